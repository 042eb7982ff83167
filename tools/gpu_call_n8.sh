# The scaling run on one 8-GPU box:  gpurun --timeout 1500 --gpus 8 -- 'bash tools/gpu_call_n8.sh'
set -x
cd $GRAFT_REPO_ROOT
for N in 8 4 2; do
  ALSO=$([ $N = 8 ] && echo 1 || echo 0)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --also $ALSO \
      > gpurun_out/bench_scale_n$N.log 2> gpurun_out/bench_scale_n$N.err
  grep -E "e2e phases|Error|error" gpurun_out/bench_scale_n$N.err | cut -c1-300
  cut -c1-400 gpurun_out/bench_scale_n$N.log
done
