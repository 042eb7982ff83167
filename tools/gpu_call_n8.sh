set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2g_n8.log 2> gpurun_out/bench_r2g_n8.err; tail -12 gpurun_out/bench_r2g_n8.err; cut -c1-1200 gpurun_out/bench_r2g_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2g_n4.log 2> gpurun_out/bench_r2g_n4.err; tail -3 gpurun_out/bench_r2g_n4.err; cut -c1-300 gpurun_out/bench_r2g_n4.log
GFASORT_GPUS=8 timeout 300 python tools/one_call_multi.py > gpurun_out/one_call_r2g_n8.log 2>&1; cat gpurun_out/one_call_r2g_n8.log
