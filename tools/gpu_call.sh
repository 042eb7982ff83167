set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2w_n8.log 2> gpurun_out/bench_r2w_n8.err; echo rc=$?; grep -E "Error|error|Traceback" gpurun_out/bench_r2w_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2w_n8.log").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e9, 2), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["launch_ms"], 3), "by rank", d["roofline"]["launch_ms_by_rank"], "gap", round(d["roofline"]["step_ms_minus_kernel_ms"], 3), "stress", d["e2e"]["stress_mean_abs_rel"])
PY
