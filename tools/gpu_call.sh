set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -s -k "p2p or multi_gpu or replica" > gpurun_out/pytest_r2l_n2.log 2>&1; tail -4 gpurun_out/pytest_r2l_n2.log; grep "stress one GPU" gpurun_out/pytest_r2l_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2l_n2_overlap.log 2> gpurun_out/bench_r2l_n2_overlap.err; grep "e2e phases" gpurun_out/bench_r2l_n2_overlap.err | cut -c1-250
GFASORT_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2l_n2_sync.log 2> gpurun_out/bench_r2l_n2_sync.err; grep "e2e phases" gpurun_out/bench_r2l_n2_sync.err | cut -c1-250
python - <<'PY'
import json
for f in ["bench_r2l_n2_overlap.log", "bench_r2l_n2_sync.log"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value", round(d["value"] / 1e9, 2), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["launch_ms"], 3), "gap", round(d["roofline"]["step_ms_minus_kernel_ms"], 3),
          "e2e", round(d["e2e"]["seconds"], 3), "stress", d["e2e"]["stress_mean_abs_rel"], d["e2e"]["stress_rms_rel"])
    for a in d["also"]:
        print("   also", a.get("workload"), round(a["value"] / 1e9, 2) if "value" in a else a, a.get("ms_per_step"), a["e2e"]["stress_mean_abs_rel"] if "e2e" in a else None)
PY
tail -3 gpurun_out/bench_r2l_n2_overlap.err
