set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_p2p.py -q -m gpu -s 2>&1 | grep -E "^\.*[12]D stress|passed|failed|FAILED|Error" | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2s_n8.log 2> gpurun_out/bench_r2s_n8.err; grep -E "e2e phases|Error|error" gpurun_out/bench_r2s_n8.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2s_n4.log 2> gpurun_out/bench_r2s_n4.err; grep -E "e2e phases|Error|error" gpurun_out/bench_r2s_n4.err | cut -c1-300
timeout 300 python tools/one_call_multi.py --reps 3 --sweep "GPUS=8" > gpurun_out/one_call_r2s_n8.log 2>&1; grep -E "one call|summary|Error|error" gpurun_out/one_call_r2s_n8.log | cut -c1-400
python - <<'PY'
import json
for f in ["bench_r2s_n8.log", "bench_r2s_n4.log"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value", round(d["value"] / 1e9, 2), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["launch_ms"], 3), "gap", round(d["roofline"]["step_ms_minus_kernel_ms"], 3),
          "e2e", round(d["e2e"]["seconds"], 3), "stress", d["e2e"]["stress_mean_abs_rel"], d["e2e"]["stress_rms_rel"], d["launch"]["overlapped_reconcile"])
    for a in d.get("also") or []:
        print("   also", a.get("workload"), round(a["value"] / 1e9, 2), a.get("ms_per_step"), a["roofline"]["launch_ms"], a["e2e"]["seconds"], a["e2e"]["stress_mean_abs_rel"])
PY
