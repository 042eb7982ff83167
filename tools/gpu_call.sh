set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_p2p.py -q -m gpu -s 2>&1 | grep -E "stress|passed|failed|FAILED|Error" | cut -c1-600
GFASORT_RC_TRACE=1 timeout 600 python tools/one_call_multi.py --reps 2 --sweep 2:0,2:1,2:2,2:3 > gpurun_out/one_call_r2p_n2.log 2>&1
grep -E "summary|rc trace.*rank 0|Error|error" gpurun_out/one_call_r2p_n2.log | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2p_n2.log 2> gpurun_out/bench_r2p_n2.err; grep -E "e2e phases|Error|error" gpurun_out/bench_r2p_n2.err | cut -c1-300
GFASORT_OVERLAP=1 GFASORT_RC_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --also 0 > gpurun_out/bench_r2p_n2_ov1.log 2> gpurun_out/bench_r2p_n2_ov1.err; grep -E "e2e phases|rc trace|Error|error" gpurun_out/bench_r2p_n2_ov1.err | cut -c1-300
python - <<'PY'
import json
for f in ["bench_r2p_n2.log", "bench_r2p_n2_ov1.log"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value", round(d["value"] / 1e9, 2), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["launch_ms"], 3), "gap", round(d["roofline"]["step_ms_minus_kernel_ms"], 3),
          "e2e", round(d["e2e"]["seconds"], 3), "stress", d["e2e"]["stress_mean_abs_rel"], d["e2e"]["stress_rms_rel"], d["launch"]["overlapped_reconcile"])
PY
