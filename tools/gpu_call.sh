set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_r2t.log 2>&1; tail -3 gpurun_out/pytest_r2t.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_r2t.log 2> gpurun_out/bench_r2t.err; tail -c 600 gpurun_out/bench_r2t.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_r2t_ref.log 2> gpurun_out/bench_r2t_ref.err; tail -c 300 gpurun_out/bench_r2t_ref.err; cat gpurun_out/bench_r2t_ref.log | cut -c1-900
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2t.log").read().strip().splitlines()[-1])
r = d["roofline"]; e = d["e2e"]
print("value", round(d["value"] / 1e9, 2), "ms/step", round(d["ms_per_step"], 3), "kernel", round(r["launch_ms"], 3), "frac", round(r["frac"], 3), "dram_frac", r.get("dram_frac"), "e2e", round(e["value"] / 1e9, 2), round(e["seconds"], 3), e.get("phases"), "stress", e["stress_mean_abs_rel"], e["stress_rms_rel"])
print("cpu", d["cpu_baseline"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("k1", r.get("k1"))
for a in d.get("also") or []:
    print("also", a.get("workload"), round(a["value"] / 1e9, 2), a.get("ms_per_step"), a["roofline"]["frac"], a["roofline"].get("dram_frac"), a["e2e"]["seconds"], a["e2e"]["stress_mean_abs_rel"])
PY
