"""CPU test (-m "not gpu") of bench.py's output contract: the reference arm (which needs no GPU) prints
ONE JSON line with the agreed keys; the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--workload", "ytiny", "--steps", "2", "--warmup", "1", "--cpu-sample", "600000")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                     # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sgd_term_updates_per_sec" and d["unit"] == "updates/s"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["scaling"] in ("weak", "strong") and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "ytiny"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                                                 # on a GPU box the arm simply runs (test_gpu_parity covers it)
    r = _run("--workload", "ytiny", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
