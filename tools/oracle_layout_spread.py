"""Run-to-run spread of the CPU oracle's 2D layout (`L`) on the graph tests/test_gpu_p2p.py::test_one_call_multi_gpu_through_the_cabi[2]
uses (synthetic, 200k nodes / 16 paths; layout-iter 30, 10 S updates per epoch), evaluated on gfs_stress's Philox sample.
Test infrastructure: the oracle only, nothing of the product library is loaded.

    python tools/oracle_layout_spread.py [--nodes 200000 --paths 16 --seeds 5 --threads 8]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O, synth_host

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=200_000)
ap.add_argument("--paths", type=int, default=16)
ap.add_argument("--seeds", type=int, default=5)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--samples", type=int, default=500_000)
a = ap.parse_args()
h, first, nlen = synth_host.synth_arrays(a.nodes, a.paths, 42)
g = O.Graph.from_dense(h, first, nlen)
counts = np.diff(first)
mx = int(counts.max())
rows = []
for k in range(a.seeds):
    p = O.params_from_graph(g, layout=True, nthreads=a.threads)
    p.iter_max = 30; p.min_term_updates = 10 * int(counts.sum()); p.eta_max = float(mx * mx); p.space = mx; p.space_max = 1000
    p.space_quantization_step = 100; p.seed = 9399220 + 1000 * k; p.nthreads = a.threads
    c0 = O.init_layout(g, 2, 9399220)
    t = time.perf_counter()
    c, st, rc = O.path_linear_sgd_layout(g, p, 2, mode=O.MODE_EXACT, coords0=c0)
    dt = time.perf_counter() - t
    r = O.layout_stress(g, c, 2, a.samples, draw=O.DRAW_PHILOX, seed=12345)
    rows.append({"seed": int(p.seed), "applied": int(st.applied), "seconds": dt, "rms_rel": float(r[0]), "mean_abs_rel": float(r[1]), "counted": int(r[2])})
    print(rows[-1], flush=True)
m = np.array([r["mean_abs_rel"] for r in rows])
print(json.dumps({"nodes": a.nodes, "paths": a.paths, "threads": a.threads, "median_mean_abs_rel": float(np.median(m)), "min": float(m.min()), "max": float(m.max()), "runs": rows}))
