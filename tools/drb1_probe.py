"""Exploration: DRB1 1D stress of the GPU path for several thread counts vs the oracle (exact / reference mode)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from oracle import oracle as O

path = os.path.join(os.path.dirname(__file__), "..", "tests", "data", "DRB1-3123.gfa")
graph = G.load_gfa(path); og = O.parse_gfa(path)
ix = G.PathIndex.from_graph(graph)
seeds = [9399220 + 1000 * k for k in range(7)]
def summ(name, vals):
    r = np.array([v[0] for v in vals]); m = np.array([v[1] for v in vals])
    print(f"{name:34s} mean_abs med {np.median(m):.5f} [{m.min():.5f},{m.max():.5f}]  rms med {np.median(r):.5f} [{r.min():.5f},{r.max():.5f}]", flush=True)
for mode, nt in ((O.MODE_EXACT, 1), (O.MODE_EXACT, 16), (O.MODE_REFERENCE, 1), (O.MODE_REFERENCE, 16)):
    vals = []; app = []
    for s in seeds:
        op = O.params_from_graph(og, nthreads=nt); op.seed = s
        x, st, _ = O.path_linear_sgd(og, op, mode=mode)
        vals.append(G.sort_stress(graph, x, 200000, ix)); app.append(st.applied)
    summ(f"oracle mode={mode} threads={nt} ({np.mean(app)/1e6:.1f}M upd)", vals)
base = G.YgsParams.from_graph(graph, 0, 1, ix).path_sgd
for thr in (1, 32, 128, 256, 512, 1024, 0):
    for agg in (1,):
        vals = []
        for s in seeds:
            p = G.PathSGDParams(**{**base.__dict__, "seed": s})
            cfg = G.LaunchCfg.default(); cfg.total_threads = thr; cfg.aggregate = agg
            x = G.path_linear_sgd_array(graph, p, ix, cfg)
            vals.append(G.sort_stress(graph, x, 200000, ix))
        summ(f"gpu threads={thr} agg={agg} grid={G.sgd.last_stats['grid']}x{G.sgd.last_stats['block']}", vals)
