"""Exploration: spread of the sampled stress after a SHORT schedule (unconverged layouts), GPU iid / sweep vs oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from oracle import oracle as O
N, P = int(sys.argv[1]), int(sys.argv[2])
iters = [int(v) for v in sys.argv[3].split(",")]
s = G.SynthGraph(N, P, seed=42)
og = O.Graph.from_dense(s.step_handles, s.path_first.copy(), s.node_len)
graph = G.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
ix = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
seeds = [9399220 + 1000 * k for k in range(7)]
def summ(name, vals):
    r = np.array([v[0] for v in vals]); m = np.array([v[1] for v in vals])
    print(f"{name:40s} mean_abs med {np.median(m):.5f} [{m.min():.5f},{m.max():.5f}]  rms med {np.median(r):.5f} [{r.min():.5f},{r.max():.5f}]", flush=True)
for it in iters:
    for nt in (1, 16):
        vals = []
        for sd in seeds:
            op = O.params_from_graph(og, nthreads=nt); op.seed = sd; op.iter_max = it
            x, st, _ = O.path_linear_sgd(og, op, mode=O.MODE_EXACT)
            vals.append(G.sort_stress(graph, x, 200000, ix))
        summ(f"iter_max {it} oracle exact threads={nt}", vals)
    op = O.params_from_graph(og); op.iter_max = it
    for win, thr in (("0", 0), ("0", 2048), ("8192", 0), ("8192", 2048)):
        os.environ["GFASORT_WINDOW"] = win
        vals = []
        for sd in seeds:
            kw = {n: getattr(op, n) for n, _ in op._fields_}; kw["progress"] = False; kw["seed"] = sd
            cfg = G.LaunchCfg.default(); cfg.total_threads = thr
            x = G.path_linear_sgd_array(graph, G.PathSGDParams(**kw), ix, cfg)
            vals.append(G.sort_stress(graph, x, 200000, ix))
        summ(f"iter_max {it} gpu window={win} grid={G.sgd.last_stats['grid']}x{G.sgd.last_stats['block']}", vals)
