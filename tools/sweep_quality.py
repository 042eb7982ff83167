"""Exploration: stress of the sweep schedule on a small graph vs threads / chunk / partner group (5 seeds, medians)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from oracle import oracle as O
N, P = int(sys.argv[1]), int(sys.argv[2])
iters = [int(v) for v in sys.argv[3].split(",")]
settings = sys.argv[4].split(";")     # "window,chunk,group,threads"
s = G.SynthGraph(N, P, seed=42)
og = O.Graph.from_dense(s.step_handles, s.path_first.copy(), s.node_len)
graph = G.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
ix = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
seeds = [9399220 + 1000 * k for k in range(5)]
def summ(name, vals):
    r = np.array([v[0] for v in vals]); m = np.array([v[1] for v in vals])
    print(f"{name:58s} mean_abs med {np.median(m):.5f} [{m.min():.5f},{m.max():.5f}]  rms med {np.median(r):.5f} [{r.min():.5f},{r.max():.5f}]", flush=True)
for it in iters:
    vals = []
    for sd in seeds:
        op = O.params_from_graph(og, nthreads=16); op.seed = sd; op.iter_max = it
        x, st, _ = O.path_linear_sgd(og, op, mode=O.MODE_EXACT)
        vals.append(G.sort_stress(graph, x, 200000, ix))
    summ(f"iter_max {it} oracle exact threads=16", vals)
    op = O.params_from_graph(og); op.iter_max = it
    for st in settings:
        win, chunk, group, thr = st.split(",")
        os.environ["GFASORT_WINDOW"] = win; os.environ["GFASORT_CHUNK"] = chunk; os.environ["GFASORT_PARTNER_GROUP"] = group
        vals = []
        for sd in seeds:
            kw = {n: getattr(op, n) for n, _ in op._fields_}; kw["progress"] = False; kw["seed"] = sd
            cfg = G.LaunchCfg.default(); cfg.total_threads = int(thr)
            x = G.path_linear_sgd_array(graph, G.PathSGDParams(**kw), ix, cfg)
            vals.append(G.sort_stress(graph, x, 200000, ix))
        summ(f"iter_max {it} gpu win={win} chunk={chunk} group={group} grid={G.sgd.last_stats['grid']}x{G.sgd.last_stats['block']}", vals)
