"""Stress reached after a full schedule for several GPU scheduling configurations, next to the
CPU oracle at the same update budget (exploration tool, not a test)."""
import argparse, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--paths", type=int, default=32)
ap.add_argument("--iter-max", type=int, default=100)
ap.add_argument("--windows", type=str, default="0,262144,1048576")
ap.add_argument("--threads", type=str, default="0")
ap.add_argument("--oracle", type=int, default=0, help="oracle threads (0 = skip)")
ap.add_argument("--dims", type=int, default=0)
ap.add_argument("--samples", type=int, default=1_000_000)
a = ap.parse_args()

s = G.SynthGraph(a.nodes, a.paths, seed=42)
graph = G.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
ix = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
counts = np.diff(s.path_first)
if a.dims == 0:
    p = G.PathSGDParams(iter_max=a.iter_max, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                        space=int(ix.path_lengths().max()), space_max=100)
else:
    p = G.LayoutSGDParams(dimensions=a.dims, iter_max=a.iter_max, min_term_updates=10 * int(counts.sum()),
                          eta_max=float(int(counts.max()) ** 2), space=int(counts.max()), space_max=1000)
x0 = s.initial_positions()
print(f"N={s.N} P={s.P} S={s.S} init stress {G.sort_stress(graph, x0, a.samples, ix)}", flush=True)
for thr in [int(t) for t in a.threads.split(",")]:
    for w in [int(v) for v in a.windows.split(",")]:
        os.environ["GFASORT_WINDOW"] = str(w)
        cfg = G.LaunchCfg.default(); cfg.total_threads = thr
        t = time.time()
        if a.dims == 0:
            x = G.path_linear_sgd_array(graph, p, ix, cfg)
            st = G.sort_stress(graph, x, a.samples, ix)
        else:
            lay = G.path_linear_sgd_layout(graph, p, ix, cfg)
            st = G.layout_stress(graph, lay.coords, a.dims, a.samples, ix)
        ls = dict(G.sgd.last_stats)
        print(f"gpu window={w} threads={ls['grid']}x{ls['block']}: mean_abs {st[1]:.6f} rms {st[0]:.6f}  "
              f"({ls['applied_updates']/ls['kernel_seconds']/1e9:.2f} G upd/s, {time.time()-t:.1f}s)", flush=True)
if a.oracle:
    from oracle import oracle as O
    og = O.Graph.from_dense(s.step_handles, s.path_first, s.node_len)
    op = O.params_from_graph(og, layout=a.dims > 0, nthreads=a.oracle); op.iter_max = a.iter_max
    t = time.time()
    if a.dims == 0:
        xo, ost, _ = O.path_linear_sgd(og, op, mode=O.MODE_EXACT)
        st = G.sort_stress(graph, xo, a.samples, ix)
    else:
        co, ost, _ = O.path_linear_sgd_layout(og, op, a.dims, mode=O.MODE_EXACT)
        st = G.layout_stress(graph, co, a.dims, a.samples, ix)
    print(f"oracle exact {a.oracle} threads: mean_abs {st[1]:.6f} rms {st[0]:.6f} ({ost.applied/ost.seconds/1e6:.1f} M upd/s, {time.time()-t:.1f}s)", flush=True)
