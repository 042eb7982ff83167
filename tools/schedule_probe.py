"""Full-budget `Y` runs of one synthetic graph under several sampling schedules, a few seeds each: sampled path stress
(mean and RMS form, same 1M-sample Philox sample as everywhere) and kernel time.  Characterises what the default
(sweep + coherent) schedule costs or gains against the reference's own sampling (GFASORT_WINDOW=0), and how much the
RMS form moves from seed to seed.

    python tools/schedule_probe.py [--nodes 10000000 --paths 90] --settings "WINDOW=-1,COHERENT=32;WINDOW=-1,COHERENT=8;WINDOW=0" --seeds 3
"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from gfasort_b200._cabi import Stats, check, f64p, lib

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--paths", type=int, default=90)
ap.add_argument("--seeds", type=int, default=3)
ap.add_argument("--iter-max", type=int, default=100)
ap.add_argument("--settings", default="WINDOW=-1,COHERENT=32;WINDOW=-1,COHERENT=8;WINDOW=-1,COHERENT=0;WINDOW=0")
ap.add_argument("--oracle-x", default="", help="an oracle result (x.npy) to evaluate on the same sample")
ap.add_argument("--graph", default="synth", choices=["synth", "hard"], help="hard: tests/hard_graph.py (tiled / perturbed DRB1-3123)")
ap.add_argument("--tiles", type=int, default=150)
ap.add_argument("--oracle", type=int, default=0, help="also run the CPU oracle (exact budget, this many threads) on the same seeds")
a = ap.parse_args()


class _Flat:
    pass


if a.graph == "hard":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from hard_graph import tiled_drb1
    s = _Flat()
    s.step_handles, s.path_first, s.node_len = tiled_drb1(G, a.tiles)
    s.N, s.P, s.S = len(s.node_len), len(s.path_first) - 1, len(s.step_handles)
    x_init = np.zeros(s.N); np.cumsum(s.node_len[:-1], dtype=np.float64, out=x_init[1:])
    s.initial_positions = lambda: x_init
else:
    s = G.SynthGraph(a.nodes, a.paths, seed=42)
ix = G.PathIndex.from_arrays(s.step_handles.astype(np.uint32), s.path_first, s.node_len)
counts = np.diff(s.path_first)
p = G.PathSGDParams(iter_max=a.iter_max, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                    space=int(ix.path_lengths().max()), space_max=100)
x0 = s.initial_positions()
print(f"N={s.N} P={s.P} S={s.S}; init stress {G.layout_stress(None, x0, 1, 1_000_000, ix, layout_order=False)}", flush=True)
if a.oracle_x and os.path.exists(a.oracle_x):
    xo = np.load(a.oracle_x)
    print(f"oracle x ({a.oracle_x}): stress (rms, mean_abs, n) {G.layout_stress(None, xo, 1, 1_000_000, ix, layout_order=False)}", flush=True)
out = []
for setting in a.settings.split(";"):
    for kv in setting.split(","):
        k, v = kv.split("=")
        os.environ["GFASORT_" + k] = v
    rows = []
    for k in range(a.seeds):
        from dataclasses import replace
        q = replace(p, seed=9399220 + 1000 * k)
        x = x0.copy()
        st = Stats()
        cp = q.c()
        check(lib().gfs_sgd_1d(ix.handle, C.byref(cp), x.ctypes.data_as(f64p), C.byref(st)))
        r = G.layout_stress(None, x, 1, 1_000_000, ix, layout_order=False)
        rows.append((r[1], r[0], st.kernel_seconds, st.window_steps, st.coherent))
        print(f"  [{setting}] seed {q.seed}: mean_abs {r[1]:.4e} rms {r[0]:.4e}  kernel {st.kernel_seconds:.2f}s "
              f"({st.applied_updates/st.kernel_seconds/1e9:.1f} G upd/s) window {st.window_steps} coherent {st.coherent}", flush=True)
    m = np.array(rows)
    print(f"[{setting}] median mean_abs {np.median(m[:,0]):.4e} [{m[:,0].min():.4e}, {m[:,0].max():.4e}]  "
          f"median rms {np.median(m[:,1]):.4e} [{m[:,1].min():.4e}, {m[:,1].max():.4e}]  kernel {np.median(m[:,2]):.2f}s", flush=True)
    out.append({"setting": setting, "mean_abs": m[:, 0].tolist(), "rms": m[:, 1].tolist(), "kernel_s": m[:, 2].tolist()})
    for kv in setting.split(","):
        os.environ.pop("GFASORT_" + kv.split("=")[0], None)
if a.oracle:
    from oracle import oracle as O
    og = O.Graph.from_dense(np.asarray(s.step_handles, dtype=np.uint64), np.asarray(s.path_first).copy(), s.node_len)
    rows = []
    for k in range(a.seeds):
        op = O.params_from_graph(og, nthreads=a.oracle); op.iter_max = a.iter_max; op.seed = 9399220 + 1000 * k
        xo, ost, _ = O.path_linear_sgd(og, op, mode=O.MODE_EXACT)
        r = G.layout_stress(None, xo, 1, 1_000_000, ix, layout_order=False)
        rows.append((r[1], r[0]))
        print(f"  [oracle] seed {op.seed}: mean_abs {r[1]:.4e} rms {r[0]:.4e} ({ost.applied/ost.seconds/1e6:.0f} M upd/s)", flush=True)
    m = np.array(rows)
    print(f"[oracle exact, {a.oracle} threads] median mean_abs {np.median(m[:,0]):.4e} [{m[:,0].min():.4e}, {m[:,0].max():.4e}]  "
          f"median rms {np.median(m[:,1]):.4e} [{m[:,1].min():.4e}, {m[:,1].max():.4e}]", flush=True)
    out.append({"setting": "oracle", "mean_abs": m[:, 0].tolist(), "rms": m[:, 1].tolist()})
print(json.dumps(out))
