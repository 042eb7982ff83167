"""A short, fixed sequence of SGD launches for ncu: builds the index of one workload, uploads, runs `--launches`
slices (1/--slices of an epoch each) of a warm epoch and of a cooling epoch.  Small launches keep ncu's ~40 replays
per kernel affordable at config 5 (a whole epoch there is 8.3e9 updates).

    python tools/ncu_target.py --workload y100m --slices 40 --launches 2        # plain first, then under ncu
"""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from gfasort_b200._cabi import Stats, check, f64p, lib

WL = {"y10m": (10_000_000, 90, 0), "l10m": (10_000_000, 90, 2), "y100m": (100_000_000, 90, 0), "y1m": (1_000_000, 32, 0)}
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="y10m", choices=sorted(WL))
ap.add_argument("--slices", type=int, default=20)
ap.add_argument("--launches", type=int, default=2)
a = ap.parse_args()
nodes, paths, dims = WL[a.workload]
t = time.time()
s = G.SynthGraph(nodes, paths, seed=42)
h32 = s.step_handles.astype(np.uint32)
ix = G.PathIndex.from_arrays(h32, s.path_first, s.node_len)
counts = np.diff(s.path_first)
mx = int(counts.max())
if dims == 0:
    p = G.PathSGDParams(iter_max=100, min_term_updates=int(counts.sum()), eta_max=float(mx * mx), space=int(ix.path_lengths().max()), space_max=100)
    pos = s.initial_positions()
else:
    p = G.LayoutSGDParams(dimensions=dims, iter_max=30, min_term_updates=10 * int(counts.sum()), eta_max=float(mx * mx), space=mx, space_max=1000)
    x0 = s.initial_positions()
    pos = np.zeros((s.N, 2, dims)); pos[:, 0, 0] = x0; pos[:, 1, 0] = x0 + s.node_len
    pos[:, :, 1:] = np.random.default_rng(1).standard_normal((s.N, 2, dims - 1)) * np.sqrt(2.0 * s.N)
    pos = pos.reshape(-1)
print(f"{a.workload}: N={s.N} S={s.S} ready in {time.time()-t:.1f}s", flush=True)
cp = p.c()
h = C.c_void_p()
check(lib().gfs_sgd_session_create(ix.handle, C.byref(cp), dims, None, C.byref(h)))
check(lib().gfs_sgd_session_upload(h, pos.ctypes.data_as(f64p)))
first_cool = int(np.floor(p.cooling_start * p.iter_max)) + 1
st = Stats()
for name, e in (("warm", 1), ("cool", first_cool + 1)):
    check(lib().gfs_sgd_session_stats(h, C.byref(st))); k0, a0 = st.kernel_seconds, st.applied_updates
    for k in range(a.launches):
        if a.slices == 1:
            check(lib().gfs_sgd_session_run(h, e + k, e + k + 1, 0, 1))       # whole epochs: consecutive epochs
        else:
            check(lib().gfs_sgd_session_run(h, e, e + 1, k, a.slices))
    check(lib().gfs_sgd_session_stats(h, C.byref(st)))
    print(f"{name}: {a.launches} launches of {(st.applied_updates-a0)//a.launches} updates: {(st.applied_updates-a0)/(st.kernel_seconds-k0)/1e9:.2f} G upd/s", flush=True)
lib().gfs_sgd_session_destroy(h)
ix.close()
