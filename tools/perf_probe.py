"""Exploratory throughput probe (not the bench): times index build and warm / cooling epochs of the
1D and 2D SGD kernels on a synthetic graph for several launch configurations."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from gfasort_b200._cabi import lib, check, f64p, Stats, LaunchCfg

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--paths", type=int, default=90)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--dims", type=int, nargs="*", default=[0])
ap.add_argument("--threads", type=int, nargs="*", default=[0])
ap.add_argument("--agg", type=int, nargs="*", default=[1, 0])
ap.add_argument("--f64", type=int, nargs="*", default=[0])
ap.add_argument("--stress", action="store_true")
ap.add_argument("--sweep", type=str, default="", help="';'-separated env settings, each 'K=V,K=V', applied in turn")
a = ap.parse_args()

print(lib().gfs_device_info().decode(), flush=True)
t = time.time(); s = G.SynthGraph(a.nodes, a.paths, seed=42); print(f"synth N={s.N} P={s.P} S={s.S} in {time.time()-t:.2f}s", flush=True)
t = time.time(); ix = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len); t_ix = time.time() - t
print(f"index build {t_ix:.3f}s ({s.S/t_ix/1e9:.2f} Gsteps/s incl. H2D of {s.S*8/1e9:.2f} GB pageable)", flush=True)
t = time.time(); ix2 = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len); print(f"index build (2nd) {time.time()-t:.3f}s"); ix2.close()
counts = np.diff(s.path_first)
lens = ix.path_lengths()
x0 = s.initial_positions()

def run(dims, threads, agg, f64):
    layout = dims > 0
    if layout:
        p = G.LayoutSGDParams(dimensions=dims, iter_max=30, min_term_updates=10 * int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                              space=int(counts.max()), space_max=1000)
    else:
        p = G.PathSGDParams(iter_max=100, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2), space=int(lens.max()), space_max=100)
    cp = p.c()
    cfg = LaunchCfg.default(); cfg.total_threads = threads; cfg.aggregate = agg; cfg.layout_f64 = f64
    h = C.c_void_p()
    check(lib().gfs_sgd_session_create(ix.handle, C.byref(cp), dims, C.byref(cfg), C.byref(h)))
    if layout:
        pos = np.zeros((s.N, 2, dims)); pos[:, 0, 0] = x0; pos[:, 1, 0] = x0 + s.node_len
        pos[:, :, 1:] = np.random.default_rng(1).standard_normal((s.N, 2, dims - 1)) * np.sqrt(2.0 * s.N)
        pos = pos.reshape(-1)
    else:
        pos = x0.copy()
    check(lib().gfs_sgd_session_upload(h, pos.ctypes.data_as(f64p)))
    st = Stats(); res = {}
    first_cool = int(np.floor(p.cooling_start * p.iter_max)) + 1
    n_slices = 10 if layout else 1
    for name, e0 in (("warm", 0), ("cool", first_cool)):
        check(lib().gfs_sgd_session_stats(h, C.byref(st))); k0, a0 = st.kernel_seconds, st.applied_updates
        for e in range(e0, e0 + a.epochs):
            check(lib().gfs_sgd_session_run(h, e, e + 1, 0, n_slices))
        check(lib().gfs_sgd_session_stats(h, C.byref(st)))
        res[name] = (st.applied_updates - a0) / (st.kernel_seconds - k0)
    print(f"dims={dims} threads={st.grid}x{st.block} agg={agg} f64={f64}: warm {res['warm']/1e9:.3f} G upd/s, cool {res['cool']/1e9:.3f} G upd/s", flush=True)
    lib().gfs_sgd_session_destroy(h)

for setting in (a.sweep.split(";") if a.sweep else [""]):
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    if setting:
        print(f"== {setting}", flush=True)
    for dims in a.dims:
        for f64 in (a.f64 if dims else [1]):
            for threads in a.threads:
                for agg in a.agg:
                    run(dims, threads, agg, f64)
ix.close()
