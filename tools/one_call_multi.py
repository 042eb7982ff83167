"""The Rust host's path on G GPUs: ONE process, GFASORT_GPUS=G, gfs_index_build32 + gfs_sgd_1d + gfs_stress
(INTEGRATION.md §4).  Prints the phases and the all-paths stress.

    GFASORT_GPUS=8 python tools/one_call_multi.py [--nodes 10000000 --paths 90] [--dims 0|2]
"""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G
from gfasort_b200._cabi import Stats, check, f64p, lib

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--paths", type=int, default=90)
ap.add_argument("--dims", type=int, default=0)
ap.add_argument("--pinned", type=int, default=0)
ap.add_argument("--reps", type=int, default=2, help="runs; run k uses SGD seed 9399220 + 1000 k")
ap.add_argument("--sweep", default="", help="comma list of GPUS:OVERLAP pairs run one after the other in this process, e.g. 8:0,8:2,4:0")
a = ap.parse_args()
s = G.SynthGraph(a.nodes, a.paths, seed=42)
if a.pinned:
    from gfasort_b200.sgd import PinnedArray
    keep = PinnedArray(s.S, np.uint32); h = keep.array; h[:] = s.step_handles
else:
    h = s.step_handles.astype(np.uint32)
counts = np.diff(s.path_first)
x0 = s.initial_positions()
# --sweep "GPUS=8,OVERLAP=0;GPUS=8,OVERLAP=2;GPUS=1": configurations run one after the other in this process; every KEY=VAL
# sets GFASORT_KEY for that configuration (keys of earlier configurations are removed); the old form 8:0,8:2 (GPUS:OVERLAP) still works
def _parse(item):
    if "=" in item:
        return dict(kv.split("=") for kv in item.split(","))
    g, o = item.split(":")
    return {"GPUS": g, "OVERLAP": o}
if a.sweep:
    sweep = [_parse(c) for c in (a.sweep.split(";") if "=" in a.sweep else a.sweep.split(","))]
else:
    sweep = [None]
swept = set()
for cfg in sweep:
  if cfg is not None:
      for k in swept:
          os.environ.pop("GFASORT_" + k, None)
      for k, v in cfg.items():
          os.environ["GFASORT_" + k] = v
          swept.add(k)
  rows = []
  for rep in range(a.reps):
      t0 = time.perf_counter()
      ix = G.PathIndex.from_arrays(h, s.path_first, s.node_len, env=True)
      t1 = time.perf_counter()
      mx = int(counts.max())
      if a.dims == 0:
          p = G.PathSGDParams(iter_max=100, min_term_updates=int(counts.sum()), eta_max=float(mx * mx), space=int(ix.path_lengths().max()), space_max=100,
                              seed=9399220 + 1000 * rep)
          x = x0.copy()
      else:
          p = G.LayoutSGDParams(dimensions=a.dims, iter_max=30, min_term_updates=10 * int(counts.sum()), eta_max=float(mx * mx), space=mx, space_max=1000,
                                seed=9399220 + 1000 * rep)
          c = np.zeros((s.N, 2, a.dims)); c[:, 0, 0] = x0; c[:, 1, 0] = x0 + s.node_len
          c[:, :, 1:] = np.random.default_rng(1).standard_normal((s.N, 2, a.dims - 1)) * np.sqrt(2.0 * s.N)
          x = c.reshape(-1)
      st = Stats(); cp = p.c()
      t2 = time.perf_counter()
      if a.dims == 0:
          check(lib().gfs_sgd_1d(ix.handle, C.byref(cp), x.ctypes.data_as(f64p), C.byref(st)))
      else:
          check(lib().gfs_sgd_nd(ix.handle, C.byref(cp), a.dims, x.ctypes.data_as(f64p), C.byref(st)))
      t3 = time.perf_counter()
      stress = G.layout_stress(None, x, max(a.dims, 1), 1_000_000, ix, layout_order=a.dims > 0)
      upd = (p.iter_max + 1) * p.min_term_updates
      assert st.applied_updates == upd
      print(f"[one call, GFASORT_GPUS={os.environ.get('GFASORT_GPUS', '1')}, rep {rep}] devices {st.n_devices}: index build {t1-t0:.3f}s {ix.build_info()} | "
            f"sgd {t3-t2:.3f}s (kernel max {st.kernel_seconds:.3f}s) | e2e {t3-t0:.3f}s = {upd/(t3-t0)/1e9:.1f} G upd/s | "
            f"stress over all paths: mean_abs {stress[1]:.4e} rms {stress[0]:.4e} n {stress[2]}", flush=True)
      rows.append((stress[1], stress[0], t3 - t2, st.syncs_per_epoch))
      ix.close()
  m = np.array([r[:3] for r in rows])
  print(f"[summary {cfg if cfg is not None else ''} GFASORT_GPUS={os.environ.get('GFASORT_GPUS', '1')} OVERLAP={os.environ.get('GFASORT_OVERLAP', '0')} SYNCS={os.environ.get('GFASORT_SYNCS', 'auto')}"
        f" -> {rows[-1][3]}/epoch; dims {a.dims}; {a.reps} runs] mean_abs median {np.median(m[:,0]):.4e} [{m[:,0].min():.4e}, {m[:,0].max():.4e}]  "
        f"rms median {np.median(m[:,1]):.4e}  sgd seconds median {np.median(m[:,2]):.3f}", flush=True)
