"""Index-build probe: builds the path index (K1 + relabelling) of a synthetic graph from pageable / pinned host
memory with 64- / 32-bit handles and prints the phases of every build (gfs_index_build_info).  Run it plain for wall
times and under `ncu --metrics gpu__time_duration.sum --clock-control none --csv` for the per-kernel launch list."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--paths", type=int, default=90)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--modes", default="pageable32,pageable64,pinned32,pinned64")
a = ap.parse_args()
for mode in a.modes.split(","):
    pinned, bits = mode.startswith("pinned"), int(mode[-2:])
    t = time.time(); s = G.SynthGraph(a.nodes, a.paths, seed=42, pinned=pinned and bits == 64)
    h = s.step_handles
    keep = None
    if bits == 32:
        if pinned:
            import torch
            keep = torch.empty(s.S, dtype=torch.int32, pin_memory=True)
            h = keep.numpy().view(np.uint32); h[:] = s.step_handles
        else:
            h = s.step_handles.astype(np.uint32)
    print(f"[{mode}] synth N={s.N} P={s.P} S={s.S} in {time.time()-t:.2f}s; step array {h.nbytes/1e9:.2f} GB", flush=True)
    for r in range(a.reps):
        t = time.time(); ix = G.PathIndex.from_arrays(h, s.path_first, s.node_len, env=True); dt = time.time() - t
        bi = ix.build_info()
        print(f"[{mode}] build #{r}: {dt:.3f}s wall = {s.S/dt/1e9:.2f} G steps/s; PCIe-copy-only bound at 55 GB/s {h.nbytes/55e9:.3f}s; "
              f"K1 {bi['kernel_seconds']*1e3:.2f} ms = {s.S*20/bi['kernel_seconds']/1e9:.0f} GB/s on 20 B/step; " + json.dumps(bi), flush=True)
        ix.close()
    s.close()
