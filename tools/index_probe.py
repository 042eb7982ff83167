"""Index-build probe: generates the config-3 graph in pinned host memory and builds the path index
(K1 + relabelling) `--reps` times.  Run it plain for wall times and under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` for the per-kernel launch list."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfasort_b200 as G

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--paths", type=int, default=90)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
t = time.time(); s = G.SynthGraph(a.nodes, a.paths, seed=42, pinned=True)
print(f"synth N={s.N} P={s.P} S={s.S} in {time.time()-t:.2f}s (pinned)", flush=True)
for r in range(a.reps):
    t = time.time(); ix = G.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len); dt = time.time() - t
    print(f"index build #{r}: {dt:.3f}s = {s.S/dt/1e9:.2f} G steps/s incl. H2D of {s.S*8/1e9:.2f} GB", flush=True)
    ix.close()
s.close()
