"""CPU emulation of a replicated `L` run with the oracle as the per-replica worker: G replicas, each samples only
its own paths for one epoch from the common base, then the replicas are combined and the next epoch starts from the result.
Lets the COMBINE RULE be studied at any G without GPUs (profiles/r2_experiments.md §6).  Test infrastructure only.

    python tools/replica_emulation.py --G 2 --rule mean --seeds 3
rules:  mean   x = base + sum_g d_g / #moved                    (the library's moved-replica mean)
        sum    x = base + sum_g d_g
        pow:A  x = base + sum_g d_g / #moved^A
        omega:W  x = base + W * sum_g d_g / #moved   (elements that at least two replicas moved)
        anneal mean while the epoch is warm, then the divisor falls linearly in the epoch index to 1 at the last epoch
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O, synth_host

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=200_000)
ap.add_argument("--paths", type=int, default=16)
ap.add_argument("--G", type=int, default=2)
ap.add_argument("--rule", default="mean")
ap.add_argument("--seeds", type=int, default=3)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--dims", type=int, default=2)
ap.add_argument("--iter-max", type=int, default=30)
ap.add_argument("--syncs", type=int, default=1, help="combines per epoch (the epoch's updates are split evenly)")
a = ap.parse_args()

h, first, nlen = synth_host.synth_arrays(a.nodes, a.paths, 42)
full = O.Graph.from_dense(h, first, nlen)
counts = np.diff(first).astype(np.int64)
mx = int(counts.max())
P = len(counts)
# contiguous, step-balanced sets of whole paths
bounds = [0]
tot = int(counts.sum())
acc = 0
for p in range(P):
    acc += int(counts[p])
    if len(bounds) < a.G and acc >= tot * len(bounds) / a.G:
        bounds.append(p + 1)
while len(bounds) < a.G:
    bounds.append(P)
bounds.append(P)
shards = []
for g in range(a.G):
    lo, hi = bounds[g], bounds[g + 1]
    f = first[lo:hi + 1] - first[lo]
    shards.append((O.Graph.from_dense(h[first[lo]:first[hi]], f.astype(np.uint64), nlen), int(counts[lo:hi].sum())))
print(f"N={a.nodes} P={P} S={tot}; G={a.G} shards of paths {[(bounds[g], bounds[g+1]) for g in range(a.G)]}; rule {a.rule}, {a.syncs} combines per epoch", flush=True)

def params(steps, seed):
    p = O.params_from_graph(full, layout=True, nthreads=a.threads)
    p.iter_max = a.iter_max; p.min_term_updates = 10 * steps // a.syncs; p.eta_max = float(mx * mx); p.space = mx; p.space_max = 1000
    p.space_quantization_step = 100; p.seed = seed; p.nthreads = a.threads
    return p

def divisor(moved, e):
    m = np.maximum(moved, 1).astype(np.float64)
    if a.rule == "mean":
        return m
    if a.rule == "sum":
        return np.ones_like(m)
    if a.rule.startswith("pow:"):
        return m ** float(a.rule[4:])
    if a.rule.startswith("omega:"):                        # over-relaxed mean: x = base + W * (sum_g d_g / #moved) where several moved it
        return np.where(m >= 2, m / float(a.rule[6:]), 1.0)
    if a.rule == "anneal":
        warm = int(0.5 * a.iter_max)                       # cooling starts behind this epoch (sgd.rs:393)
        t = 0.0 if e <= warm else (e - warm) / max(1, a.iter_max - warm)
        return m ** (1.0 - t)
    raise SystemExit("unknown rule")

rows = []
for k in range(a.seeds):
    base = O.init_layout(full, a.dims, 9399220)
    t0 = time.perf_counter()
    for e in range(a.iter_max + 1):
        O.set_epoch_window(e, e + 1)
        for j in range(a.syncs):
            ds = []
            for g, (sg, steps) in enumerate(shards):
                c, st, rc = O.path_linear_sgd_layout(sg, params(steps, 9399220 + 1000 * k + 100_000 * (e * a.syncs + j) + g), a.dims, mode=O.MODE_EXACT, coords0=base)
                assert st.applied == 10 * steps // a.syncs, (st.applied, steps)
                ds.append(c - base)
            moved = sum((d != 0).astype(np.int32) for d in ds)
            base = base + sum(ds) / divisor(moved, e)
    O.set_epoch_window()
    r = O.layout_stress(full, base, a.dims, 500_000, draw=O.DRAW_PHILOX, seed=12345)
    rows.append({"seed": k, "mean_abs_rel": float(r[1]), "rms_rel": float(r[0]), "seconds": time.perf_counter() - t0})
    print(rows[-1], flush=True)
m = np.array([r["mean_abs_rel"] for r in rows])
print(json.dumps({"G": a.G, "rule": a.rule, "syncs": a.syncs, "nodes": a.nodes, "median": float(np.median(m)), "min": float(m.min()), "max": float(m.max())}))
