"""One full-budget run of the CPU oracle (the C++ restatement of reference src/sgd.rs:237-614) on a
synthetic graph of a named shape, and the sampled path stress of its result on the SAME Philox sample
gfs_stress uses on the GPU (stream 2, counter k, seed 12345) — SURVEY.md §8d: "stress-parity runs use the
full budget on DRB1, config 2 and (once) config 3".

    python tools/oracle_runs.py --name config3_10M_90                                        # ~1 h of 8 cores, 45 GB
    python tools/oracle_runs.py --name config2_1M_32 --nodes 1000000 --paths 32 --seeds 9399220,9400220,9401220

The index of the oracle is the reference's four 8-byte-per-step arrays (27 GB at config 3), so the stress
is evaluated here with numpy from per-path prefix sums instead of oracle_layout_stress (which would build
a second index).  Writes oracle/_runs/<name>/x_<seed>.npy (final positions, dense node order; not committed) and
adds the results to tests/golden/oracle_stress.json, which the -m gpu parity tests compare the CUDA path with.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def philox_stress(step_handles, path_first, node_len, x, samples, seed=12345):
    """gfs_stress / stress_kernel restated with numpy: returns (rms_rel, mean_abs_rel, counted)."""
    from oracle import oracle as O
    S = int(path_first[-1])
    P = len(path_first) - 1
    sa = np.zeros(samples, dtype=np.int64)
    r23 = np.zeros(samples, dtype=object)
    key = (seed & 0xffffffff, seed >> 32)
    for k in range(samples):
        r = O.philox((k & 0xffffffff, k >> 32, 0, O.STREAM_STRESS), key)
        sa[k] = ((int(r[1]) << 32 | int(r[0])) * S) >> 64
        r23[k] = int(r[3]) << 32 | int(r[2])
    pa = np.searchsorted(path_first, sa.astype(np.uint64), side="right") - 1
    f = path_first[pa].astype(np.int64)
    n = (path_first[pa + 1] - path_first[pa]).astype(np.int64)
    ra = sa - f
    rb = np.array([(int(r23[k]) * int(n[k])) >> 64 for k in range(samples)], dtype=np.int64)
    sb = f + rb
    ok = (n >= 2) & (ra != rb)
    pos_a = np.zeros(samples, dtype=np.float64)
    pos_b = np.zeros(samples, dtype=np.float64)
    for p in range(P):
        lo, hi = int(path_first[p]), int(path_first[p + 1])
        sel = np.nonzero(pa == p)[0]
        if not len(sel) or hi <= lo:
            continue
        lens = node_len[(step_handles[lo:hi] >> np.uint64(1)).astype(np.int64)].astype(np.uint64)
        pos = np.zeros(hi - lo, dtype=np.uint64)
        np.cumsum(lens[:-1], out=pos[1:])
        pos_a[sel] = pos[ra[sel]].astype(np.float64)
        pos_b[sel] = pos[np.minimum(rb[sel], hi - lo - 1)].astype(np.float64)
    dp = np.abs(pos_a - pos_b)
    ok &= dp != 0
    ia = (step_handles[sa] >> np.uint64(1)).astype(np.int64)
    ib = (step_handles[np.where(ok, sb, sa)] >> np.uint64(1)).astype(np.int64)
    dl = np.abs(x[ia] - x[ib])
    err = (dl - dp)[ok]
    d = dp[ok]
    cnt = int(ok.sum())
    return float(np.sqrt(np.sum(err * err / (d * d)) / cnt)), float(np.sum(np.abs(err) / d) / cnt), cnt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="config3_10M_90", help="key in tests/golden/oracle_stress.json")
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--paths", type=int, default=90)
    ap.add_argument("--graph-seed", type=int, default=42)
    ap.add_argument("--seeds", default="9399220", help="comma-separated SGD seeds, one oracle run each")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--iter-max", type=int, default=100)
    ap.add_argument("--samples", type=int, default=1_000_000)
    ap.add_argument("--mode", default="exact", choices=["exact", "reference"])
    ap.add_argument("--out", default="", help="directory for x_<seed>.npy (default oracle/_runs/<name>)")
    ap.add_argument("--fixture", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_stress.json"))
    a = ap.parse_args()
    out_dir = a.out or os.path.join("oracle", "_runs", a.name)
    os.makedirs(out_dir, exist_ok=True)
    from oracle import oracle as O
    from oracle.synth_host import synth_arrays
    t0 = time.time()
    handles, path_first, node_len = synth_arrays(a.nodes, a.paths, a.graph_seed)
    x0 = np.zeros(a.nodes)
    np.cumsum(node_len[:-1], dtype=np.float64, out=x0[1:])
    print(f"graph N={a.nodes} P={a.paths} S={len(handles)} in {time.time()-t0:.0f}s", flush=True)
    st0 = philox_stress(handles, path_first, node_len, x0, a.samples)
    print(f"initial stress: rms {st0[0]:.6e} mean_abs {st0[1]:.6e} counted {st0[2]}", flush=True)
    og = O.Graph.from_dense(handles, path_first.copy(), node_len)
    entry = None
    for seed in [int(v) for v in a.seeds.split(",")]:
        op = O.params_from_graph(og, nthreads=a.threads)
        op.iter_max = a.iter_max
        op.seed = seed
        t1 = time.time()
        x, st, rc = O.path_linear_sgd(og, op, mode=O.MODE_EXACT if a.mode == "exact" else O.MODE_REFERENCE, x0=x0)
        assert rc == 0
        print(f"oracle {a.mode} seed {seed}: {st.applied} updates in {st.seconds:.0f}s ({st.applied/st.seconds/1e6:.1f} M/s), wall {time.time()-t1:.0f}s", flush=True)
        np.save(os.path.join(out_dir, f"x_{seed}.npy"), x)
        st1 = philox_stress(handles, path_first, node_len, x, a.samples)
        print(f"final stress: rms {st1[0]:.6e} mean_abs {st1[1]:.6e} counted {st1[2]}", flush=True)
        if entry is None:
            entry = {"nodes": a.nodes, "paths": a.paths, "steps": int(len(handles)), "graph_seed": a.graph_seed,
                     "stress_sample": {"samples": a.samples, "seed": 12345, "draw": "philox stream 2 (same sample as gfs_stress)"},
                     "initial": {"rms_rel": st0[0], "mean_abs_rel": st0[1], "counted": st0[2]},
                     "params": {k: v for k, v in op.as_dict().items() if k not in ("seed", "nthreads")},
                     "command": "python tools/oracle_runs.py " + " ".join(sys.argv[1:]), "runs": []}
        entry["runs"].append({"sgd_seed": seed, "threads": a.threads, "mode": a.mode, "applied": int(st.applied),
                              "sgd_seconds": st.seconds, "updates_per_s": st.applied / st.seconds,
                              "final": {"rms_rel": st1[0], "mean_abs_rel": st1[1], "counted": st1[2]}})
        fx = {}
        if os.path.exists(a.fixture):
            with open(a.fixture) as f:
                fx = json.load(f)
        fx.setdefault("_provenance", "Sampled path stress reached by the CPU oracle (oracle/gfs_oracle.cpp: the C++ restatement of reference "
                      "src/sgd.rs; exact-count epochs, all host cores) on synthetic graphs, evaluated with numpy on the SAME Philox sample "
                      "gfs_stress uses.  ORACLE-derived (the Rust reference cannot be built here); the runs are multi-threaded Hogwild, so "
                      "values reproduce statistically, not bit for bit.  Made by tools/oracle_runs.py.")
        fx[a.name] = entry
        with open(a.fixture, "w") as f:
            json.dump(fx, f, indent=1)
    print(json.dumps(entry), flush=True)


if __name__ == "__main__":
    main()
