#!/usr/bin/env python
"""bench.py — SGD term updates/sec of the path-guided SGD hot path on synthetic pangenome graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload y10m|l10m|y1m|y100m] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A STEP is one epoch of the schedule: `min_term_updates` applied term updates over the whole graph
(BASELINE.json config 3: 10M nodes / 90 paths / ~0.83e9 steps => 0.83e9 updates per step), taken
at K points spread evenly over the reference's 101-epoch eta schedule (so warm epochs, with uniform
partners, and cooling epochs, with theta = 0.001, are both in the timed region).  With N > 1 the
terms of every step are sharded over the ranks (disjoint slices of the step array), every rank
keeps a replica of the positions, and replicas are averaged by an NCCL all-reduce `syncs` times per
step — total work is fixed, so scaling is "strong".

Printed keys (one JSON line, rank 0): see the task contract; `value` is device-timed with the index
and positions resident in HBM; `e2e` is the same metric through the public host-buffer API
(index build from pinned host arrays + the full schedule + download), wall-clocked around
synchronised calls; `roofline` uses 192 algorithmic bytes per update against
MEASURED_PEAKS.json; `cpu_baseline` is the C++ restatement of the reference (oracle/, kind "port" —
the Rust reference cannot be built in this image) on all host cores on a bounded sample.

`--impl reference` times that same CPU restatement as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_UPDATE = 192      # 6 sector transactions x 32 B
WORKLOADS = {
    #  name: (nodes, paths, dims, description)
    "y10m": (10_000_000, 90, 0, "config3: synthetic 10M-node / 90-path graph (~0.83e9 steps), 1D Y, iter_max 100"),
    "l10m": (10_000_000, 90, 2, "config4: same 10M-node graph, 2D L (float2 per node end), layout-iter 30"),
    "y1m": (1_000_000, 32, 0, "config2: synthetic 1M-node / 32-path graph, 1D Y, iter_max 100"),
    "y100m": (100_000_000, 90, 0, "config5: synthetic 100M-node / 90-path graph, 1D Y, replicated positions"),
    "ytiny": (50_000, 8, 0, "tiny 50k-node / 8-path graph (plumbing check only)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md "clocks line")
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self) -> int:
        return len(self.lines)

    def stop(self, lo: int = 0, hi: int | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[lo:hi]
        scope = "timed region"
        if not window:          # region shorter than one sampling period: use the samples around it
            window, scope = self.lines[max(0, lo - 3):(hi or 0) + 3], "around the timed region"
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def derive_params(G, dims, counts, max_path_bp, iter_max=None):
    """YgsParams::from_graph (ygs.rs:50-92) / LayoutSGDParams::from_graph (sgd.rs:733-763) from the
    per-path step counts and the longest path in bp."""
    total, mx = int(counts.sum()), int(counts.max())
    if dims == 0:
        return G.PathSGDParams(iter_max=iter_max or 100, min_term_updates=total, eta_max=float(mx * mx),
                               space=int(max_path_bp), space_max=100, space_quantization_step=100)
    return G.LayoutSGDParams(dimensions=dims, iter_max=iter_max or 30, min_term_updates=10 * total,
                             eta_max=float(mx * mx), space=mx, space_max=1000, space_quantization_step=100)


def initial_positions(node_len, dims, seed=9399220):
    n = len(node_len)
    cum = np.zeros(n, dtype=np.float64)
    np.cumsum(node_len[:-1], dtype=np.float64, out=cum[1:])
    if dims == 0:
        return cum
    c = np.zeros((n, 2, dims), dtype=np.float64)
    c[:, 0, 0] = cum
    c[:, 1, 0] = cum + node_len
    if dims > 1:
        rng = np.random.Generator(np.random.PCG64(seed))
        c[:, :, 1:] = rng.standard_normal((n, 2, dims - 1)) * np.sqrt(2.0 * n)
    return c.reshape(-1)


def epoch_of_step(k, K, n_epochs):
    return (k * n_epochs) // K if K <= n_epochs else k % n_epochs


# --------------------------------------------------------------------------------------------------
# CPU legs (the oracle: test infrastructure, used here only as the reported baseline / reference arm)
# --------------------------------------------------------------------------------------------------
def cpu_leg(nodes, paths, dims, steps, warmup, sample_updates, seed=42):
    """Times the C++ restatement of path_linear_sgd on all host cores.  Each step applies
    `sample_updates` updates (3 short epochs: 2 warm + 1 cooling) on the SAME graph as the GPU arm."""
    import gfasort_b200 as G
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    t0 = time.time()
    s = G.SynthGraph(nodes, paths, seed=seed)
    og = O.Graph.from_dense(s.step_handles, s.path_first.copy(), s.node_len)
    N, S = s.N, s.S
    s.close()
    op = O.params_from_graph(og, layout=dims > 0, nthreads=cores)
    op.iter_max = 2
    op.min_term_updates = max(1, sample_updates // 3)
    per_step = 3 * op.min_term_updates
    log(f"[cpu] graph N={N} S={S} built in {time.time()-t0:.1f}s; {cores} threads; {per_step} updates per step")
    rates, times = [], []
    if dims == 0:
        t0 = time.time()
        pix = O.PrebuiltIndex(og)
        log(f"[cpu] oracle PathIndex + handle map in {time.time()-t0:.1f}s")
        x0 = O.init_x(og)
        for k in range(warmup + steps):
            x, st, rc = O.path_linear_sgd(og, op, mode=O.MODE_REFERENCE, x0=x0, index=pix)
            assert rc == 0
            if k >= warmup:
                rates.append(st.applied / st.seconds); times.append(st.seconds)
        pix.close()
    else:
        c0 = O.init_layout(og, dims, op.seed)
        for k in range(warmup + steps):
            c, st, rc = O.path_linear_sgd_layout(og, op, dims, mode=O.MODE_REFERENCE, coords0=c0)
            assert rc == 0
            if k >= warmup:
                rates.append(st.applied / st.seconds); times.append(st.seconds)
    total_updates = sum(r * t for r, t in zip(rates, times))
    return {"value": total_updates / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "cores": cores,
            "sample": f"{len(times)} x {per_step} applied updates (3 epochs: 2 warm + 1 cooling, reference-mode "
                      f"checker thread) on the same {nodes}-node / {paths}-path graph"}


def run_reference(a, wl):
    nodes, paths, dims, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = a.cpu_sample or (60_000_000 if nodes >= 5_000_000 else 30_000_000)
    r = cpu_leg(nodes, paths, dims, a.steps, a.warmup, sample)
    out = {"impl": "reference", "metric": "sgd_term_updates_per_sec", "value": r["value"], "unit": "updates/s",
           "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": desc, "nodes": nodes, "paths": paths, "dims": dims,
                      "note": "C++ restatement of reference src/sgd.rs (oracle/gfs_oracle.cpp): the Rust reference "
                              "cannot be built in this image (no cargo/rustc)"},
           "cpu_baseline": {"value": r["value"], "unit": "updates/s", "cores": r["cores"], "kind": "port",
                            "sample": r["sample"]},
           "e2e": {"value": r["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(a, wl):
    import torch
    import torch.distributed as dist

    import gfasort_b200 as G
    from gfasort_b200 import multi
    from gfasort_b200.synth import synth_path_counts

    nodes, paths, dims, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — gfasort_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if world != a.gpus and rank == 0:
        log(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={world}; using {world}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- graph: only this rank's paths are materialised -----------------------------------------
    t0 = time.time()
    counts = synth_path_counts(nodes, paths, a.seed)
    path_first = np.zeros(paths + 1, dtype=np.uint64)
    np.cumsum(counts, out=path_first[1:])
    S = int(path_first[-1])
    shard = multi.shard_steps(path_first, rank, world)
    sg = G.SynthGraph(nodes, paths, seed=a.seed, path_begin=shard.path_begin, path_end=shard.path_end, pinned=True)
    if rank == 0:
        log(f"[bench] {desc}: N={nodes} P={paths} S={S}; rank 0 holds paths [{shard.path_begin},{shard.path_end}) "
            f"= {sg.S} steps; generated in {time.time()-t0:.1f}s")
    node_len = sg.node_len
    x0 = initial_positions(node_len, dims)
    sampler = {"window": os.environ.get("GFASORT_WINDOW", "auto"), "chunk": os.environ.get("GFASORT_CHUNK", "256"), "coherent": os.environ.get("GFASORT_COHERENT", "1"),
               "relabel": os.environ.get("GFASORT_RELABEL", "1")}

    def build_run(iter_max=None):
        ix = multi.build_shard_index(sg.step_handles, sg.path_first, node_len, device=local, rank=rank, world=world)
        max_bp = int(ix.path_lengths().max()) if sg.P else 0
        if world > 1:
            t = torch.tensor([max_bp], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            max_bp = int(t.item())
        params = derive_params(G, dims, counts, max_bp, iter_max)
        run = multi.ReplicaRun(ix, nodes, shard, S, params, dims=dims, device=local,
                               syncs_per_epoch=a.syncs if world > 1 else 1, mode=a.reconcile)
        return ix, params, run

    # ---- device-resident measurement -------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ix, params, run = build_run()
    run.upload(x0)
    n_epochs = params.iter_max + 1
    M = params.min_term_updates
    K, W = a.steps, max(a.warmup, 3)      # timing rule: at least 3 untimed warm-up steps, whatever was asked for
    for k in range(W):
        run.run_epoch(epoch_of_step(k, max(W, 1), n_epochs))
    barrier()
    st0 = run.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c_lo = clocks.mark()
    with torch.cuda.stream(run.stream):
        ev0.record()
    for k in range(K):
        run.run_epoch(epoch_of_step(k, K, n_epochs))
    with torch.cuda.stream(run.stream):
        ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    c_hi = clocks.mark()
    clk = clocks.stop(c_lo, c_hi) if rank == 0 else None
    st1 = run.stats()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    kern_s = torch.tensor([st1["kernel_seconds"] - st0["kernel_seconds"]], dtype=torch.float64, device=dev)
    applied = torch.tensor([st1["applied_updates"] - st0["applied_updates"]], dtype=torch.int64, device=dev)
    attempts = torch.tensor([st1["attempts"] - st0["attempts"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(applied, op=dist.ReduceOp.SUM)
        dist.all_reduce(attempts, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    total_applied = int(applied.item())
    assert total_applied == K * M, f"applied {total_applied} != {K} x {M}"
    launches_per_rank = st1["launches"] - st0["launches"]
    assert bool(torch.isfinite(run.x).all().item()), "positions are not finite"
    value = total_applied / (ms * 1e-3)
    # roofline of the dominant kernel (the SGD term kernel): algorithmic bytes per launch / mean launch time
    launch_s = float(kern_s.item()) / launches_per_rank
    upd_per_launch = total_applied / (launches_per_rank * world)
    achieved = upd_per_launch * ALGO_BYTES_PER_UPDATE / launch_s / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(a.workload)
            if tj:
                traffic = tj["dram_bytes_per_update"] * upd_per_launch
    except Exception:
        pass
    grid, block = st1["grid"], st1["block"]
    run.close(); ix.close()

    # ---- end to end through the host-buffer API ----------------------------------------------------
    e2e = None
    if a.e2e_epochs != 0:
        e_iter = params.iter_max if a.e2e_epochs < 0 else max(2, a.e2e_epochs - 1)
        barrier()
        t0 = time.perf_counter()
        ix, p2, run = build_run(e_iter)               # H2D: step handles, first_step, node lengths (pinned host)
        t1 = time.perf_counter()
        run.upload(x0)                                # H2D: initial positions
        t2 = time.perf_counter()
        for e in range(p2.iter_max + 1):
            run.run_epoch(e)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        xf = run.download()                           # D2H: final positions
        barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            log(f"[bench] e2e phases: index build + session {t1-t0:.3f}s, upload {t2-t1:.3f}s, "
                f"{p2.iter_max+1} epochs {t3-t2:.3f}s, download {time.perf_counter()-t3:.3f}s")
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert np.all(np.isfinite(xf))
        h2d = sg.S * 8 + (sg.P + 1) * 8 + nodes * 4 + x0.nbytes
        e2e_updates = (p2.iter_max + 1) * M
        stress = None
        if a.stress and rank == 0:      # N > 1: sampled over the paths rank 0 holds
            stress = G.layout_stress(None, xf, max(dims, 1), 1_000_000, ix, layout_order=dims > 0)
        stress_k = None
        if a.stress and rank == 0 and world == 1 and a.stress_paths:
            k = min(a.stress_paths, paths)
            six = G.PathIndex.from_arrays(sg.step_handles, sg.path_first, node_len, path_begin=0, path_end=k, device=local, relabel=0)
            stress_k = G.layout_stress(None, xf, max(dims, 1), 1_000_000, six, layout_order=dims > 0)
            six.close()
            log(f"[bench] stress over paths [0,{k}): mean_abs {stress_k[1]:.4e} rms {stress_k[0]:.4e}")
        e2e = {"value": e2e_updates / dt, "unit": "updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(xf.nbytes), "seconds": dt, "epochs": p2.iter_max + 1,
               "what": "gfs_index_build from pinned host arrays + upload + full schedule + download, wall clock "
                       "around synchronised calls; one e2e step = one complete Y/L call",
               "stress_mean_abs_rel": stress[1] if stress else None, "stress_rms_rel": stress[0] if stress else None,
               "stress_over": ("all paths" if world == 1 else f"paths [{shard.path_begin},{shard.path_end}) of rank 0") if stress else None}
        run.close(); ix.close()
    sg.close()

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if world == 1 and not a.no_cpu:
        r = cpu_leg(nodes, paths, dims, 2, 1, a.cpu_sample or (300_000_000 if nodes >= 5_000_000 else 100_000_000), a.seed)
        cpu = {"value": r["value"], "unit": "updates/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        out = {"metric": "sgd_term_updates_per_sec", "value": value, "unit": "updates/s", "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64" if dims == 0 else "f32", "data": "synthetic",
               "config": {"workload": desc, "nodes": nodes, "paths": paths, "total_steps": S, "dims": dims,
                          "updates_per_step": M, "step": "one epoch of the eta schedule (min_term_updates applied updates)",
                          "l2": "inputs larger than L2: step records %.1f GB + positions %.0f MB vs 126 MB L2"
                                % (S * 16 / 1e9, nodes * (8 if dims == 0 else 16) / 1e6),
                          "sampler": sampler, "grid": grid, "block": block,
                          "parallelism": f"replicas x{world}, terms sharded by step slice, all-reduce({a.reconcile}) "
                                         f"x{a.syncs}/epoch" if world > 1 else "single GPU"},
               "clocks": clk, "e2e": e2e, "gpu_launches": int(launches_per_rank * world),
               "attempts_per_update": float(attempts.item()) / total_applied,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                            "updates_per_launch": upd_per_launch, "launch_ms": launch_s * 1e3},
               "cpu_baseline": cpu}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GFASORT_BENCH_WORKLOAD", "y10m"), choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--syncs", type=int, default=int(os.environ.get("GFASORT_SYNCS", "1")), help="replica reconciles per epoch (N > 1)")
    ap.add_argument("--reconcile", default=os.environ.get("GFASORT_RECONCILE", "tavg"), choices=["avg", "tavg", "delta", "p2p"])
    ap.add_argument("--stress-paths", type=int, default=0, help="N = 1: also report the stress over the first K paths only (to compare with rank 0 of a multi-GPU run)")
    ap.add_argument("--e2e-epochs", type=int, default=-1, help="-1 = the full schedule, 0 = skip the e2e leg")
    ap.add_argument("--stress", type=int, default=1, help="report the sampled path stress of the e2e result")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="applied updates per CPU step (0 = auto)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)
    wl = WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, wl)
    else:
        run_ours(a, wl)


if __name__ == "__main__":
    main()
