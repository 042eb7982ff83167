#!/usr/bin/env python
"""bench.py — SGD term updates/sec of the path-guided SGD hot path on synthetic pangenome graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload y10m|l10m|y1m|y100m] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A STEP is one epoch of the schedule: `min_term_updates` applied term updates over the whole graph
(BASELINE.json config 3: 10M nodes / 90 paths / ~0.83e9 steps => 0.83e9 updates per step), taken
at K points spread evenly over the reference's 101-epoch eta schedule (so warm epochs, with uniform
partners, and cooling epochs, with theta = 0.001, are both in the timed region).  With N > 1 the
terms of every step are sharded over the ranks (disjoint slices of the step array), every rank
keeps a replica of the positions, and replicas are reconciled `syncs` times per step over NVLink peer
memory — total work is fixed, so scaling is "strong".  Every rank drives its GPU through the
C ABI's gfs_replica_* entry points; torch.distributed only carries region handles, the shared node order and the
timing / stress reductions.

Printed keys (one JSON line, rank 0): see the task contract; `value` is device-timed with the index
and positions resident in HBM; `e2e` is the same metric through the public host-buffer API — gfs_index_build32
from PAGEABLE host arrays + the full schedule + download (N = 1: exactly the two calls the Rust host makes) —
wall-clocked around synchronised calls; `roofline` uses 192 algorithmic bytes per update against
MEASURED_PEAKS.json, with the DRAM-side fraction beside it; `cpu_baseline` is the C++ restatement of the reference
(oracle/, kind "port" — the Rust reference cannot be built in this image) on all host cores on a bounded sample;
`also` carries the other BASELINE.json configs that fit the run (config 4 `L` 2D; config 5 at N = 8).

`--impl reference` times that same CPU restatement as the reference arm: same config, same epochs of the same
schedule, each step a bounded sample of the epoch's updates.  It loads nothing of the product library.
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
K1_ALGO_BYTES_PER_STEP = 20      # 8 B handle + 4 B gathered length + 8 B offset
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


def nvlink_bytes(gpu: int):
    """Cumulative NVLink data counters of one GPU (sum over links), bytes: (tx, rx) or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
        tx = rx = 0
        found = False
        for ln in out.splitlines():
            ln = ln.strip()
            if "Data Tx:" in ln:
                tx += int(ln.split("Data Tx:")[1].split()[0]); found = True
            elif "Data Rx:" in ln:
                rx += int(ln.split("Data Rx:")[1].split()[0]); found = True
        if not found:
            log("[bench] nvidia-smi nvlink -gt d: no data counters in the output: " + " | ".join(out.splitlines()[:4]))
        return (tx * 1024, rx * 1024) if found else None
    except Exception as e:
        log(f"[bench] nvidia-smi nvlink failed: {e}")
        return None


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


def workload_config(wl, total_steps, world, syncs, reconcile):
    """The `config` object: identical in both arms (it names the workload and the schedule, not the machine)."""
    nodes, paths, dims, desc = wl
    iter_max = 100 if dims == 0 else 30
    M = total_steps if dims == 0 else 10 * total_steps
    if not syncs:                       # the library's default (gfs_default_syncs_per_epoch): one reconcile per S applied updates
        syncs = min(max((M + total_steps // 2) // max(total_steps, 1), 1), 64)
    return {"workload": desc, "nodes": nodes, "paths": paths, "total_steps": int(total_steps), "dims": dims,
            "updates_per_step": int(M), "iter_max": iter_max,
            "step": "one epoch of the eta schedule (min_term_updates applied updates); step k of K runs epoch "
                    "floor(k*(iter_max+1)/K), so warm and cooling epochs are both timed",
            "l2": "inputs larger than L2: step records %.1f GB + positions %.0f MB vs 126 MB L2"
                  % (total_steps * 16 / 1e9, nodes * (8 if dims == 0 else 16) / 1e6),
            "parallelism": (f"replicas x{world}, terms sharded by step slice, peer-memory reconcile({reconcile}) "
                            f"x{syncs}/epoch behind the C ABI (gfs_replica_*)") if world > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
# CPU legs (the oracle: test infrastructure, used here only as the reported baseline / reference arm).
# Nothing of the product library is loaded: the graph comes from oracle/libgfs_synth.so (the same generator
# source compiled alone).
# --------------------------------------------------------------------------------------------------
def cpu_leg(nodes, paths, dims, steps, warmup, sample_updates, seed=42):
    """Times the C++ restatement of path_linear_sgd on all host cores.  Step k runs epoch
    floor(k*(iter_max+1)/K) of the reference's schedule — the epochs the GPU arm times — on the SAME graph,
    bounded to `sample_updates` applied updates (reference mode: checker thread, free-running workers)."""
    from oracle import oracle as O
    from oracle.synth_host import synth_arrays
    cores = os.cpu_count() or 1
    t0 = time.time()
    handles, first, node_len = synth_arrays(nodes, paths, seed)
    og = O.Graph.from_dense(handles, first.copy(), node_len)
    N, S = len(node_len), len(handles)
    del handles
    op = O.params_from_graph(og, layout=dims > 0, nthreads=cores)
    n_epochs = op.iter_max + 1
    full = op.min_term_updates
    op.min_term_updates = max(1, min(sample_updates, full))
    log(f"[cpu] graph N={N} S={S} built in {time.time()-t0:.1f}s; {cores} threads; {op.min_term_updates} of {full} updates per step")
    rates, times = [], []
    K = max(steps, 1)
    try:
        if dims == 0:
            t0 = time.time()
            pix = O.PrebuiltIndex(og)
            log(f"[cpu] oracle PathIndex + handle map in {time.time()-t0:.1f}s")
            x = O.init_x(og)
            for k in range(warmup + steps):
                e = epoch_of_step(k, max(warmup, 1), n_epochs) if k < warmup else epoch_of_step(k - warmup, K, n_epochs)
                O.set_epoch_window(e, e + 1)
                x, st, rc = O.path_linear_sgd(og, op, mode=O.MODE_REFERENCE, x0=x, index=pix)
                assert rc == 0
                if k >= warmup:
                    rates.append(st.applied / st.seconds); times.append(st.seconds)
            pix.close()
        else:
            c = O.init_layout(og, dims, op.seed)
            for k in range(warmup + steps):
                e = epoch_of_step(k, max(warmup, 1), n_epochs) if k < warmup else epoch_of_step(k - warmup, K, n_epochs)
                O.set_epoch_window(e, e + 1)
                c, st, rc = O.path_linear_sgd_layout(og, op, dims, mode=O.MODE_REFERENCE, coords0=c)
                assert rc == 0
                if k >= warmup:
                    rates.append(st.applied / st.seconds); times.append(st.seconds)
    finally:
        O.set_epoch_window()
    total_updates = sum(r * t for r, t in zip(rates, times))
    return {"value": total_updates / sum(times), "ms_per_step": 1e3 * sum(times) / len(times), "cores": cores, "total_steps": S,
            "sample": f"{len(times)} steps x {op.min_term_updates} applied updates (a bounded sample of the epoch's {full}; epochs "
                      f"floor(k*{n_epochs}/{K}) of the reference schedule, reference-mode checker thread) on the same "
                      f"{nodes}-node / {paths}-path graph"}


def run_reference(a, wl):
    nodes, paths, dims, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = a.cpu_sample or (60_000_000 if nodes >= 5_000_000 else 30_000_000)
    r = cpu_leg(nodes, paths, dims, a.steps, a.warmup, sample, a.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "sgd_term_updates_per_sec", "value": r["value"], "unit": "updates/s",
           "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": workload_config(wl, r["total_steps"], world, a.syncs, a.reconcile),
           "note": "C++ restatement of reference src/sgd.rs (oracle/gfs_oracle.cpp) on the host cores: the Rust reference "
                   "cannot be built in this image (no cargo/rustc)",
           "cpu_baseline": {"value": r["value"], "unit": "updates/s", "cores": r["cores"], "kind": "port",
                            "sample": r["sample"]},
           "e2e": {"value": r["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def measure_workload(a, wl, name, rank, world, local, dev, K, W, with_cpu, e2e_epochs):
    """Device-resident value + e2e + stress for one workload.  Returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist

    import gfasort_b200 as G
    from gfasort_b200 import multi
    from gfasort_b200.synth import synth_path_counts

    nodes, paths, dims, desc = wl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- graph: only this rank's paths are materialised; PAGEABLE host memory (what a Rust Vec is) --------
    t0 = time.time()
    counts = synth_path_counts(nodes, paths, a.seed)
    path_first = np.zeros(paths + 1, dtype=np.uint64)
    np.cumsum(counts, out=path_first[1:])
    S = int(path_first[-1])
    shard = multi.shard_steps(path_first, rank, world)
    sg = G.SynthGraph(nodes, paths, seed=a.seed, path_begin=shard.path_begin, path_end=shard.path_end, pinned=bool(a.pinned))
    # the host flattens Vec<Handle> once either way; dense idx < 2^31, so it flattens to 32-bit handles
    handles = sg.step_handles
    pin_keep = None
    if a.handles == 32:
        if a.pinned:
            from gfasort_b200.sgd import PinnedArray
            pin_keep = PinnedArray(sg.S, np.uint32)              # gfs_host_alloc: what the binding flattens into
            handles = pin_keep.array
            handles[:] = sg.step_handles
        else:
            handles = sg.step_handles.astype(np.uint32)
    shard_steps_held = sg.S
    if rank == 0:
        log(f"[bench] {desc}: N={nodes} P={paths} S={S}; rank 0 holds paths [{shard.path_begin},{shard.path_end}) "
            f"= {sg.S} steps; generated in {time.time()-t0:.1f}s; {a.handles}-bit handles, {'pinned' if a.pinned else 'pageable'} host memory")
    node_len = sg.node_len
    x0 = initial_positions(node_len, dims)
    sampler = {"window": os.environ.get("GFASORT_WINDOW", "auto"), "chunk": os.environ.get("GFASORT_CHUNK", "256"),
               "coherent": os.environ.get("GFASORT_COHERENT", "1"), "relabel": os.environ.get("GFASORT_RELABEL", "1")}
    syncs = a.syncs if world > 1 else 1          # 0 = the library's default (one reconcile per S applied updates: 1 for Y, 10 for L)

    phases = {}

    def build_run(iter_max=None):
        t0 = time.perf_counter()
        ix = multi.build_shard_index(handles, sg.path_first, node_len, device=local, rank=rank, world=world)
        t1 = time.perf_counter()
        max_bp = int(ix.path_lengths().max()) if sg.P else 0
        if world > 1:
            t = torch.tensor([max_bp], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            max_bp = int(t.item())
        params = derive_params(G, dims, counts, max_bp, iter_max)
        t2 = time.perf_counter()
        run = multi.ReplicaRun(ix, nodes, shard, S, params, dims=dims, device=local, syncs_per_epoch=syncs, mode=a.reconcile)
        t3 = time.perf_counter()
        phases.update(index=t1 - t0, params=t2 - t1, replica=t3 - t2)
        return ix, params, run

    def syncs_of(run):
        return run.syncs if world > 1 else 1

    # ---- device-resident measurement -------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ix, params, run = build_run()
    syncs = syncs_of(run)
    binfo = ix.build_info()
    run.upload(x0)
    n_epochs = params.iter_max + 1
    M = params.min_term_updates
    for k in range(W):
        run.run_epoch(epoch_of_step(k, max(W, 1), n_epochs))
    run.flush()
    barrier()
    st0 = run.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv0 = nvlink_bytes(local) if (world > 1 and rank == 0) else None     # a subprocess: BEFORE the barrier, or the peers would wait for it inside the timed region
    barrier()
    c_lo = clocks.mark()
    with torch.cuda.stream(run.stream):
        ev0.record()
    for k in range(K):
        run.run_epoch(epoch_of_step(k, K, n_epochs))
    run.flush()                          # the last overlapped reconcile belongs to the timed region
    with torch.cuda.stream(run.stream):
        ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    c_hi = clocks.mark()
    nv1 = nvlink_bytes(local) if (world > 1 and rank == 0) else None
    clk = clocks.stop(c_lo, c_hi) if rank == 0 else None
    st1 = run.stats()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    kern_s = torch.tensor([st1["kernel_seconds"] - st0["kernel_seconds"]], dtype=torch.float64, device=dev)
    applied = torch.tensor([st1["applied_updates"] - st0["applied_updates"]], dtype=torch.int64, device=dev)
    attempts = torch.tensor([st1["attempts"] - st0["attempts"]], dtype=torch.int64, device=dev)
    kern_ranks = None
    if world > 1:
        gathered = [torch.zeros_like(kern_s) for _ in range(world)]
        dist.all_gather(gathered, kern_s)                      # per-rank SGD kernel time: the skew the reconcile's start barrier waits for
        kern_ranks = [float(g.item()) for g in gathered]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(applied, op=dist.ReduceOp.SUM)
        dist.all_reduce(attempts, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    total_applied = int(applied.item())
    assert total_applied == K * M, f"applied {total_applied} != {K} x {M}"
    launches_per_rank = st1["launches"] - st0["launches"]
    sgd_launches_per_rank = K * syncs
    assert bool(torch.isfinite(run.x).all().item()), "positions are not finite"
    value = total_applied / (ms * 1e-3)
    # roofline of the dominant kernel (the SGD term kernel): algorithmic bytes per launch / mean launch time
    launch_s = float(kern_s.item()) / sgd_launches_per_rank
    upd_per_launch = total_applied / (sgd_launches_per_rank * world)
    achieved = upd_per_launch * ALGO_BYTES_PER_UPDATE / launch_s / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        pass
    traffic = traffic_src = dram_frac = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(name)
            if tj:
                traffic = tj["dram_bytes_per_update"] * upd_per_launch
                traffic_src = "static: " + tj.get("source", "ncu --set full capture under profiles/") + (
                    "; scaled to this launch's update count" if world > 1 else "")
                dram_frac = traffic / launch_s / 1e9 / peak
    except Exception:
        pass
    grid, block = st1["grid"], st1["block"]
    window_steps, coherent = st1.get("window_steps"), st1.get("coherent")
    nvlink = None
    if nv0 and nv1:
        n_elems = nodes if dims == 0 else nodes * 4
        esz = 8 if dims == 0 else 4
        nvlink = {"tx_bytes_per_reconcile": (nv1[0] - nv0[0]) / (K * syncs), "rx_bytes_per_reconcile": (nv1[1] - nv0[1]) / (K * syncs),
                  "algorithmic_bytes_per_reconcile_per_direction": n_elems * esz * (world - 1) / world,
                  "source": "nvidia-smi nvlink -gt d on rank 0's GPU, around the timed region"}
    run.close(); ix.close()

    # ---- end to end through the host-buffer API ----------------------------------------------------
    e2e = None
    if e2e_epochs != 0:
        e_iter = params.iter_max if e2e_epochs < 0 else max(2, e2e_epochs - 1)
        stats_e = None
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            # exactly the Rust host's two calls: gfs_index_build(32) + gfs_sgd_1d / gfs_sgd_nd, host buffers in and out
            ix = G.PathIndex.from_arrays(handles, sg.path_first, node_len, env=True)
            t1 = time.perf_counter()
            p2 = derive_params(G, dims, counts, int(ix.path_lengths().max()), e_iter)
            xf = x0.copy()
            t2 = time.perf_counter()
            import ctypes as C

            from gfasort_b200._cabi import Stats, check, f64p, lib
            stats_e = Stats()
            cp = p2.c()
            if dims == 0:
                check(lib().gfs_sgd_1d(ix.handle, C.byref(cp), xf.ctypes.data_as(f64p), C.byref(stats_e)))
            else:
                check(lib().gfs_sgd_nd(ix.handle, C.byref(cp), dims, xf.ctypes.data_as(f64p), C.byref(stats_e)))
            t3 = time.perf_counter()
            dt = t3 - t0                                   # params derivation + x0 copy stay inside: the host does them too
            phases_s = f"gfs_index_build{a.handles if a.handles == 32 else ''} {t1-t0:.3f}s, params + x0 copy {t2-t1:.3f}s, gfs_sgd_{'1d' if dims == 0 else 'nd'} {t3-t2:.3f}s"
            e2e_binfo = ix.build_info()
        else:
            ix, p2, run = build_run(e_iter)               # H2D: step handles, first_step, node lengths
            t1 = time.perf_counter()
            run.upload(x0)                                # H2D: initial positions
            t2 = time.perf_counter()
            for e in range(p2.iter_max + 1):
                run.run_epoch(e)
            xf = run.download()                           # D2H: final positions (synchronises)
            t3 = time.perf_counter()
            barrier()
            dt = time.perf_counter() - t0
            phases_s = (f"index build (K1 + shared node order) {phases['index']:.3f}s, params {phases['params']:.3f}s, replica create + connect "
                        f"{phases['replica']:.3f}s, upload {t2-t1:.3f}s, {p2.iter_max+1} epochs + download {t3-t2:.3f}s")
            e2e_binfo = ix.build_info()
        if rank == 0:
            log(f"[bench] e2e phases: {phases_s}")
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert np.all(np.isfinite(xf))
        h2d = handles.nbytes + (sg.P + 1) * 8 + nodes * 4 + x0.nbytes
        e2e_updates = (p2.iter_max + 1) * M
        stress = None
        if a.stress:
            if world == 1:
                stress = G.layout_stress(None, xf, max(dims, 1), 1_000_000, ix, layout_order=dims > 0)
            else:
                stress = multi.all_paths_stress(ix, shard, S, xf, dims, 1_000_000, dims > 0, device=local)
        e2e = {"value": e2e_updates / dt, "unit": "updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(xf.nbytes), "seconds": dt, "epochs": p2.iter_max + 1,
               "host_memory": "pinned" if a.pinned else "pageable", "handle_bits": a.handles,
               "index_build": e2e_binfo,
               "what": ("gfs_index_build32 + gfs_sgd_1d/gfs_sgd_nd: the reference-facing C-ABI calls, host buffers in and out, "
                        "wall clock; one e2e step = one complete Y/L call") if world == 1 else
                       ("per rank: gfs_index_build_shard32 + gfs_replica_* (upload, full schedule with reconciles, download), wall "
                        "clock, max over ranks; one e2e step = one complete Y/L call"),
               "stress_mean_abs_rel": stress[1] if stress else None, "stress_rms_rel": stress[0] if stress else None,
               "stress_over": "all paths (1M Philox samples, seed 12345)" if stress else None}
        if world > 1:
            run.close()
        ix.close()
    sg.close()

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if with_cpu:
        r = cpu_leg(nodes, paths, dims, 4, 1, a.cpu_sample or (150_000_000 if nodes >= 5_000_000 else 50_000_000), a.seed)
        cpu = {"value": r["value"], "unit": "updates/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank != 0:
        return None
    k1 = None
    if binfo["kernel_seconds"] > 0:
        k1_gbs = shard_steps_held * K1_ALGO_BYTES_PER_STEP / binfo["kernel_seconds"] / 1e9
        k1 = {"kernel": "k1_scan_write (path index, one pass)", "achieved": k1_gbs, "frac": k1_gbs / peak, "unit": "GB/s",
              "algorithmic_bytes_per_step": K1_ALGO_BYTES_PER_STEP, "kernel_ms": binfo["kernel_seconds"] * 1e3,
              "build_ms": binfo["build_seconds"] * 1e3, "note": "K1 runs under the host->device copy of the next chunk"}
    cfg = workload_config(wl, S, world, syncs, a.reconcile)
    return {"metric": "sgd_term_updates_per_sec", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if dims == 0 else "f32", "data": "synthetic",
            "config": cfg,
            "launch": {"sampler": sampler, "grid": grid, "block": block, "window_steps": window_steps, "coherent": coherent,
                       "reconcile": a.reconcile if world > 1 else None, "syncs_per_epoch": syncs,
                       "overlapped_reconcile": {"0": "off (stop-the-world exchange; the default)", "1": "when the exchange is short against a slice",
                                                "2": "always", "3": "overlapped arithmetic, joined at once"}.get(os.environ.get("GFASORT_OVERLAP", "0"))
                                               if (world > 1 and a.reconcile == "p2p") else None},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches_per_rank * world),
            "attempts_per_update": float(attempts.item()) / total_applied,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "dram_frac": dram_frac, "peak_source": peak_src,
                         "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                         "updates_per_launch": upd_per_launch, "launch_ms": launch_s * 1e3,
                         "launch_ms_by_rank": [round(k / sgd_launches_per_rank * 1e3, 4) for k in kern_ranks] if kern_ranks else None,
                         "step_ms_minus_kernel_ms": ms / K - launch_s * 1e3 * syncs, "k1": k1},
            "nvlink": nvlink, "cpu_baseline": cpu}


def run_ours(a, wl):
    import torch
    import torch.distributed as dist

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
    K, W = a.steps, max(a.warmup, 3)      # timing rule: at least 3 untimed warm-up steps, whatever was asked for
    out = measure_workload(a, wl, a.workload, rank, world, local, dev, K, W, with_cpu=(world == 1 and not a.no_cpu),
                           e2e_epochs=a.e2e_epochs)
    # ---- the other BASELINE.json configs that fit this run, as compact records ----------------------------
    also = []
    if a.also and a.workload == "y10m":
        extra = ["l10m"] + (["y100m"] if world == 8 else [])
        for name in extra:
            t0 = time.time()
            try:
                r = measure_workload(a, WORKLOADS[name], name, rank, world, local, dev, min(K, 5), 3, with_cpu=False,
                                     e2e_epochs=a.e2e_epochs)
            except Exception as e:          # an extra must never cost the headline line
                r = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            if rank == 0 and r is not None:
                if "error" not in r:
                    r = {"workload": name, "config": r["config"], "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                         "steps": r["steps"], "dtype": r["dtype"], "roofline": {k: r["roofline"][k] for k in ("achieved", "frac", "dram_frac", "launch_ms")},
                         "e2e": r["e2e"], "launch": r["launch"], "nvlink": r["nvlink"], "seconds_spent": time.time() - t0}
                else:
                    r["workload"] = name
                also.append(r)
    if rank == 0:
        out["also"] = also
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
    ap.add_argument("--syncs", type=int, default=int(os.environ.get("GFASORT_SYNCS", "0")),
                    help="replica reconciles per epoch (N > 1); 0 = the library's default: one per S applied updates (1 for Y, 10 for L)")
    ap.add_argument("--reconcile", default=os.environ.get("GFASORT_RECONCILE", "p2p"), choices=["p2p", "tavg", "avg", "delta"],
                    help="p2p: one peer-memory kernel per rank behind the C ABI (default); tavg/avg/delta: NCCL all-reduce driven from torch")
    ap.add_argument("--e2e-epochs", type=int, default=-1, help="-1 = the full schedule, 0 = skip the e2e leg")
    ap.add_argument("--stress", type=int, default=1, help="report the sampled path stress of the e2e result (all paths)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="applied updates per CPU step (0 = auto)")
    ap.add_argument("--handles", type=int, default=32, choices=[32, 64], help="width of the flattened step handles the host passes")
    ap.add_argument("--pinned", type=int, default=0, help="1 = the host step array is page-locked (default: pageable, like a Rust Vec)")
    ap.add_argument("--also", type=int, default=int(os.environ.get("GFASORT_BENCH_ALSO", "1")),
                    help="append compact records of the other configs (l10m; y100m at 8 GPUs) to the y10m line")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)
    wl = WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, wl)
    else:
        run_ours(a, wl)


if __name__ == "__main__":
    main()
