"""GPU tests (-m gpu) of the peer-memory reconcile kernel (gfasort_b200/csrc/gfs_p2p.cu, K5b) and of the replicated
multi-GPU run behind the C ABI (gfs_multi.cu).

One device: G replicas live on the same GPU, are connected with gfs_p2p_region_connect_local and reconciled by
gfs_p2p_reconcile_local — ONE cooperative launch in which block group g plays rank g, so the ranks meet at the same
in-kernel barriers, use the same partition and the same arithmetic as G ranks over NVLink, without ever having two
kernels of one GPU wait on one another.  Two or more devices (skipped otherwise): the real thing — one process
driving G GPUs through GFASORT_GPUS, and tools/p2p_ipc_check.py covers one process per GPU."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected(xs, xr, rel_thr=0.0):
    """x_sync + sum of displacements / #replicas that moved the element; exact when at most one moved it.
    rel_thr: the overlapped form calls an element moved when |x - x_sync| > rel_thr * |x_sync| (2.5 ulps)."""
    d = np.stack([x.astype(np.float64) - xs.astype(np.float64) for x in xr])
    if rel_thr:
        moved = np.stack([np.abs(x.astype(np.float64) - xs.astype(np.float64)) > np.abs(xs.astype(np.float64)) * rel_thr for x in xr])
    else:
        moved = np.stack([x != xs for x in xr])
    cnt = moved.sum(0)
    mean = xs.astype(np.float64) + (d * moved).sum(0) / np.maximum(cnt, 1)
    out = mean.astype(xs.dtype)
    one = cnt == 1
    which = moved.argmax(0)
    out[one] = np.stack(xr)[which[one], np.nonzero(one)[0]]
    out[cnt == 0] = xs[cnt == 0]
    return out


@pytest.mark.skipif("_n_gpus() < 2")
def test_one_call_multi_gpu_many_slices_per_epoch(gfs, monkeypatch):
    """GFASORT_SYNCS above the session's ring of 16 timing events: the one host thread must hand every device its slice
    before it waits for any of them (it once enqueued a whole epoch per device, blocked on device 0's ring while device 1
    had no work yet, and the reconcile barrier's bounded spin failed the run)."""
    s, counts, x0 = _synth(gfs, 50_000, 8)
    graph = gfs.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
    ix1 = gfs.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
    monkeypatch.setenv("GFASORT_GPUS", "2")
    monkeypatch.setenv("GFASORT_SYNCS", "40")
    ix2 = gfs.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len, env=True)
    p = gfs.PathSGDParams(iter_max=30, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                          space=int(ix1.path_lengths().max()), space_max=100)
    xb = gfs.path_linear_sgd_array(graph, p, ix2)
    st = dict(gfs.sgd.last_stats)
    assert st["applied_updates"] == (p.iter_max + 1) * p.min_term_updates and st["syncs_per_epoch"] == 40
    monkeypatch.delenv("GFASORT_SYNCS")
    xa = gfs.path_linear_sgd_array(graph, p, ix1)
    sa, sb = gfs.sort_stress(graph, xa, 200_000, ix1)[1], gfs.sort_stress(graph, xb, 200_000, ix1)[1]
    print(f"1D stress one GPU {sa:.5e} vs GFASORT_GPUS=2 GFASORT_SYNCS=40 {sb:.5e}")
    assert sb <= sa * 1.05
    ix1.close(); ix2.close()


def _own_slice(n, elem_bytes, G, g):
    """Element range of rank g's slice: the 16-byte vectors [nvec g/G, nvec (g+1)/G) of gfs_p2p.cu's slice_begin."""
    per = 16 // elem_bytes
    nvec = (n * elem_bytes + 15) // 16
    q, rem = divmod(nvec, G)
    lo = q * g + min(g, rem)
    hi = q * (g + 1) + min(g + 1, rem)
    return slice(min(lo * per, n), min(hi * per, n))


@pytest.mark.parametrize("dtype,G,n", [("float64", 2, 100_003), ("float64", 4, 1_000_000), ("float32", 3, 65_537),
                                       ("float32", 2, 1_000_001), ("float64", 1, 999), ("float64", 8, 123_457)])
def test_p2p_reconcile_matches_moved_replica_mean(dtype, G, n, gfs, monkeypatch):
    import torch
    from gfasort_b200.multi import PeerRegion
    monkeypatch.setenv("GFASORT_P2P_SPIN_CAP", str(1 << 21))          # ~1 s: a missing peer is an error, not a hang
    f64 = dtype == "float64"
    regions = [PeerRegion(0, n, f64, max_blocks=8) for _ in range(G)]
    try:
        PeerRegion.connect_local(regions)
        stream = torch.cuda.Stream(device=0)
        rng = np.random.default_rng(5)
        xs = (rng.standard_normal(n) * 1e6).astype(dtype)
        for r in regions:
            r.x_sync.copy_(torch.from_numpy(xs))
            r.x.copy_(torch.from_numpy(xs))
        for rnd in range(3):                                           # consecutive reconciles reuse the flags (tags)
            xr = []
            for g, r in enumerate(regions):
                mask = rng.random(n) < (0.6 if rnd < 2 else 0.05)
                x = xs.copy()
                x[mask] += (rng.standard_normal(int(mask.sum())) * 100).astype(dtype)
                xr.append(x)
                r.x.copy_(torch.from_numpy(x))
            torch.cuda.synchronize()
            PeerRegion.reconcile_local(regions, stream.cuda_stream)    # all ranks, one cooperative launch
            torch.cuda.synchronize()
            for r in regions:
                r.check()
            want = _expected(xs, xr)
            got = [r.x.cpu().numpy() for r in regions]
            for g in range(G):
                assert np.array_equal(got[g], got[0]), "replicas differ after the reconcile"
                own = _own_slice(n, 8 if f64 else 4, G, g)          # the base of a slice is kept by its owner only
                assert np.array_equal(regions[g].x_sync.cpu().numpy()[own], got[g][own]), "the owner's x_sync is not the new base"
            tol = 1e-9 if f64 else 1e-1
            assert np.allclose(got[0], want, rtol=0, atol=tol)
            one = np.stack([x != xs for x in xr]).sum(0) <= 1
            assert np.array_equal(got[0][one], want[one])              # exact where at most one replica moved it
            xs = got[0]
    finally:
        for r in regions:
            r.close()


def test_p2p_missing_peer_fails_the_run_on_every_rank(gfs, monkeypatch):
    """A rank that never shows up: the barrier's bounded spin gives up, the error word of EVERY rank is raised
    (the run failed; nobody may use its replica), and later reconciles return at once instead of timing out again."""
    import time

    import torch
    from gfasort_b200.multi import PeerRegion
    monkeypatch.setenv("GFASORT_P2P_SPIN_CAP", str(1 << 12))          # a few ms
    regions = [PeerRegion(0, 1000, True, max_blocks=2) for _ in range(2)]
    try:
        PeerRegion.connect_local(regions)
        st = torch.cuda.current_stream().cuda_stream
        regions[0].reconcile(st)                                       # rank 1 never launches: a lone kernel with a bounded spin
        torch.cuda.synchronize()
        for r in regions:
            with pytest.raises(gfs.GfsError, match="barrier timed out"):
                r.check()
        t0 = time.perf_counter()
        regions[0].reconcile(st)                                       # sticky: returns without waiting
        torch.cuda.synchronize()
        assert time.perf_counter() - t0 < 0.5
    finally:
        for r in regions:
            r.close()


def test_p2p_region_argument_checks(gfs):
    import ctypes as C

    from gfasort_b200._cabi import GFS_P2P_HANDLE_BYTES, lib, u8p
    from gfasort_b200.multi import PeerRegion
    a = PeerRegion(0, 10, True, max_blocks=2)
    b = PeerRegion(0, 11, True, max_blocks=2)
    c = PeerRegion(0, 10, True, max_blocks=3)
    try:
        with pytest.raises(gfs.GfsError):
            a.reconcile(0)                                             # not connected
        with pytest.raises(gfs.GfsError):
            PeerRegion.connect_local([a, b])                           # sizes differ
        with pytest.raises(gfs.GfsError):
            PeerRegion.connect_local([a, c])                           # grids differ
        # the IPC path checks the same agreement from the exchanged blobs, before opening any handle
        ha, hb, hc = a.ipc_handle(), b.ipc_handle(), c.ipc_handle()
        assert len(ha) == GFS_P2P_HANDLE_BYTES
        for other in (hb, hc):
            blob = (C.c_uint8 * (2 * GFS_P2P_HANDLE_BYTES)).from_buffer_copy(ha + other)
            assert lib().gfs_p2p_region_connect_ipc(a._h, C.cast(blob, u8p), 2, 0) != 0
            assert b"differs in size, element type or grid" in lib().gfs_last_error()
    finally:
        a.close(); b.close(); c.close()


# ------------------------------------------------------------------------------------------------
# the replicated run behind the C ABI
# ------------------------------------------------------------------------------------------------
def _synth(gfs, nodes, paths):
    s = gfs.SynthGraph(nodes, paths, seed=42)
    counts = np.diff(s.path_first)
    x0 = s.initial_positions()
    return s, counts, x0


def test_replica_world_one_equals_plain_run(gfs):
    """A gfs_replica with world = 1 is the plain session: same applied count, finite positions, same stress."""
    from gfasort_b200 import multi
    s, counts, x0 = _synth(gfs, 50_000, 8)
    ix = gfs.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
    p = gfs.PathSGDParams(iter_max=30, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                          space=int(ix.path_lengths().max()), space_max=100)
    shard = multi.shard_steps(s.path_first, 0, 1)
    run = multi.ReplicaRun(ix, s.N, shard, s.S, p, dims=0, device=0, mode="p2p")
    run.upload(x0)
    for e in range(p.iter_max + 1):
        run.run_epoch(e)
    x = run.download()
    st = run.stats()
    run.close()
    assert st["applied_updates"] == (p.iter_max + 1) * p.min_term_updates
    assert np.all(np.isfinite(x))
    graph = gfs.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
    x1 = gfs.path_linear_sgd_array(graph, p, ix)
    a, b = gfs.sort_stress(graph, x, 200_000, ix)[1], gfs.sort_stress(graph, x1, 200_000, ix)[1]
    assert abs(a - b) <= 0.05 * b
    ix.close()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2")
@pytest.mark.parametrize("dims", [0, 2])
def test_one_call_multi_gpu_through_the_cabi(dims, gfs, monkeypatch):
    """GFASORT_GPUS=2: gfs_index_build builds one shard per device, gfs_sgd_1d / gfs_sgd_nd run the replicated
    schedule from this one process, gfs_stress covers all paths.  Index bit-identical to one GPU; 1D stress within 2 % of
    the one-GPU run.  2D: the mean of two replicas of a 2D layout is slightly contracted wherever their local orientation
    differs, which a 31-epoch run on a graph this small does not fully repair — measured (9 seeds, 2 B200s) 4.93e-4
    [4.2e-4, 6.4e-4] against 3.91e-4 [3.4e-4, 4.7e-4] on one GPU and 4.30e-4 [3.7e-4, 5.7e-4] for the CPU oracle (5 seeds,
    tools/oracle_layout_spread.py); at config 4's size the difference is below the run-to-run spread.
    About one two-GPU run in nine ends 2-3 x worse (a fold that the 31 epochs do not repair; profiles/r2_experiments.md §6).
    Medians of 9 seeds, two GPUs over one, in four runs of this test / tools/one_call_multi.py: 1.26, 1.37, 1.40, 1.56;
    the bound (2 x) only guards against a broken exchange, it is not a parity claim."""
    s, counts, x0 = _synth(gfs, 200_000, 16)
    graph = gfs.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
    ix1 = gfs.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len)
    monkeypatch.setenv("GFASORT_GPUS", "2")
    ix2 = gfs.PathIndex.from_arrays(s.step_handles, s.path_first, s.node_len, env=True)
    assert ix2.build_info()["devices"] == 2
    assert np.array_equal(ix1.step_positions(), ix2.step_positions())
    assert np.array_equal(ix1.path_lengths(), ix2.path_lengths())
    if dims == 0:
        p = gfs.PathSGDParams(iter_max=100, min_term_updates=int(counts.sum()), eta_max=float(int(counts.max()) ** 2),
                              space=int(ix1.path_lengths().max()), space_max=100)
        xa = gfs.path_linear_sgd_array(graph, p, ix1)
        xb = gfs.path_linear_sgd_array(graph, p, ix2)
        st = dict(gfs.sgd.last_stats)
        assert st["applied_updates"] == (p.iter_max + 1) * p.min_term_updates
        sa, sb = gfs.sort_stress(graph, xa, 500_000, ix1), gfs.sort_stress(graph, xb, 500_000, ix2)
        sb1 = gfs.sort_stress(graph, xb, 500_000, ix1)
        assert sb[2] == sb1[2] and abs(sb[1] - sb1[1]) <= 1e-9 * sb1[1]       # sharded stress == one-GPU stress, same sample
        print(f"1D stress one GPU {sa[1]:.5e} vs GFASORT_GPUS=2 {sb[1]:.5e}")
        assert sb[1] <= sa[1] * 1.02          # measured on 2 B200s: 2.724e-4 vs 2.822e-4 (the replicated run ends slightly lower)
    else:
        from dataclasses import replace
        p = gfs.LayoutSGDParams(dimensions=2, iter_max=30, min_term_updates=10 * int(counts.sum()),
                                eta_max=float(int(counts.max()) ** 2), space=int(counts.max()), space_max=1000)
        c0 = gfs.initial_layout(graph, 2, p.seed)
        one, two = [], []
        for k in range(9):              # a 31-epoch 2D layout of this size moves +-30 % from run to run: medians of 9 seeds
            q = replace(p, seed=p.seed + 1000 * k)
            one.append(gfs.layout_stress(graph, gfs.path_linear_sgd_layout(graph, q, ix1, coords0=c0).coords, 2, 500_000, ix1)[1])
            c2 = gfs.path_linear_sgd_layout(graph, q, ix2, coords0=c0).coords
            two.append(gfs.layout_stress(graph, c2, 2, 500_000, ix2)[1])
            if k == 0:                  # sharded 2D stress == one-GPU stress of the same layout on the same sample
                s2, s1 = gfs.layout_stress(graph, c2, 2, 500_000, ix2), gfs.layout_stress(graph, c2, 2, 500_000, ix1)
                assert s2[2] == s1[2] and abs(s2[1] - s1[1]) <= 1e-9 * s1[1] and abs(s2[0] - s1[0]) <= 1e-9 * s1[0]
        a1, a2 = float(np.median(one)), float(np.median(two))
        print(f"2D stress, medians of 9: one GPU {a1:.5e} {one} vs GFASORT_GPUS=2 {a2:.5e} {two}")
        assert a2 <= a1 * 2.0
    ix1.close(); ix2.close()


@pytest.mark.parametrize("dtype,G,n", [("float64", 2, 100_003), ("float64", 4, 300_001), ("float32", 3, 65_537), ("float64", 8, 200_003),
                                       ("float32", 8, 70_001), ("float32", 4, 50_001), ("float32", 2, 33_333), ("float64", 5, 40_001)])
def test_p2p_overlapped_reconcile_arithmetic(dtype, G, n, gfs, monkeypatch):
    """The overlapped form (rc_p2p_async) on one device, all ranks in one cooperative launch: the exchange works on the
    SNAPSHOTS, every live replica receives (new base - its own snapshot) on top of whatever it did since the snapshot,
    and the owner of each slice stores the new base into its x_sync."""
    import torch
    from gfasort_b200.multi import PeerRegion
    monkeypatch.setenv("GFASORT_P2P_SPIN_CAP", str(1 << 21))
    f64 = dtype == "float64"
    regions = [PeerRegion(0, n, f64, max_blocks=8) for _ in range(G)]
    try:
        PeerRegion.connect_local(regions)
        stream = torch.cuda.Stream(device=0)
        rng = np.random.default_rng(9)
        base = (rng.standard_normal(n) * 1e6).astype(dtype)
        for rnd in range(2):
            snaps, lives = [], []
            for g, r in enumerate(regions):
                mask = rng.random(n) < (0.6 if rnd == 0 else 0.05)
                snap = base.copy()
                snap[mask] += (rng.standard_normal(int(mask.sum())) * 100).astype(dtype)
                since = np.zeros(n, dtype=dtype)                     # what the rank did after taking its snapshot
                m2 = rng.random(n) < 0.3
                since[m2] = (rng.standard_normal(int(m2.sum())) * 10).astype(dtype)
                live = (snap + since).astype(dtype)
                snaps.append(snap); lives.append(live)
                r.x_sync.copy_(torch.from_numpy(base)); r.x_snap.copy_(torch.from_numpy(snap)); r.x.copy_(torch.from_numpy(live))
            torch.cuda.synchronize()
            PeerRegion.reconcile_async_local(regions, stream.cuda_stream)
            torch.cuda.synchronize()
            for r in regions:
                r.check()
            want_base = _expected(base, snaps, 5.6e-16 if f64 else 3e-7)
            tol = 1e-8 if f64 else 0.5
            new_base = base.copy()
            for g, r in enumerate(regions):
                own = _own_slice(n, 8 if f64 else 4, G, g)          # every rank keeps the base of its own slice
                new_base[own] = r.x_sync.cpu().numpy()[own]
            assert np.allclose(new_base, want_base, rtol=0, atol=tol), "the owners' x_sync slices are not the new common base"
            for g, r in enumerate(regions):
                want_live = lives[g].astype(np.float64) + (want_base.astype(np.float64) - snaps[g].astype(np.float64))
                assert np.allclose(r.x.cpu().numpy().astype(np.float64), want_live, rtol=0, atol=tol * 4), "live replica: wrong correction"
            base = new_base
    finally:
        for r in regions:
            r.close()
