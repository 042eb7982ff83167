"""GPU tests (-m gpu) of the peer-memory reconcile kernel (gfasort_b200/csrc/gfs_p2p.cu, K5b) on ONE device:
G replicas live on the same GPU and are connected with gfs_p2p_region_connect_local, one stream per "rank",
so the G kernels run concurrently and meet at their in-kernel barriers exactly as G ranks over NVLink would.
(The IPC path between processes needs >= 2 GPUs: bench.py --gpus N --reconcile p2p.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected(xs, xr):
    """x_sync + sum of displacements / #replicas that moved the element; exact when at most one moved it."""
    d = np.stack([x.astype(np.float64) - xs.astype(np.float64) for x in xr])
    moved = np.stack([x != xs for x in xr])
    cnt = moved.sum(0)
    mean = xs.astype(np.float64) + (d * moved).sum(0) / np.maximum(cnt, 1)
    out = mean.astype(xs.dtype)
    one = cnt == 1
    which = moved.argmax(0)
    out[one] = np.stack(xr)[which[one], np.nonzero(one)[0]]
    out[cnt == 0] = xs[cnt == 0]
    return out


@pytest.mark.parametrize("dtype,G,n", [("float64", 2, 100_003), ("float64", 4, 1_000_000), ("float32", 3, 65_537), ("float64", 1, 999)])
def test_p2p_reconcile_matches_moved_replica_mean(dtype, G, n, gfs, monkeypatch):
    import torch
    from gfasort_b200.multi import PeerRegion
    monkeypatch.setenv("GFASORT_P2P_SPIN_CAP", str(1 << 21))          # ~1 s: a missing peer is an error, not a hang
    f64 = dtype == "float64"
    regions = [PeerRegion(0, n, f64, max_blocks=8) for _ in range(G)]
    try:
        PeerRegion.connect_local(regions)
        streams = [torch.cuda.Stream(device=0) for _ in range(G)]
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
            for g, r in enumerate(regions):
                r.reconcile(streams[g].cuda_stream)
            torch.cuda.synchronize()
            for r in regions:
                r.check()
            want = _expected(xs, xr)
            got = [r.x.cpu().numpy() for r in regions]
            for g in range(G):
                assert np.array_equal(got[g], got[0]), "replicas differ after the reconcile"
                assert np.array_equal(regions[g].x_sync.cpu().numpy(), got[g]), "x_sync not refreshed"
            tol = 1e-9 if f64 else 1e-1
            assert np.allclose(got[0], want, rtol=0, atol=tol)
            one = np.stack([x != xs for x in xr]).sum(0) <= 1
            assert np.array_equal(got[0][one], want[one])              # exact where at most one replica moved it
            xs = got[0]
    finally:
        for r in regions:
            r.close()


def test_p2p_missing_peer_is_an_error_not_a_hang(gfs, monkeypatch):
    import torch
    from gfasort_b200.multi import PeerRegion
    monkeypatch.setenv("GFASORT_P2P_SPIN_CAP", str(1 << 12))          # a few ms
    regions = [PeerRegion(0, 1000, True, max_blocks=2) for _ in range(2)]
    try:
        PeerRegion.connect_local(regions)
        x0 = regions[0].x.clone()
        regions[0].reconcile(torch.cuda.current_stream().cuda_stream)  # rank 1 never shows up
        torch.cuda.synchronize()
        with pytest.raises(gfs.GfsError, match="barrier timed out"):
            regions[0].check()
        assert torch.equal(regions[0].x, x0)                           # nothing was touched
        regions[1].check()
    finally:
        for r in regions:
            r.close()


def test_p2p_region_argument_checks(gfs):
    from gfasort_b200.multi import PeerRegion
    a = PeerRegion(0, 10, True, max_blocks=2)
    b = PeerRegion(0, 11, True, max_blocks=2)
    try:
        with pytest.raises(gfs.GfsError):
            a.reconcile(0)                                             # not connected
        with pytest.raises(gfs.GfsError):
            PeerRegion.connect_local([a, b])                           # sizes differ
    finally:
        a.close(); b.close()
