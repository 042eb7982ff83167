"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/gfasort_cuda.h
declares, fails loudly without a device, and the host-side mirror of the reference interface
(graph flattening, Layout, parameter derivation, synthetic generator) behaves like the reference.
No compute entry point is exercised for results here — that is tests/test_gpu_parity.py (-m gpu).
"""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest

from conftest import DATA, ROOT


def _has_cuda() -> bool:
    import torch
    return torch.cuda.is_available()


def test_header_symbols_all_exported(gfs):
    from gfasort_b200 import _cabi
    with open(os.path.join(ROOT, "include", "gfasort_cuda.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(gfs_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    L = C.CDLL(_cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.SIGNATURES), "ctypes binding and header disagree"


def test_struct_layouts_match_header(gfs):
    from gfasort_b200 import _cabi
    assert C.sizeof(_cabi.SgdParams) == 14 * 8
    assert C.sizeof(_cabi.Stats) == 8 * 8 + 4 * 4 + 8 + 2 * 4
    assert C.sizeof(_cabi.SynthSpec) == 3 * 8 + 2 * 4
    assert C.sizeof(_cabi.LaunchCfg) == 4 * 4 + 8 + 8 + 8 + 8 + 8


def test_product_never_imports_oracle():
    """The shipped package and library must not reference oracle/ (prompt ③)."""
    pkg = os.path.join(ROOT, "gfasort_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "libgfs_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.skipif(_has_cuda(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly(gfs):
    from gfasort_b200._cabi import GFS_ERR_NO_DEVICE, lib, u32p, u64p
    steps = np.array([0, 2], dtype=np.uint64)
    first = np.array([0, 2], dtype=np.uint64)
    nl = np.array([1, 1], dtype=np.uint32)
    h = C.c_void_p()
    rc = lib().gfs_index_build(steps.ctypes.data_as(u64p), first.ctypes.data_as(u64p), nl.ctypes.data_as(u32p),
                               2, 1, 2, C.byref(h))
    assert rc == GFS_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib().gfs_last_error()
    with pytest.raises(gfs.GfsError):
        gfs.PathIndex.from_arrays(steps, first, nl)
    # every other device entry point says the same
    x = np.zeros(4)
    order = np.zeros(4, dtype=np.uint32)
    from gfasort_b200._cabi import f64p
    assert lib().gfs_sort_positions(x.ctypes.data_as(f64p), 4, order.ctypes.data_as(u32p)) == GFS_ERR_NO_DEVICE
    assert lib().gfs_debug_fast_precise_pow(x.ctypes.data_as(f64p), x.ctypes.data_as(f64p), x.ctypes.data_as(f64p), 4) == GFS_ERR_NO_DEVICE
    r = C.c_void_p()
    assert lib().gfs_p2p_region_create(-1, 16, 8, 0, C.byref(r)) == GFS_ERR_NO_DEVICE and not r.value
    assert b"no CPU fallback" in lib().gfs_last_error()


def test_index_build_rejects_bad_arguments(gfs):
    from gfasort_b200._cabi import GFS_ERR_INVALID, lib, u32p, u64p
    steps = np.array([0, 2], dtype=np.uint64)
    nl = np.array([1, 1], dtype=np.uint32)
    h = C.c_void_p()
    bad_first = np.array([1, 2], dtype=np.uint64)          # must start at 0
    rc = lib().gfs_index_build(steps.ctypes.data_as(u64p), bad_first.ctypes.data_as(u64p), nl.ctypes.data_as(u32p),
                               2, 1, 2, C.byref(h))
    assert rc == GFS_ERR_INVALID and b"path_first_step" in lib().gfs_last_error()
    rc = lib().gfs_index_build(None, None, None, 2, 1, 2, C.byref(h))
    assert rc == GFS_ERR_INVALID


# ---- host-side mirror ---------------------------------------------------------------------------
def test_load_gfa_and_dense_flattening(gfs, oracle):
    for name in ("simple", "lil", "DRB1-3123"):
        path = os.path.join(DATA, f"{name}.gfa")
        g = gfs.load_gfa(path)
        og = oracle.parse_gfa(path)
        assert np.array_equal(g.steps, og.steps) and np.array_equal(g.path_first, og.path_first)
        assert np.array_equal(g.node_order, og.node_order) and np.array_equal(g.seq_len, og.seq_len)
        h, f, nl = g.dense()
        oh, of, onl = og.dense()
        assert np.array_equal(h, oh) and np.array_equal(f, of) and np.array_equal(nl, onl)
        assert np.array_equal(gfs.initial_positions(g), oracle.init_x(og))


def test_dense_handles_missing_nodes(gfs):
    present = np.array([0, 1, 0, 1], dtype=np.uint8)               # ids 1 and 3 live, 2 dead
    g = gfs.BidirectedGraph(present, np.array([0, 4, 9, 6]), np.array([1, 2, 3]),
                            np.array([2, 5, 6, 40]), np.array([0, 4]))
    h, f, nl = g.dense()
    assert nl.tolist() == [4, 6]
    assert h.tolist() == [0 << 1, (2 << 1) | 1, 1 << 1, 2 << 1]    # dead/unknown ids -> sentinel N = 2


def test_layout_params_from_graph(gfs):
    g = gfs.load_gfa(os.path.join(DATA, "DRB1-3123.gfa"))
    p = gfs.LayoutSGDParams.from_graph(g, 2, 4)
    assert (p.min_term_updates, p.space, p.eta_max, p.iter_max, p.space_max) == (350590, 3100, 9610000.0, 30, 1000)
    d = gfs.PathSGDParams()
    assert (d.iter_max, d.theta, d.eps, d.seed) == (100, 0.99, 0.01, 9399220)      # ygs.rs:226-231


def test_layout_value_type(gfs):
    """src/layout.rs:258-341 — the reference's five unit tests."""
    L = gfs.Layout
    lay = L.new(2, 10)
    assert len(lay.coords) == 40
    lay.set(3, 1, 1, 7.5)
    assert lay.get(3, 1, 1) == 7.5 and lay.coords[3 * 4 + 2 + 1] == 7.5
    lay = L.new(2, 2)
    lay.set(0, 0, 0, 0.0); lay.set(0, 0, 1, 0.0); lay.set(1, 0, 0, 3.0); lay.set(1, 0, 1, 4.0)
    assert abs(lay.distance(0, 0, 1, 0) - 5.0) < 1e-10
    lay = L.from_vectors([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]])
    assert lay.num_nodes == 2 and lay.coords.tolist() == [1.0, 5.0, 2.0, 6.0, 3.0, 7.0, 4.0, 8.0]
    lay = L(2, 2, np.array([1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5]))
    buf = io.StringIO()
    lay.write_tsv(buf)
    assert buf.getvalue().split("\n")[0] == "idx\tx+\ty+\tx-\ty-"
    back = L.read_tsv(io.StringIO(buf.getvalue()))
    assert back.dimensions == 2 and np.allclose(back.coords, lay.coords, atol=1e-10)


def test_apply_ordering_renumbers(gfs):
    g = gfs.load_gfa(os.path.join(DATA, "simple.gfa"))
    n, e, s = g.node_count(), len(g.edges), len(g.steps)
    order = (g.live_node_ids()[::-1].astype(np.uint64)) << np.uint64(1)          # reverse file order
    old_steps = g.steps.copy()
    g.apply_ordering(order)
    assert (g.node_count(), len(g.edges), len(g.steps)) == (n, e, s)             # integration_tests.rs:47-50
    assert np.array_equal(g.steps >> np.uint64(1), np.uint64(n + 1) - (old_steps >> np.uint64(1)))
    assert np.array_equal(g.steps & np.uint64(1), old_steps & np.uint64(1))


# ---- synthetic generator (host code inside the library; needs no GPU) ----------------------------
def test_synth_graph_shape_and_determinism(gfs):
    a = gfs.SynthGraph(20_000, 6, seed=42)
    b = gfs.SynthGraph(20_000, 6, seed=42)
    assert (a.N, a.P) == (20_000, 6) and a.S == int(a.path_first[-1])
    assert np.array_equal(a.step_handles, b.step_handles) and np.array_equal(a.node_len, b.node_len)
    counts = np.diff(a.path_first)
    assert counts.min() > 0.8 * a.N and counts.max() < 1.05 * a.N               # steps/path ~ 0.9 N
    nodes = (a.step_handles >> np.uint64(1))
    assert int(nodes.max()) < a.N
    assert a.node_len.min() >= 1 and a.node_len.max() <= 1024
    rev = (a.step_handles & np.uint64(1)).mean()
    assert 0 < rev < 0.2                                                        # inversions exist, are rare
    c = gfs.SynthGraph(20_000, 6, seed=43)
    assert not np.array_equal(a.step_handles[:1000], c.step_handles[:1000])
    # a rank's path range is the same walk as in the whole graph
    r = gfs.SynthGraph(20_000, 6, seed=42, path_begin=2, path_end=4)
    lo, hi = int(a.path_first[2]), int(a.path_first[4])
    assert np.array_equal(r.step_handles, a.step_handles[lo:hi])
    from gfasort_b200.synth import synth_path_counts
    assert np.array_equal(synth_path_counts(20_000, 6, 42), counts)


@pytest.mark.parametrize("space,max_steps,theta,space_max", [(50, 10, 0.99, 100), (15931, 3100, 0.99, 100), (3100, 3100, 0.99, 1000),
                                                              (27624835, 1_000_000, 0.99, 100), (400_000, 400_000, 0.001, 100)])
def test_host_zeta_table_matches_oracle_bit_for_bit(space, max_steps, theta, space_max, gfs, oracle):
    """The library's zeta table (terms computed block-wise on several threads, added up in the reference's serial order,
    truncated at the reachable index) against the oracle's plain serial loop (src/sgd.rs:311-331): identical bits."""
    import time
    from gfasort_b200._cabi import check, f64p, lib, u64p
    p = gfs.PathSGDParams(iter_max=100, min_term_updates=1, eta_max=1.0, theta=theta, space=space, space_max=space_max,
                          space_quantization_step=100)
    cp = p.c()
    n = C.c_uint64()
    check(lib().gfs_debug_zetas_host(C.byref(cp), max_steps, None, 0, C.byref(n)))
    got = np.zeros(n.value)
    t0 = time.perf_counter()
    check(lib().gfs_debug_zetas_host(C.byref(cp), max_steps, got.ctypes.data_as(f64p), n.value, C.byref(n)))
    dt = time.perf_counter() - t0
    want = oracle.zetas(space, space_max, 100, theta, iter_cap=min(space, max_steps), size=n.value)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"first difference at {int(np.nonzero(got != want)[0][0])}"
    assert dt < 2.0
