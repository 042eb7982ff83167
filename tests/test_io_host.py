"""CPU tests (-m "not gpu") of the flat GFA ingest and the buffered writers: the
native one-pass parser must build the same graph as the line-by-line restatement of the CLI's parse_gfa
(gfasort_b200.graph.load_gfa, src/bin/gfasort.rs:88-167); the TSV writer must produce the bytes of
Layout::write_tsv (src/layout.rs:138-163, Rust `{}` float formatting); write_gfa must round-trip."""
import io
import os

import numpy as np
import pytest

from conftest import DATA


def _same_graph(a, b):
    assert np.array_equal(a.present, b.present) and np.array_equal(a.seq_len, b.seq_len)
    assert np.array_equal(a.node_order, b.node_order)
    assert np.array_equal(a.steps, b.steps) and np.array_equal(a.path_first, b.path_first)
    assert np.array_equal(a.edges, b.edges)
    assert a.path_names == b.path_names
    assert a.sequences == b.sequences


@pytest.mark.parametrize("name", ["simple", "lil", "DRB1-3123"])
def test_flat_ingest_matches_line_parser_on_fixtures(name, gfs):
    path = os.path.join(DATA, f"{name}.gfa")
    _same_graph(gfs.load_gfa_flat(path), gfs.load_gfa(path))


def test_flat_ingest_odd_inputs(gfs, tmp_path):
    text = ("H\tVN:Z:1.0\r\n"
            "P\tp1\t3+,1-, 2+ ,,7+\t*\n"                 # P before S; blanks and empty steps; a step on a missing node
            "S\t3\tACGT\tDP:i:4\tRC:i:9\n"               # extra tags ignored
            "S\t1\tA\n"
            "L\t3\t+\t1\t-\t0M\n"
            "L\t1\t+\t3\t-\t0M\n"                        # the complement of the previous edge: dropped
            "L\t3\t+\t1\t-\t5M\n"                        # duplicate: dropped
            "L\t1\t-\t2\t+\t0M\r\n"
            "S\t2\tGG\n"
            "S\t3\tTTTTT\n"                              # redefinition: sequence replaced, order kept
            "S\tshort\n"                                 # fewer than 3 fields: skipped
            "L\t1\t+\n"                                  # fewer than 5 fields: skipped
            "#comment\n"
            "P\tempty\t\t*\n"
            "P\tlast\t2-")                               # no trailing newline
    p = tmp_path / "odd.gfa"
    p.write_bytes(text.encode())
    a = gfs.load_gfa_flat(str(p))
    b = gfs.load_gfa(str(p))
    _same_graph(a, b)
    assert a.node_order.tolist() == [3, 1, 2] and a.seq_len.tolist() == [0, 1, 2, 5]
    assert a.edges.tolist() == [[6, 3], [3, 4]]
    assert a.steps.tolist() == [6, 3, 4, 14, 5] and a.path_first.tolist() == [0, 4, 4, 5]
    assert a.path_names == ["p1", "empty", "last"]
    c = gfs.load_gfa_flat(text=text.encode())
    _same_graph(a, c)
    with pytest.raises(gfs.GfsError):
        gfs.load_gfa_flat(text=b"S\tx1\tACGT\n")         # non-numeric id: the CLI errors out too
    # an absurd node id must come back as an error, not as an exception unwinding through the C ABI
    with pytest.raises(gfs.GfsError, match="out of host memory"):
        gfs.load_gfa_flat(text=b"S\t9999999999999999\tA\n")


def test_layout_tsv_bytes_match_rust_formatting(gfs, tmp_path):
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 1e6, 4000), rng.integers(-1000, 1000, 400).astype(float),
                           10.0 ** rng.uniform(-12, 15, 200), [0.0, -0.0, 1.0, -1.0, 0.1, 1e21, 1e-7, 123456789.125,
                           5e-324, 1.7976931348623157e308, float("nan"), float("inf"), float("-inf"), 0.30000000000000004,
                           2.5, 1e15, 1e16, 1e17, 9007199254740993.0]])
    for dims in (1, 2, 3, 5):
        n = len(vals) // (2 * dims)
        lay = gfs.Layout(dims, n, vals[:n * 2 * dims].copy())
        ref = io.StringIO()
        lay.write_tsv(ref)                                # the Python restatement of write_tsv + Rust `{}`
        path = str(tmp_path / f"l{dims}.tsv")
        nbytes = gfs.write_layout_tsv(lay, path)
        got = open(path, "rb").read()
        assert nbytes == len(got)
        assert got.decode() == ref.getvalue()
        back = gfs.Layout.read_tsv(io.StringIO(got.decode()))
        ok = np.isfinite(lay.coords)
        assert np.array_equal(back.coords[ok], lay.coords[ok])     # shortest digits round-trip exactly


@pytest.mark.parametrize("name", ["simple", "lil", "DRB1-3123"])
def test_write_gfa_round_trip(name, gfs, tmp_path):
    g = gfs.load_gfa(os.path.join(DATA, f"{name}.gfa"))
    out = str(tmp_path / "out.gfa")
    n = gfs.write_gfa(g, out)
    assert n == os.path.getsize(out)
    lines = open(out).read().split("\n")
    assert lines[0] == "H\tVN:Z:1.0"
    assert sum(l.startswith("S\t") for l in lines) == g.node_count()
    assert all(l.endswith("\t0M") for l in lines if l.startswith("L\t"))
    assert all(l.endswith("\t*") for l in lines if l.startswith("P\t"))
    h = gfs.load_gfa_flat(out)
    assert np.array_equal(h.present, g.present) and np.array_equal(h.seq_len, g.seq_len)
    assert np.array_equal(h.steps, g.steps) and h.path_names == g.path_names and h.sequences == g.sequences
    assert sorted(map(tuple, h.edges.tolist())) == sorted(map(tuple, g.edges.tolist()))     # L-line order is unspecified in the reference


def test_flat_ingest_is_fast_at_scale(gfs, tmp_path):
    """A 200k-node / 1.5M-step GFA parses in well under a second natively."""
    import time
    from gfasort_b200.graph import edges_from_paths
    s = gfs.SynthGraph(200_000, 8, seed=3)
    g = gfs.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
    g.edges = edges_from_paths(g.steps, g.path_first)
    g.sequences = {int(i): b"A" * int(g.seq_len[i]) for i in np.nonzero(g.present)[0]}
    g.path_names = [f"hap{k}" for k in range(g.num_paths)]
    out = str(tmp_path / "big.gfa")
    gfs.write_gfa(g, out)
    t0 = time.time()
    h = gfs.load_gfa_flat(out, with_sequences=False)
    dt = time.time() - t0
    assert np.array_equal(h.steps, g.steps) and np.array_equal(h.seq_len, g.seq_len) and len(h.edges) == len(g.edges)
    assert dt < 5.0
