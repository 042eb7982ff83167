"""CPU tests (-m "not gpu") of the host steps downstream of `Y`: the library's
linear-time find_head_nodes / groom / topological sort must emit EXACTLY the orders of the literal
O(N*E) restatement of the reference in oracle/graph_oracle.cpp, on the reference's fixtures and on
random bidirected graphs (cycles, inversions, self loops, edges stored in complement form, edges to
missing nodes, several components, graphs without heads); plus the reference's pipeline invariants
(tests/integration_tests.rs: node / edge counts kept; path sequences unchanged)."""
import os
import time

import numpy as np
import pytest

from conftest import DATA

FIXTURES = ["simple", "lil", "DRB1-3123"]


def _both(gfs, oracle, g):
    h = gfs.find_head_nodes(g), oracle.find_head_nodes(g.present, g.edges, g.steps, g.path_first)
    gr = gfs.groom(g), oracle.groom(g.present, g.edges, g.steps, g.path_first)[0]
    t = gfs.exact_odgi_topological_order(g), oracle.topological_order(g.present, g.edges, g.steps, g.path_first)
    return h, gr, t


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_orders_match_reference_restatement(name, gfs, oracle):
    g = gfs.load_gfa(os.path.join(DATA, f"{name}.gfa"))
    for got, ref in _both(gfs, oracle, g):
        assert np.array_equal(got, ref)
    # and again on the graph the pipeline really hands to `s`: renumbered by a scrambled order, groomed
    rng = np.random.default_rng(1)
    order = rng.permutation(g.live_node_ids()).astype(np.uint64) << np.uint64(1)
    g.apply_ordering(order)
    gfs.groom_only(g)
    for got, ref in _both(gfs, oracle, g):
        assert np.array_equal(got, ref)


def _random_graph(gfs, rng, nodes_len, n_edges, n_paths, p_present=0.9, p_rev=0.3, p_missing_edge=0.05):
    present = (rng.random(nodes_len) < p_present).astype(np.uint8)
    present[0] = 0
    if present.sum() == 0:
        present[1 % nodes_len] = 1
    live = np.nonzero(present)[0]
    edges, seen = [], set()
    for _ in range(n_edges):
        a = int(rng.choice(live)) if rng.random() > p_missing_edge else int(rng.integers(1, nodes_len + 3))
        b = int(rng.choice(live)) if rng.random() > p_missing_edge else int(rng.integers(1, nodes_len + 3))
        fh = (a << 1) | int(rng.random() < p_rev)
        th = (b << 1) | int(rng.random() < p_rev)
        if (fh, th) not in seen and (th ^ 1, fh ^ 1) not in seen:          # add_edge's rule
            seen.add((fh, th)); edges.append((fh, th))
    steps, first = [], [0]
    for _ in range(n_paths):
        for _ in range(int(rng.integers(0, 12))):
            steps.append((int(rng.choice(live)) << 1) | int(rng.random() < p_rev))
        first.append(len(steps))
    seq_len = np.where(present, rng.integers(1, 9, nodes_len), 0)
    return gfs.BidirectedGraph(present, seq_len, live.astype(np.uint64), np.array(steps, dtype=np.uint64),
                               np.array(first, dtype=np.uint64), [], np.array(edges, dtype=np.uint64).reshape(-1, 2))


@pytest.mark.parametrize("seed", range(60))
def test_random_bidirected_graphs_match(seed, gfs, oracle):
    rng = np.random.default_rng(seed)
    nodes_len = int(rng.integers(2, 60))
    shape = seed % 4
    n_edges = [nodes_len // 2, nodes_len, 2 * nodes_len, 4 * nodes_len][shape]
    g = _random_graph(gfs, rng, nodes_len, n_edges, int(rng.integers(0, 5)),
                      p_present=[1.0, 0.9, 0.7, 0.95][shape], p_rev=[0.0, 0.3, 0.5, 0.1][shape],
                      p_missing_edge=[0.0, 0.05, 0.1, 0.0][shape])
    for got, ref in _both(gfs, oracle, g):
        assert np.array_equal(got, ref), (seed, got, ref)


def test_cycle_without_heads_and_self_loops(gfs, oracle):
    present = np.array([0, 1, 1, 1, 1], dtype=np.uint8)
    #  1+ -> 2+ -> 3+ -> 1+ (cycle), 3+ -> 3+ (self loop), 4- -> 4+ (inverting self loop), 2- -> 4+
    edges = np.array([[2, 4], [4, 6], [6, 2], [6, 6], [9, 8], [5, 8]], dtype=np.uint64)
    g = gfs.BidirectedGraph(present, np.array([0, 3, 3, 3, 3]), np.array([1, 2, 3, 4]), np.array([2, 4, 6, 2]),
                            np.array([0, 4]), [], edges)
    (h, ho), (gr, go), (t, to) = _both(gfs, oracle, g)
    assert len(ho) == 0 and np.array_equal(h, ho)
    assert np.array_equal(gr, go) and np.array_equal(t, to)
    assert sorted((t >> np.uint64(1)).tolist()) == [1, 2, 3, 4]


@pytest.mark.parametrize("name", FIXTURES)
def test_groom_and_sort_keep_the_graph(name, gfs):
    """integration_tests.rs:23-51, 91-145: counts preserved; plus: path sequences are unchanged."""
    from gfasort_b200 import ygs
    g = gfs.load_gfa(os.path.join(DATA, f"{name}.gfa"))
    n, e, s = g.node_count(), len(g.edges), len(g.steps)
    seqs = ygs.path_sequences(g)
    rng = np.random.default_rng(3)
    # flip a few nodes by hand first so that grooming has something to undo
    some = rng.choice(g.live_node_ids(), size=min(5, n), replace=False).astype(np.uint64)
    handles = g.live_node_ids().astype(np.uint64) << np.uint64(1)
    handles[np.isin(g.live_node_ids(), some)] |= np.uint64(1)
    gfs.apply_grooming_with_reorder(g, handles, True)
    assert ygs.path_sequences(g) == seqs
    gfs.groom_only(g)
    assert (g.node_count(), len(g.edges), len(g.steps)) == (n, e, s)
    assert ygs.path_sequences(g) == seqs
    gfs.topological_sort_only(g)
    assert (g.node_count(), len(g.edges), len(g.steps)) == (n, e, s)
    assert ygs.path_sequences(g) == seqs
    assert sorted(g.live_node_ids().tolist()) == list(range(1, n + 1))            # renumbered 1..N


def test_topological_sort_makes_dag_edges_forward(gfs):
    # simple.gfa and lil.gfa are DAGs: after `s` every edge goes from a lower to a higher id
    for name in ("simple", "lil"):
        g = gfs.load_gfa(os.path.join(DATA, f"{name}.gfa"))
        rng = np.random.default_rng(7)
        g.apply_ordering(rng.permutation(g.live_node_ids()).astype(np.uint64) << np.uint64(1))   # scramble ids
        gfs.groom_only(g)
        gfs.topological_sort_only(g)
        fwd, bwd = gfs.count_edge_directions(g)
        assert bwd == 0 and fwd == len(g.edges)


def test_linear_time_at_scale(gfs):
    """200k nodes / 280k edges: the reference's algorithms would need ~1e11 edge visits; these finish in seconds."""
    from gfasort_b200.graph import edges_from_paths
    s = gfs.SynthGraph(200_000, 8, seed=11)
    g = gfs.BidirectedGraph.from_dense(s.step_handles, s.path_first, s.node_len)
    g.edges = edges_from_paths(g.steps, g.path_first)
    n, e = g.node_count(), len(g.edges)
    lens_along = g.seq_len[(g.steps >> np.uint64(1)).astype(np.int64)].copy()
    t0 = time.time()
    gfs.groom_only(g)
    gfs.topological_sort_only(g)
    dt = time.time() - t0
    assert dt < 30, f"g + s took {dt:.1f}s"
    assert g.node_count() == n and len(g.edges) == e
    assert sorted(g.live_node_ids().tolist()) == list(range(1, n + 1))
    assert np.array_equal(g.seq_len[(g.steps >> np.uint64(1)).astype(np.int64)], lens_along)     # paths spell the same lengths
    fwd, bwd = gfs.count_edge_directions(g)
    assert fwd > 0.9 * (fwd + bwd)             # a bubble chain with a few inversions: almost everything forward
    ids = (g.steps >> np.uint64(1)).astype(np.int64)
    p0 = ids[int(g.path_first[0]):int(g.path_first[1])]
    assert (np.diff(p0) > 0).mean() > 0.9       # and the paths run (mostly) left to right
