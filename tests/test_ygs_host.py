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


# ------------------------------------------------------------------------------------------------
# apply_ordering / apply_grooming_with_reorder (the flat gfs_remap_handles passes) against a literal,
# dict-based restatement of the reference (src/graph_ops.rs:36-84, 1939-2025; src/groom.rs:533-605)
# ------------------------------------------------------------------------------------------------
def _ref_state(g):
    nodes = {int(i): (int(g.seq_len[i]), g.sequences.get(int(i))) for i in np.nonzero(g.present)[0]}
    edges = {(int(a), int(b)) for a, b in g.edges.tolist()}
    return nodes, len(g.present), edges, [int(h) for h in g.steps.tolist()]


def _ref_apply_ordering(state, ordering):
    nodes, nodes_len, edges, steps = state
    if not ordering:
        return state
    old_to_new = {}
    for i, h in enumerate(ordering):
        old_to_new[h >> 1] = i + 1                                   # HashMap insert: the last rank wins
    max_new = max(old_to_new.values())
    new_nodes = {new: nodes[old] for old, new in old_to_new.items() if old in nodes}
    new_edges = {((old_to_new[f >> 1] << 1) | (f & 1), (old_to_new[t >> 1] << 1) | (t & 1))
                 for f, t in edges if (f >> 1) in old_to_new and (t >> 1) in old_to_new}
    new_steps = [((old_to_new[h >> 1] << 1) | (h & 1)) if (h >> 1) in old_to_new else h for h in steps]
    return new_nodes, max_new + 1, new_edges, new_steps


def _ref_apply_grooming(state, groomed, reorder, rc):
    nodes, nodes_len, edges, steps = state
    flips = {h >> 1 for h in groomed if h & 1}
    nodes = {i: (l, rc(s) if (i in flips and s is not None) else s) for i, (l, s) in nodes.items()}
    fl = lambda h: h ^ 1 if (h >> 1) in flips else h
    edges = {(fl(f), fl(t)) for f, t in edges}
    steps = [fl(h) for h in steps]
    if reorder:
        m = {h >> 1: i + 1 for i, h in enumerate(groomed)}
        mp = lambda h: (m.get(h >> 1, h >> 1) << 1) | (h & 1)
        max_new = max(m.values()) if m else 0
        nodes = {m.get(i, i): v for i, v in nodes.items()}
        nodes_len = max(max_new, max(nodes) if nodes else 0) + 1      # the reference would panic past max_new
        edges = {(mp(f), mp(t)) for f, t in edges}
        steps = [mp(h) for h in steps]
    return nodes, nodes_len, edges, steps


def _check_state(g, want, tag):
    nodes, nodes_len, edges, steps = want
    got_nodes, got_len, got_edges, got_steps = _ref_state(g)
    assert got_nodes == nodes, tag
    assert got_len == nodes_len, tag
    assert got_edges == edges, tag
    assert got_steps == steps, tag
    assert len(g.edges) == len(edges), tag                           # HashSet semantics: no duplicate rows


@pytest.mark.parametrize("seed", range(80))
def test_apply_ordering_and_grooming_match_reference_semantics(seed, gfs):
    rng = np.random.default_rng(seed)
    g = _random_graph(gfs, rng, int(rng.integers(2, 50)), int(rng.integers(0, 120)), int(rng.integers(0, 5)),
                      p_present=0.85, p_rev=0.3, p_missing_edge=0.1)
    live = np.nonzero(g.present)[0]
    g.sequences = {int(i): bytes(rng.choice(list(b"ACGT"), int(g.seq_len[i])).tolist()) for i in live}
    # steps on ids that are not in the graph, too
    extra = (rng.integers(1, len(g.present) + 4, 5).astype(np.uint64) << np.uint64(1)) | rng.integers(0, 2, 5).astype(np.uint64)
    g.steps = np.concatenate([g.steps, extra]); g.path_first = np.append(g.path_first, np.uint64(len(g.steps)))
    kind = seed % 4
    if kind == 0:
        ids = rng.permutation(live)                                             # a full permutation
    elif kind == 1:
        ids = rng.permutation(live)[:max(1, len(live) // 2)]                    # partial: unmapped nodes / steps / edges
    elif kind == 2:
        ids = np.concatenate([rng.permutation(live), rng.choice(live, 2)])      # a node listed twice: the last rank wins
    else:
        ids = np.concatenate([rng.permutation(live), [len(g.present) + 3]])     # an id that is not in the graph
    order = (ids.astype(np.uint64) << np.uint64(1)) | (rng.random(len(ids)) < 0.3).astype(np.uint64)
    import copy
    for reorder in (None, True, False):
        h = copy.deepcopy(g)
        before = _ref_state(h)
        if reorder is None:
            h.apply_ordering(order)
            _check_state(h, _ref_apply_ordering(before, [int(x) for x in order]), ("apply_ordering", seed))
        else:
            groomed = (rng.permutation(live).astype(np.uint64) << np.uint64(1)) | (rng.random(len(live)) < 0.4).astype(np.uint64)
            gfs.apply_grooming_with_reorder(h, groomed, reorder)
            _check_state(h, _ref_apply_grooming(before, [int(x) for x in groomed], reorder, gfs.ygs.reverse_complement),
                         ("apply_grooming", seed, reorder))


def test_remap_handles_native(gfs):
    from gfasort_b200.graph import UNMAPPED, remap_handles
    rng = np.random.default_rng(3)
    n = 3_000_000                                                   # several threads
    h = rng.integers(0, 2 * 1000 + 40, n).astype(np.uint64)
    table = rng.integers(1, 5000, 1000).astype(np.uint64)
    table[rng.random(1000) < 0.2] = UNMAPPED
    flip = (rng.random(1010) < 0.5).astype(np.uint8)
    got = remap_handles(h, table, flip)
    ids = (h >> np.uint64(1)).astype(np.int64)
    rev = h & np.uint64(1)
    rev = np.where(ids < 1010, rev ^ flip[np.minimum(ids, 1009)], rev)
    mapped = np.where((ids < 1000) & (table[np.minimum(ids, 999)] != UNMAPPED), table[np.minimum(ids, 999)], ids.astype(np.uint64))
    assert np.array_equal(got, (mapped << np.uint64(1)) | rev)
    assert np.array_equal(remap_handles(h[:7], table), ((np.where((ids[:7] < 1000) & (table[np.minimum(ids[:7], 999)] != UNMAPPED), table[np.minimum(ids[:7], 999)], ids[:7].astype(np.uint64))) << np.uint64(1)) | (h[:7] & np.uint64(1)))
    assert len(remap_handles(np.zeros(0, dtype=np.uint64), table)) == 0


@pytest.mark.parametrize("seed", range(40))
def test_edges_from_paths_matches_add_edge_semantics(seed, gfs):
    """gfs_edges_from_paths == add_edge (graph_ops.rs:626-638) applied to every consecutive step pair in path order:
    one edge per {edge, complement} class, first occurrence wins, kept as first walked; empty / single-step paths."""
    from gfasort_b200.graph import edges_from_paths
    rng = np.random.default_rng(seed)
    steps, first = [], [0]
    for _ in range(int(rng.integers(0, 6))):
        steps += [(int(rng.integers(1, 8)) << 1) | int(rng.random() < 0.4) for _ in range(int(rng.integers(0, 15)))]
        first.append(len(steps))
    seen, want = set(), []
    for p in range(len(first) - 1):
        for s in range(first[p], first[p + 1] - 1):
            a, b = steps[s], steps[s + 1]
            if (a, b) not in seen and (b ^ 1, a ^ 1) not in seen:
                seen.add((a, b)); want.append((a, b))
    got = edges_from_paths(np.array(steps, dtype=np.uint64), np.array(first, dtype=np.uint64))
    assert got.shape == (len(want), 2) and got.tolist() == [list(e) for e in want]
