"""The `g` and `s` steps and the `Ygs` pipeline (reference src/ygs.rs, src/groom.rs, src/graph_ops.rs)
over the library's linear-time host algorithms (gfasort_b200/csrc/gfs_host_graph.cpp).

    find_head_nodes                  src/graph_ops.rs:1138-1183  -> gfs_find_head_nodes
    groom (BFS mode)                 src/groom.rs:49-275         -> gfs_groom_order
    apply_grooming_with_reorder      src/groom.rs:533-605
    exact_odgi_topological_order     src/graph_ops.rs:1232-1485  -> gfs_topological_order
    groom_only / topological_sort_only / ygs_sort      src/ygs.rs:97-206

`Y` runs on the GPU (gfasort_b200.sgd); `g` and `s` are sequential graph walks and stay on the host,
as in the reference — only their O(N*E) cost is gone.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from ._cabi import check, lib, u8p, u64p
from .graph import BidirectedGraph
from .sgd import YgsParams, path_sgd_sort

_COMP = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def reverse_complement(seq: bytes) -> bytes:
    """src/graph.rs reverse_complement."""
    return seq.translate(_COMP)[::-1]


def _args(graph: BidirectedGraph):
    present = np.ascontiguousarray(graph.present, dtype=np.uint8)
    edges = np.ascontiguousarray(graph.edges, dtype=np.uint64).reshape(-1, 2)
    ef = np.ascontiguousarray(edges[:, 0])
    et = np.ascontiguousarray(edges[:, 1])
    steps = np.ascontiguousarray(graph.steps, dtype=np.uint64)
    first = np.ascontiguousarray(graph.path_first, dtype=np.uint64)
    keep = (present, ef, et, steps, first)
    p = lambda a, t: a.ctypes.data_as(t)
    return keep, (p(present, u8p), len(present), p(ef, u64p), p(et, u64p), len(ef), p(steps, u64p), p(first, u64p),
                  len(first) - 1)


def find_head_nodes(graph: BidirectedGraph) -> np.ndarray:
    keep, args = _args(graph)
    out = np.zeros(graph.node_count() + 1, dtype=np.uint64)
    n = C.c_uint64()
    check(lib().gfs_find_head_nodes(*args, out.ctypes.data_as(u64p), C.byref(n)))
    return out[:n.value]


def groom(graph: BidirectedGraph, verbose: bool = False) -> np.ndarray:
    """Handles of all nodes in increasing id; reverse = the node must be flipped (src/groom.rs:49-199)."""
    keep, args = _args(graph)
    out = np.zeros(graph.node_count(), dtype=np.uint64)
    nf = C.c_uint64()
    check(lib().gfs_groom_order(*args, out.ctypes.data_as(u64p), C.byref(nf)))
    if verbose:
        print(f"[groom] Flipped {nf.value} nodes", file=sys.stderr)
    return out


def apply_grooming_with_reorder(graph: BidirectedGraph, groomed: np.ndarray, reorder: bool = True) -> None:
    """src/groom.rs:533-605: reverse-complement flipped nodes, XOR the orientation of every edge end and
    path step on them, then renumber 1..N in the order given (apply_node_id_mapping, graph_ops.rs:36-84).
    Flip and renumbering are one pass of gfs_remap_handles over the steps and one over the edge ends."""
    from .graph import UNMAPPED, remap_handles, unique_rows
    groomed = np.asarray(groomed, dtype=np.uint64)
    flip_ids = (groomed[(groomed & np.uint64(1)) == 1] >> np.uint64(1)).astype(np.int64)
    size = len(graph.present)
    flip = np.zeros(max(size, int(flip_ids.max()) + 1 if len(flip_ids) else 0), dtype=np.uint8)
    flip[flip_ids] = 1                                       # the flip set holds ids, live or not (groom.rs:535-541)
    for nid in flip_ids.tolist():
        if nid in graph.sequences:
            graph.sequences[nid] = reverse_complement(graph.sequences[nid])
    mapping = np.full(size, UNMAPPED, dtype=np.uint64)      # unmapped ids keep their id (graph_ops.rs:44, 58-59, 78)
    max_new = 0
    if reorder:
        old_ids = (groomed >> np.uint64(1)).astype(np.int64)
        ranks = np.arange(1, len(groomed) + 1, dtype=np.uint64)
        if len(old_ids) and int(old_ids.max()) >= size:
            mapping = np.concatenate([mapping, np.full(int(old_ids.max()) + 1 - size, UNMAPPED, dtype=np.uint64)])
        mapping[old_ids] = ranks
        max_new = len(groomed)
        _renumber_nodes(graph, mapping, max_new)
    graph.steps = remap_handles(graph.steps, mapping, flip)
    if len(graph.edges):
        e = np.ascontiguousarray(graph.edges.T)
        graph.edges = unique_rows(remap_handles(e[0], mapping, flip), remap_handles(e[1], mapping, flip))


def _renumber_nodes(graph: BidirectedGraph, mapping: np.ndarray, max_new: int) -> None:
    """The node-table half of apply_node_id_mapping (graph_ops.rs:36-50)."""
    from .graph import UNMAPPED
    live = np.nonzero(graph.present)[0]
    new_ids = mapping[live]
    new_ids = np.where(new_ids == UNMAPPED, live.astype(np.uint64), new_ids).astype(np.int64)
    n_len = max(max_new, int(new_ids.max()) if len(new_ids) else 0) + 1
    present = np.zeros(n_len, dtype=np.uint8)
    seq_len = np.zeros(n_len, dtype=np.uint64)
    present[new_ids] = 1
    seq_len[new_ids] = graph.seq_len[live]
    if graph.sequences:
        graph.sequences = {int(n): graph.sequences[int(o)] for o, n in zip(live, new_ids) if int(o) in graph.sequences}
    graph.present, graph.seq_len = present, seq_len


def groom_only(graph: BidirectedGraph, verbose: int = 0) -> None:
    """src/ygs.rs:180-192."""
    if verbose >= 2:
        print("[groom] Starting grooming", file=sys.stderr)
    order = groom(graph, verbose >= 2)
    apply_grooming_with_reorder(graph, order, True)
    if verbose >= 2:
        print("[groom] Complete", file=sys.stderr)


def exact_odgi_topological_order(graph: BidirectedGraph) -> np.ndarray:
    """use_heads = true, use_tails = false (src/graph_ops.rs:1232-1485): forward handles in emitted order."""
    keep, args = _args(graph)
    out = np.zeros(graph.node_count(), dtype=np.uint64)
    n = C.c_uint64()
    check(lib().gfs_topological_order(*args, out.ctypes.data_as(u64p), C.byref(n)))
    return out[:n.value]


def topological_sort_only(graph: BidirectedGraph, verbose: int = 0) -> None:
    """src/ygs.rs:147-159."""
    if verbose >= 2:
        print("[topological_sort] Starting topological sort (heads only)", file=sys.stderr)
    graph.apply_ordering(exact_odgi_topological_order(graph))
    if verbose >= 2:
        print("[topological_sort] Complete", file=sys.stderr)


def ygs_sort(graph: BidirectedGraph, params: YgsParams) -> None:
    """src/ygs.rs:97-143: Y (GPU) -> g -> s, renumbering after every step."""
    if params.verbose >= 1:
        print("[ygs_sort] Starting Ygs pipeline (Y=SGD, g=groom, s=topological_sort)", file=sys.stderr)
    graph.apply_ordering(path_sgd_sort(graph, params.path_sgd))
    apply_grooming_with_reorder(graph, groom(graph, params.verbose >= 2), True)
    graph.apply_ordering(exact_odgi_topological_order(graph))
    if params.verbose >= 1:
        print("[ygs_sort] Ygs pipeline complete", file=sys.stderr)


def count_edge_directions(graph: BidirectedGraph):
    """(forward, backward) by node id (src/graph_ops.rs:1215-1227); self loops are not counted."""
    if len(graph.edges) == 0:
        return 0, 0
    f = graph.edges[:, 0] >> np.uint64(1)
    t = graph.edges[:, 1] >> np.uint64(1)
    return int((f < t).sum()), int((f > t).sum())


def path_sequences(graph: BidirectedGraph) -> list:
    """Spelled sequence of every path (what the reference hashes, src/graph_ops.rs:781-800)."""
    out = []
    for p in range(graph.num_paths):
        parts = []
        for h in graph.path_steps(p).tolist():
            s = graph.sequences[h >> 1]
            parts.append(reverse_complement(s) if h & 1 else s)
        out.append(b"".join(parts))
    return out
