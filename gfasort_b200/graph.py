"""The slice of the reference's graph model that the SGD hot path reads and that `Y` writes back.

Mirrors (by name and meaning) `Handle` (reference src/graph.rs:9-19) and the three fields of
`BidirectedGraph` the path touches: `nodes[*].sequence.len()`, `paths[*].steps`, `node_order`
(src/graph_ops.rs:10-16; read at src/sgd.rs:41-55, 276-294).  Everything else in the reference's
graph container (edges, grooming, topological sorts, unchop) is out of scope here; edges are only
carried along so `apply_ordering` can renumber them the way the reference does.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def handle_forward(node_id):
    return np.uint64(node_id) << np.uint64(1)


def handle_node_id(h):
    return h >> np.uint64(1)


def handle_is_reverse(h):
    return (h & np.uint64(1)) == 1


@dataclass
class BidirectedGraph:
    """nodes as parallel arrays indexed by node id; paths as one concatenated handle array."""
    present: np.ndarray                      # uint8[max_id+1], 1 <=> nodes[id].is_some()
    seq_len: np.ndarray                      # uint64[max_id+1]
    node_order: np.ndarray                   # uint64[], insertion order of add_node (graph_ops.rs:613-623)
    steps: np.ndarray                        # uint64[S] Handle = id<<1 | is_rev, all paths concatenated
    path_first: np.ndarray                   # uint64[P+1]
    path_names: list = field(default_factory=list)
    edges: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), dtype=np.uint64))   # (from, to) handles
    sequences: dict = field(default_factory=dict)   # id -> bytes, only when loaded from GFA text

    def __post_init__(self):
        self.present = np.ascontiguousarray(self.present, dtype=np.uint8)
        self.seq_len = np.ascontiguousarray(self.seq_len, dtype=np.uint64)
        self.node_order = np.ascontiguousarray(self.node_order, dtype=np.uint64)
        self.steps = np.ascontiguousarray(self.steps, dtype=np.uint64)
        self.path_first = np.ascontiguousarray(self.path_first, dtype=np.uint64)

    # ---- reference-named queries ----
    def node_count(self) -> int:
        return int(self.present.sum())

    @property
    def num_paths(self) -> int:
        return len(self.path_first) - 1

    def path_steps(self, p: int) -> np.ndarray:
        return self.steps[int(self.path_first[p]):int(self.path_first[p + 1])]

    def node_ids(self) -> np.ndarray:
        """`node_order` if non-empty, else sorted live ids (src/sgd.rs:276-284)."""
        if len(self.node_order):
            return self.node_order
        return np.nonzero(self.present)[0].astype(np.uint64)

    # ---- what the host does before crossing the C ABI (SURVEY.md §8b "Rust side of the call") ----
    def live_node_ids(self) -> np.ndarray:
        ids = self.node_ids()
        if len(ids) == 0:
            return ids
        inb = ids < np.uint64(len(self.present))
        keep = np.zeros(len(ids), dtype=bool)
        keep[inb] = self.present[ids[inb].astype(np.int64)] != 0
        return ids[keep]

    def dense(self):
        """(step_handles, path_first_step, node_len) in dense-idx space (src/sgd.rs:286-293).

        Steps on ids without a dense idx get idx == N, the library's missing-node sentinel."""
        live = self.live_node_ids()
        n = len(live)
        id2idx = np.full(len(self.present) + 1, n, dtype=np.uint64)
        id2idx[live.astype(np.int64)] = np.arange(n, dtype=np.uint64)
        nid = np.minimum((self.steps >> np.uint64(1)).astype(np.int64), len(self.present))
        handles = (id2idx[nid] << np.uint64(1)) | (self.steps & np.uint64(1))
        node_len = self.seq_len[live.astype(np.int64)].astype(np.uint32)
        return np.ascontiguousarray(handles, dtype=np.uint64), self.path_first.copy(), node_len

    @staticmethod
    def from_dense(step_handles, path_first, node_len) -> "BidirectedGraph":
        """Synthetic graphs: node id = dense idx + 1, ids 1..N in file order."""
        n = len(node_len)
        present = np.zeros(n + 1, dtype=np.uint8)
        present[1:] = 1
        seq_len = np.zeros(n + 1, dtype=np.uint64)
        seq_len[1:] = node_len
        steps = np.asarray(step_handles, dtype=np.uint64) + np.uint64(2)
        return BidirectedGraph(present, seq_len, np.arange(1, n + 1, dtype=np.uint64), steps,
                               np.asarray(path_first, dtype=np.uint64))

    def apply_ordering(self, ordering: np.ndarray) -> None:
        """Renumber nodes 1..N in `ordering` (handles); rewrite edges and path steps.

        Follows src/graph_ops.rs:1939-2025: new id = rank + 1, orientations are kept, steps/edges on
        ids not in the ordering are left/dropped as the reference does, and `node_order` is NOT
        updated (SURVEY.md §8 quirk 7 — harmless for contiguous ids).  The per-step rewrite is the
        library's flat, multi-threaded gfs_remap_handles (the reference: one HashMap probe per step)."""
        ordering = np.asarray(ordering, dtype=np.uint64)
        if len(ordering) == 0:
            return
        old_ids = (ordering >> np.uint64(1)).astype(np.int64)
        new_ids = np.arange(1, len(ordering) + 1, dtype=np.uint64)
        size = max(len(self.present), int(old_ids.max()) + 1)
        old_to_new = np.full(size, UNMAPPED, dtype=np.uint64)
        old_to_new[old_ids] = new_ids                          # a repeated node keeps its LAST rank (HashMap insert)
        in_range = old_ids < len(self.present)
        live = np.zeros(len(old_ids), dtype=bool)
        live[in_range] = self.present[old_ids[in_range]] != 0
        # a node listed twice lands where its last entry says
        dst = old_to_new[old_ids[live]].astype(np.int64)
        present = np.zeros(len(ordering) + 1, dtype=np.uint8)
        seq_len = np.zeros(len(ordering) + 1, dtype=np.uint64)
        present[dst] = 1
        seq_len[dst] = self.seq_len[old_ids[live]]
        if self.sequences:
            self.sequences = {int(n): self.sequences[int(o)] for o, n in zip(old_ids[live], dst)
                              if int(o) in self.sequences}
        self.present, self.seq_len = present, seq_len
        self.steps = remap_handles(self.steps, old_to_new)
        if len(self.edges):
            f = (self.edges[:, 0] >> np.uint64(1)).astype(np.int64)
            t = (self.edges[:, 1] >> np.uint64(1)).astype(np.int64)
            keep = (f < size) & (t < size)
            keep[keep] = (old_to_new[f[keep]] != UNMAPPED) & (old_to_new[t[keep]] != UNMAPPED)   # an unmapped end drops the edge
            e = np.ascontiguousarray(self.edges[keep].T)       # (2, E'): one contiguous array per end
            self.edges = unique_rows(remap_handles(e[0], old_to_new), remap_handles(e[1], old_to_new), sort=True)


UNMAPPED = np.uint64(0xFFFFFFFFFFFFFFFF)


def remap_handles(handles: np.ndarray, new_id: np.ndarray, flip: np.ndarray | None = None) -> np.ndarray:
    """A rewritten copy of `handles`: orientation ^= flip[id], id -> new_id[id] unless new_id[id] == UNMAPPED or
    id >= len(new_id).  gfs_remap_handles: flat and multi-threaded."""
    import ctypes as C

    from ._cabi import check, lib, u8p, u64p
    out = np.array(handles, dtype=np.uint64, copy=True, order="C")
    new_id = np.ascontiguousarray(new_id, dtype=np.uint64)
    fl = None if flip is None else np.ascontiguousarray(flip, dtype=np.uint8)
    check(lib().gfs_remap_handles(out.ctypes.data_as(u64p), len(out), new_id.ctypes.data_as(u64p), len(new_id),
                                  fl.ctypes.data_as(u8p) if fl is not None else None, len(fl) if fl is not None else 0))
    return out


def unique_rows(a: np.ndarray, b: np.ndarray, sort: bool = False) -> np.ndarray:
    """HashSet<BiEdge> semantics for the pairs (a[i], b[i]): exact duplicates collapse.  (E', 2) array; rows in
    increasing (a, b) order when `sort`, else in order of first occurrence."""
    if len(a) == 0:
        return np.zeros((0, 2), dtype=np.uint64)
    if sort and int(a.max()) < (1 << 32) and int(b.max()) < (1 << 32):
        key = np.sort((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64))     # one 1-D sort, then adjacent dedupe
        keep = np.ones(len(key), dtype=bool)
        keep[1:] = key[1:] != key[:-1]
        key = key[keep]
        return np.stack([key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)], axis=1)
    idx = unique_pair_index(a, b)
    e = np.stack([a[idx], b[idx]], axis=1)
    return e[np.lexsort((e[:, 1], e[:, 0]))] if sort else e


def edges_from_paths(steps: np.ndarray, path_first: np.ndarray) -> np.ndarray:
    """The edge set a GFA writer would emit for a graph given only by its paths: every pair of consecutive
    steps, stored once per {edge, complement} like add_edge (graph_ops.rs:626-638), in order of first occurrence
    and in the form in which a path first walks it.  For synthetic graphs (gfs_edges_from_paths)."""
    import ctypes as C

    from ._cabi import check, lib, u64p
    steps = np.ascontiguousarray(steps, dtype=np.uint64)
    first = np.ascontiguousarray(path_first, dtype=np.uint64)
    h = C.c_void_p()
    check(lib().gfs_edges_from_paths(steps.ctypes.data_as(u64p), first.ctypes.data_as(u64p), len(first) - 1, C.byref(h)))
    try:
        pf, pt, n = u64p(), u64p(), C.c_uint64()
        check(lib().gfs_edge_list_get(h, C.byref(pf), C.byref(pt), C.byref(n)))
        if n.value == 0:
            return np.zeros((0, 2), dtype=np.uint64)
        return np.stack([np.ctypeslib.as_array(pf, shape=(n.value,)), np.ctypeslib.as_array(pt, shape=(n.value,))], axis=1).copy()
    finally:
        lib().gfs_edge_list_free(h)


def unique_pair_index(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Indices (increasing) of the first occurrence of every distinct pair (a[i], b[i])."""
    if len(a) == 0:
        return np.zeros(0, dtype=np.int64)
    if int(a.max()) < (1 << 32) and int(b.max()) < (1 << 32):
        key = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)          # one 1-D sort instead of a row sort
        order = np.argsort(key, kind="stable")                                      # stable: the first occurrence leads its run
        sk = key[order]
        lead = np.ones(len(sk), dtype=bool)
        lead[1:] = sk[1:] != sk[:-1]
        idx = order[lead]
    else:
        _, idx = np.unique(np.stack([a, b], axis=1), axis=0, return_index=True)
    idx.sort()
    return idx


def load_gfa(path: str) -> BidirectedGraph:
    """The CLI's `parse_gfa` (src/bin/gfasort.rs:88-167): numeric ids; S, then L, then P lines."""
    with open(path) as f:
        lines = f.read().split("\n")
    order, seqs = [], {}
    for line in lines:
        if line.startswith("S"):
            parts = line.split("\t")
            if len(parts) >= 3:
                nid = int(parts[1])
                if nid not in seqs:
                    order.append(nid)
                seqs[nid] = parts[2].encode()
    nodes_len = (max(seqs) + 1) if seqs else 0
    present = np.zeros(nodes_len, dtype=np.uint8)
    seq_len = np.zeros(nodes_len, dtype=np.uint64)
    for nid, s in seqs.items():
        present[nid] = 1
        seq_len[nid] = len(s)
    edges, seen = [], set()
    for line in lines:
        if line.startswith("L"):
            parts = line.split("\t")
            if len(parts) >= 5:
                fh = (int(parts[1]) << 1) | (0 if parts[2] == "+" else 1)
                th = (int(parts[3]) << 1) | (0 if parts[4] == "+" else 1)
                # add_edge (graph_ops.rs:626-638): skipped when the edge or its complement is already stored
                if (fh, th) not in seen and (th ^ 1, fh ^ 1) not in seen:
                    seen.add((fh, th))
                    edges.append((fh, th))
    steps, first, names = [], [0], []
    for line in lines:
        if line.startswith("P"):
            parts = line.split("\t")
            if len(parts) >= 3:
                names.append(parts[1])
                for s in parts[2].split(","):
                    s = s.strip()
                    if s:
                        steps.append((int(s[:-1]) << 1) | (0 if s[-1] == "+" else 1))
                first.append(len(steps))
    return BidirectedGraph(present, seq_len, np.array(order, dtype=np.uint64), np.array(steps, dtype=np.uint64),
                           np.array(first, dtype=np.uint64), names,
                           np.array(edges, dtype=np.uint64).reshape(-1, 2), seqs)
