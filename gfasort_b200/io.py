"""Flat GFA ingest and buffered writers over the library (gfasort_b200/csrc/gfs_io.cpp).

    load_gfa_flat      src/bin/gfasort.rs:88-167 (parse_gfa) in one native pass -> BidirectedGraph
    write_layout_tsv   src/layout.rs:138-163 (Layout::write_tsv), byte for byte
    write_gfa          src/graph_ops.rs:693-738 (BidirectedGraph::write_gfa)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._cabi import check, f64p, lib, u8p, u64p
from .graph import BidirectedGraph
from .layout import Layout


def load_gfa_flat(path: str | None = None, text: bytes | None = None, with_sequences: bool = True) -> BidirectedGraph:
    h = C.c_void_p()
    if text is not None:
        check(lib().gfs_gfa_parse_text(text, len(text), C.byref(h)))
    else:
        check(lib().gfs_gfa_parse_file(path.encode(), C.byref(h)))
    try:
        d = [C.c_uint64() for _ in range(5)]
        check(lib().gfs_gfa_dims(h, *[C.byref(x) for x in d]))
        nodes_len, n_nodes, n_edges, n_steps, n_paths = [x.value for x in d]
        pp, ps, po, pf, pt, pst, pfi = u8p(), u64p(), u64p(), u64p(), u64p(), u64p(), u64p()
        check(lib().gfs_gfa_arrays(h, C.byref(pp), C.byref(ps), C.byref(po), C.byref(pf), C.byref(pt), C.byref(pst), C.byref(pfi)))

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dtype=dt)

        present = arr(pp, nodes_len, np.uint8)
        seq_len = arr(ps, nodes_len, np.uint64)
        order = arr(po, n_nodes, np.uint64)
        edges = np.stack([arr(pf, n_edges, np.uint64), arr(pt, n_edges, np.uint64)], axis=1) if n_edges else np.zeros((0, 2), dtype=np.uint64)
        steps = arr(pst, n_steps, np.uint64)
        first = arr(pfi, n_paths + 1, np.uint64)
        txt, so, no, nl = C.c_void_p(), u64p(), u64p(), u64p()
        check(lib().gfs_gfa_text(h, C.byref(txt), C.byref(so), C.byref(no), C.byref(nl)))
        names, seqs = [], {}
        if n_paths:
            noff, nlen = arr(no, n_paths, np.uint64), arr(nl, n_paths, np.uint64)
            names = [C.string_at(txt.value + int(o), int(l)).decode() for o, l in zip(noff, nlen)]
        if with_sequences and nodes_len:
            soff = arr(so, nodes_len, np.uint64)
            for nid in np.nonzero(present)[0].tolist():
                seqs[nid] = C.string_at(txt.value + int(soff[nid]), int(seq_len[nid]))
        return BidirectedGraph(present, seq_len, order, steps, first, names, edges, seqs)
    finally:
        lib().gfs_gfa_free(h)


def write_layout_tsv(layout: Layout, path: str) -> int:
    coords = np.ascontiguousarray(layout.coords, dtype=np.float64)
    n = C.c_uint64()
    check(lib().gfs_layout_write_tsv(coords.ctypes.data_as(f64p), layout.num_nodes, layout.dimensions, path.encode(), C.byref(n)))
    return n.value


def write_gfa(graph: BidirectedGraph, path: str) -> int:
    nodes_len = len(graph.present)
    live = np.nonzero(graph.present)[0]
    blob = bytearray()
    seq_off = np.zeros(nodes_len, dtype=np.uint64)
    seq_len = np.zeros(nodes_len, dtype=np.uint64)
    for nid in live.tolist():
        s = graph.sequences.get(nid, b"")
        seq_off[nid] = len(blob); seq_len[nid] = len(s)
        blob += s
    names = list(graph.path_names) + [f"path{k}" for k in range(len(graph.path_names), graph.num_paths)]
    nblob = bytearray()
    noff = np.zeros(max(graph.num_paths, 1), dtype=np.uint64)
    nlen = np.zeros(max(graph.num_paths, 1), dtype=np.uint64)
    for k, nm in enumerate(names[:graph.num_paths]):
        b = nm.encode()
        noff[k] = len(nblob); nlen[k] = len(b)
        nblob += b
    edges = np.ascontiguousarray(graph.edges, dtype=np.uint64).reshape(-1, 2)
    ef, et = np.ascontiguousarray(edges[:, 0]), np.ascontiguousarray(edges[:, 1])
    present = np.ascontiguousarray(graph.present, dtype=np.uint8)
    steps = np.ascontiguousarray(graph.steps, dtype=np.uint64)
    first = np.ascontiguousarray(graph.path_first, dtype=np.uint64)
    p = lambda a, t: a.ctypes.data_as(t)
    n = C.c_uint64()
    check(lib().gfs_gfa_write(path.encode(), p(present, u8p), nodes_len, bytes(blob), p(seq_off, u64p), p(seq_len, u64p),
                              p(ef, u64p), p(et, u64p), len(ef), p(steps, u64p), p(first, u64p), graph.num_paths,
                              bytes(nblob), p(noff, u64p), p(nlen, u64p), C.byref(n)))
    return n.value
