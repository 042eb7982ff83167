"""Host-side mirror of the reference's SGD interface (src/sgd.rs, src/ygs.rs) over the C ABI.

Same names, argument meaning and error behaviour as the Rust functions; every body does what the
Rust side of the boundary would do: build the dense node numbering, flatten the
paths, compute the initial positions, call libgfasort_cuda, wrap the result.  All compute happens
in the CUDA library — nothing here falls back to the CPU.

    PathIndex.from_graph           src/sgd.rs:34-71     -> gfs_index_build / gfs_index_export
    PathSGDParams / YgsParams      src/sgd.rs:196-234, src/ygs.rs:16-92
    path_linear_sgd                src/sgd.rs:237-614   -> gfs_sgd_1d
    path_sgd_sort                  src/sgd.rs:641-672
    sgd_sort_only                  src/ygs.rs:195-206
    LayoutSGDParams                src/sgd.rs:676-763
    path_linear_sgd_layout         src/sgd.rs:773-1188  -> gfs_sgd_nd
    calculate_layout_stress        src/sgd.rs:1196-1283 -> gfs_stress
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass, replace

import numpy as np

from . import _cabi
from ._cabi import LaunchCfg, SgdParams, Stats, check, f64p, lib, u32p, u64p
from .graph import BidirectedGraph
from .layout import Layout


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------------------
# PathIndex
# ------------------------------------------------------------------------------------------------
class PathIndex:
    """Device-resident path index + the nine O(1) accessors of src/sgd.rs:73-107.

    The per-step offsets live on the GPU (one 16-byte record per step); `step_to_path`,
    `step_to_rank`, `first_step` and `step_count` are functions of `path_first_step` and are kept
    on the host.  Accessors that need offsets export them once, lazily."""

    def __init__(self, handle, step_handles, path_first, n_nodes, owner_graph=None):
        self._h = handle
        self._step_handles = step_handles        # dense handles (host copy, for get_handle_of_step)
        self._first = path_first
        self._n = n_nodes
        self._pos = None
        self._len = None
        self._graph = owner_graph

    @staticmethod
    def from_graph(graph: BidirectedGraph) -> "PathIndex":
        handles, first, node_len = graph.dense()
        return PathIndex.from_arrays(handles, first, node_len, graph, env=True)     # gfs_index_build: what the Rust host calls

    @staticmethod
    def from_arrays(step_handles: np.ndarray, path_first: np.ndarray, node_len: np.ndarray, graph=None,
                    path_begin: int = 0, path_end: int | None = None, device: int = -1,
                    relabel: int | None = None, new_of_old: np.ndarray | None = None, env: bool = False) -> "PathIndex":
        """step_handles: uint64 (Handle as the reference stores it) or uint32 (same value, half the copy).
        env=True: gfs_index_build / gfs_index_build32 — the whole graph, device / GPU count / relabelling from the
        environment (GFASORT_DEVICE, GFASORT_GPUS, GFASORT_RELABEL), i.e. what the Rust host's call does."""
        h32 = np.asarray(step_handles).dtype == np.uint32
        step_handles = np.ascontiguousarray(step_handles, dtype=np.uint32 if h32 else np.uint64)
        hp = u32p if h32 else u64p
        path_first = np.ascontiguousarray(path_first, dtype=np.uint64)
        node_len = np.ascontiguousarray(node_len, dtype=np.uint32)
        P = len(path_first) - 1
        if path_end is None:
            path_end = P
        h = C.c_void_p()
        if env:
            assert path_begin == 0 and path_end == P and new_of_old is None and relabel is None
            fn = lib().gfs_index_build32 if h32 else lib().gfs_index_build
            check(fn(_p(step_handles, hp), _p(path_first, u64p), _p(node_len, u32p), len(step_handles), P, len(node_len), C.byref(h)))
            return PathIndex(h, step_handles, path_first, len(node_len), graph)
        if relabel is None:
            relabel = 2 if new_of_old is not None else int(os.environ.get("GFASORT_RELABEL", "1") != "0")
        perm = None
        if new_of_old is not None:
            perm = np.ascontiguousarray(new_of_old, dtype=np.uint32)
        fn = lib().gfs_index_build_shard32 if h32 else lib().gfs_index_build_shard
        check(fn(_p(step_handles, hp), _p(path_first, u64p), _p(node_len, u32p),
                 len(step_handles), P, len(node_len), path_begin, path_end, device,
                 relabel, _p(perm, u32p) if perm is not None else None, C.byref(h)))
        s0, s1 = int(path_first[path_begin]), int(path_first[path_end])
        return PathIndex(h, step_handles[s0:s1], path_first[path_begin:path_end + 1] - path_first[path_begin],
                         len(node_len), graph)

    def close(self):
        if self._h:
            lib().gfs_index_free(self._h)
            self._h = None

    def apply_relabel(self, new_of_old: np.ndarray) -> None:
        """Adopt another rank's node order (an index built with relabel=0; gfs_index_apply_relabel)."""
        perm = np.ascontiguousarray(new_of_old, dtype=np.uint32)
        check(lib().gfs_index_apply_relabel(self._h, _p(perm, u32p)))

    def build_info(self) -> dict:
        """Wall / copy / K1-kernel seconds, launches and devices of the build (gfs_index_build_info)."""
        b, c, k, al, rl, n, d = C.c_double(), C.c_double(), C.c_double(), C.c_double(), C.c_double(), C.c_uint64(), C.c_uint32()
        check(lib().gfs_index_build_info(self._h, C.byref(b), C.byref(c), C.byref(k), C.byref(al), C.byref(rl), C.byref(n), C.byref(d)))
        return {"build_seconds": b.value, "copy_seconds": c.value, "kernel_seconds": k.value, "alloc_seconds": al.value,
                "relabel_seconds": rl.value, "launches": n.value, "devices": d.value}

    def relabel_permutation(self) -> np.ndarray:
        """new_of_old[N]: the library's internal node order (identity when relabelling is off)."""
        out = np.zeros(self._n, dtype=np.uint32)
        check(lib().gfs_index_export_relabel(self._h, _p(out, u32p)))
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _export(self):
        """Step offsets (S x 8 bytes device -> host): only the accessors that need them pay for it."""
        if self._pos is None:
            self._pos = np.empty(self.get_total_steps(), dtype=np.uint64)
            check(lib().gfs_index_export(self._h, _p(self._pos, u64p), None))

    def _export_lengths(self):
        if self._len is None:
            self._len = np.zeros(self.num_paths(), dtype=np.uint64)
            check(lib().gfs_index_export(self._h, None, _p(self._len, u64p)))

    # --- accessors, names as in the reference ---
    def get_total_steps(self) -> int:
        return int(self._first[-1])

    def get_handle_of_step(self, step_idx: int) -> int:
        """Handle in the caller's id space when built from a graph, else the dense handle."""
        h = int(self._step_handles[step_idx])
        if self._graph is not None:
            live = self._graph.live_node_ids()
            idx = h >> 1
            if idx < len(live):
                return (int(live[idx]) << 1) | (h & 1)
            return int(self._graph.steps[step_idx])
        return h

    def get_position_of_step(self, step_idx: int) -> int:
        self._export()
        return int(self._pos[step_idx])

    def get_path_of_step(self, step_idx: int) -> int:
        return int(np.searchsorted(self._first, step_idx, side="right") - 1)

    def get_rank_of_step(self, step_idx: int) -> int:
        return step_idx - int(self._first[self.get_path_of_step(step_idx)])

    def get_path_step_count(self, path_idx: int) -> int:
        return int(self._first[path_idx + 1] - self._first[path_idx])

    def get_step_at_path_position(self, path_idx: int, rank: int) -> int:
        return int(self._first[path_idx]) + rank

    def num_paths(self) -> int:
        return len(self._first) - 1

    def get_path_length(self, path_idx: int) -> int:
        self._export_lengths()
        return int(self._len[path_idx])

    # bulk views (numpy) for tests / parameter derivation
    def step_positions(self) -> np.ndarray:
        self._export()
        return self._pos

    def path_lengths(self) -> np.ndarray:
        self._export_lengths()
        return self._len

    def path_step_counts(self) -> np.ndarray:
        return np.diff(self._first)


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
@dataclass
class PathSGDParams:
    """src/sgd.rs:196-234 (defaults = the reference's `Default`)."""
    iter_max: int = 100
    iter_with_max_learning_rate: int = 0
    min_term_updates: int = 100
    delta: float = 0.0
    eps: float = 0.01
    eta_max: float = 100.0
    theta: float = 0.99
    space: int = 100
    space_max: int = 100
    space_quantization_step: int = 100
    cooling_start: float = 0.5
    nthreads: int = 1
    progress: bool = False
    seed: int = 9399220

    def c(self) -> SgdParams:
        return SgdParams(self.iter_max, self.iter_with_max_learning_rate, self.min_term_updates, self.delta,
                         self.eps, self.eta_max, self.theta, self.space, self.space_max,
                         self.space_quantization_step, self.cooling_start, self.nthreads, int(self.progress),
                         self.seed)


@dataclass
class LayoutSGDParams(PathSGDParams):
    """src/sgd.rs:676-729."""
    dimensions: int = 2
    iter_max: int = 30
    space_max: int = 1000

    @staticmethod
    def from_graph(graph: BidirectedGraph, dimensions: int, nthreads: int) -> "LayoutSGDParams":
        """src/sgd.rs:733-763."""
        counts = np.diff(graph.path_first)
        total = int(counts.sum()) if len(counts) else 0
        mx = int(counts.max()) if len(counts) else 0
        return LayoutSGDParams(dimensions=dimensions, iter_max=30, iter_with_max_learning_rate=0,
                               min_term_updates=10 * total, delta=0.0, eps=0.01, eta_max=float(mx * mx),
                               theta=0.99, space=mx, space_max=1000, space_quantization_step=100,
                               cooling_start=0.5, nthreads=nthreads, progress=False, seed=9399220)


@dataclass
class YgsParams:
    """src/ygs.rs:16-46."""
    path_sgd: PathSGDParams
    verbose: int = 0

    @staticmethod
    def default() -> "YgsParams":
        return YgsParams(PathSGDParams(iter_max=100, min_term_updates=0, eta_max=0.0, space=0, space_max=100,
                                       space_quantization_step=100, cooling_start=0.5, nthreads=1,
                                       progress=False, seed=9399220), 0)

    @staticmethod
    def from_graph(graph: BidirectedGraph, verbose: int, nthreads: int, path_index: PathIndex | None = None) -> "YgsParams":
        """src/ygs.rs:50-92: min_term_updates = sum of path step counts, eta_max = (max step count)^2,
        space = max path length in bp (which needs the index: built on the GPU)."""
        params = YgsParams.default()
        params.verbose = verbose
        params.path_sgd.nthreads = nthreads
        params.path_sgd.progress = verbose >= 2
        own = path_index is None
        ix = PathIndex.from_graph(graph) if own else path_index
        counts = ix.path_step_counts()
        lens = ix.path_lengths()
        params.path_sgd.min_term_updates = int(counts.sum()) if len(counts) else 0
        mx = int(counts.max()) if len(counts) else 0
        params.path_sgd.eta_max = float(mx * mx)
        params.path_sgd.space = int(lens.max()) if len(lens) else 0
        if own:
            ix.close()
        if verbose >= 2:
            print("[ygs_sort] Calculated parameters:", file=sys.stderr)
            print(f"  min_term_updates: {params.path_sgd.min_term_updates}", file=sys.stderr)
            print(f"  eta_max: {params.path_sgd.eta_max}", file=sys.stderr)
            print(f"  space: {params.path_sgd.space}", file=sys.stderr)
        return params


# ------------------------------------------------------------------------------------------------
# 1D  `Y`
# ------------------------------------------------------------------------------------------------
def initial_positions(graph: BidirectedGraph) -> np.ndarray:
    """X[idx] = sum of the lengths of the nodes before idx in node_ids order (src/sgd.rs:286-293)."""
    live = graph.live_node_ids()
    lens = graph.seq_len[live.astype(np.int64)].astype(np.uint64)
    x = np.zeros(len(live), dtype=np.float64)
    if len(live) > 1:
        x[1:] = np.cumsum(lens[:-1]).astype(np.float64)
    return x


last_stats: dict = {}


def path_linear_sgd(graph: BidirectedGraph, params: PathSGDParams, path_index: PathIndex | None = None,
                    cfg: LaunchCfg | None = None) -> dict:
    """src/sgd.rs:237-614.  Returns {dense idx: position}; empty when the graph has no nodes or no
    path with more than one step (src/sgd.rs:242-244, 258-261)."""
    if graph.node_count() == 0:
        return {}
    x = path_linear_sgd_array(graph, params, path_index, cfg)
    if x is None:
        return {}
    return dict(enumerate(x.tolist()))


def path_linear_sgd_array(graph: BidirectedGraph, params: PathSGDParams, path_index: PathIndex | None = None,
                          cfg: LaunchCfg | None = None):
    """Same as path_linear_sgd but returns the positions as a numpy array (None when the reference
    would return an empty map)."""
    if graph.node_count() == 0:
        return None
    own = path_index is None
    ix = PathIndex.from_graph(graph) if own else path_index
    try:
        x = initial_positions(graph)
        st = Stats()
        cp = params.c()
        rc = lib().gfs_sgd_1d_cfg(ix.handle, C.byref(cp), C.byref(cfg) if cfg is not None else None,
                                  _p(x, f64p), C.byref(st))
        if rc == _cabi.GFS_ERR_NO_VALID_PATH:
            print("[path_sgd] No paths with multiple steps found", file=sys.stderr)
            return None
        check(rc)
        last_stats.clear()
        last_stats.update(st.as_dict())
        if params.progress:
            print(f"[path_sgd] Complete: {st.applied_updates} term updates", file=sys.stderr)
        return x
    finally:
        if own:
            ix.close()


def sort_positions(x: np.ndarray) -> np.ndarray:
    """Dense indices ordered by position (stable, ties by idx) — on the GPU (gfs_sort_positions)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    order = np.zeros(len(x), dtype=np.uint32)
    check(lib().gfs_sort_positions(_p(x, f64p), len(x), _p(order, u32p)))
    return order


def path_sgd_sort(graph: BidirectedGraph, params: PathSGDParams, path_index: PathIndex | None = None) -> np.ndarray:
    """src/sgd.rs:641-672: forward handles of all nodes, sorted by final position (stable; ties by
    dense idx — the reference's tie order is HashMap iteration order, i.e. unspecified).  SGD and sort
    both run on the device (gfs_sgd_sort_1d); only the order comes back."""
    if graph.node_count() == 0:
        return np.zeros(0, dtype=np.uint64)
    own = path_index is None
    ix = PathIndex.from_graph(graph) if own else path_index
    try:
        x = initial_positions(graph)
        order = np.zeros(len(x), dtype=np.uint32)
        st = Stats()
        cp = params.c()
        rc = lib().gfs_sgd_sort_1d(ix.handle, C.byref(cp), _p(x, f64p), _p(order, u32p), C.byref(st))
        if rc == _cabi.GFS_ERR_NO_VALID_PATH:
            print("[path_sgd] No paths with multiple steps found", file=sys.stderr)
            return np.zeros(0, dtype=np.uint64)
        check(rc)
        last_stats.clear()
        last_stats.update(st.as_dict())
        last_stats["positions"] = x
        node_ids = graph.node_ids()
        order = order[order < len(node_ids)]
        return node_ids[order.astype(np.int64)].astype(np.uint64) << np.uint64(1)
    finally:
        if own:
            ix.close()


def sgd_sort_only(graph: BidirectedGraph, params: PathSGDParams, verbose: int = 0) -> None:
    """src/ygs.rs:195-206."""
    if verbose >= 2:
        print("[path_sgd] Starting path-guided SGD", file=sys.stderr)
    ordering = path_sgd_sort(graph, params)
    graph.apply_ordering(ordering)
    if verbose >= 2:
        print("[path_sgd] Complete", file=sys.stderr)


# ------------------------------------------------------------------------------------------------
# nD  `L`
# ------------------------------------------------------------------------------------------------
def initial_layout(graph: BidirectedGraph, dims: int, seed: int) -> np.ndarray:
    """src/sgd.rs:816-854, Layout order.  dim 0: cumulative length (+ end) / + node length (- end);
    dims >= 1: N(0,1)*sqrt(2N) drawn node-major, + end's dims then - end's.  (The reference draws
    from xoshiro256+ through rand_distr's ziggurat; the stream here is numpy's — the init is random
    noise and parity on it is statistical)"""
    node_ids = graph.node_ids()
    n = len(node_ids)
    num_nodes = graph.node_count()
    coords = np.zeros((n, 2, dims), dtype=np.float64)
    inb = node_ids < np.uint64(len(graph.present))
    live = np.zeros(n, dtype=bool)
    live[inb] = graph.present[node_ids[inb].astype(np.int64)] != 0
    lens = np.zeros(n, dtype=np.uint64)
    lens[live] = graph.seq_len[node_ids[live].astype(np.int64)]
    cum = np.zeros(n, dtype=np.uint64)
    if n > 1:
        cum[1:] = np.cumsum(lens[:-1])
    coords[:, 0, 0] = cum.astype(np.float64)
    coords[:, 1, 0] = (cum + lens).astype(np.float64)
    if dims > 1:
        rng = np.random.Generator(np.random.PCG64(seed))
        noise = rng.standard_normal((n, 2, dims - 1)) * np.sqrt(num_nodes * 2.0)
        coords[:, :, 1:] = noise
    coords[~live] = 0.0
    return coords.reshape(-1)


def path_linear_sgd_layout(graph: BidirectedGraph, params: LayoutSGDParams, path_index: PathIndex | None = None,
                           cfg: LaunchCfg | None = None, coords0: np.ndarray | None = None) -> Layout:
    """src/sgd.rs:773-1188."""
    num_nodes = graph.node_count()
    dims = params.dimensions
    if num_nodes == 0:
        return Layout.new(dims, 0)
    own = path_index is None
    ix = PathIndex.from_graph(graph) if own else path_index
    try:
        coords = initial_layout(graph, dims, params.seed) if coords0 is None else np.array(coords0, dtype=np.float64)
        st = Stats()
        cp = params.c()
        rc = lib().gfs_sgd_nd_cfg(ix.handle, C.byref(cp), C.byref(cfg) if cfg is not None else None, dims,
                                  _p(coords, f64p), C.byref(st))
        if rc == _cabi.GFS_ERR_NO_VALID_PATH:
            print("[path_sgd_layout] No paths with multiple steps found", file=sys.stderr)
            return Layout.new(dims, num_nodes)
        check(rc)
        last_stats.clear()
        last_stats.update(st.as_dict())
        return Layout(dims, len(coords) // (2 * dims), coords)
    finally:
        if own:
            ix.close()


def calculate_layout_stress(graph: BidirectedGraph, layout: Layout, sample_count: int,
                            path_index: PathIndex | None = None, seed: int = 12345) -> float:
    """src/sgd.rs:1196-1283 (RMS relative error over a fixed seeded sample, + ends)."""
    return layout_stress(graph, layout.coords, layout.dimensions, sample_count, path_index, seed)[0]


def layout_stress(graph, coords: np.ndarray, dims: int, sample_count: int, path_index: PathIndex | None = None,
                  seed: int = 12345, layout_order: bool = True):
    """(rms_rel, mean_abs_rel, counted) — the reference's form and BASELINE.json's form."""
    own = path_index is None
    ix = PathIndex.from_graph(graph) if own else path_index
    try:
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        rms, mar, cnt = C.c_double(), C.c_double(), C.c_uint64()
        check(lib().gfs_stress(ix.handle, dims, int(layout_order), _p(coords, f64p), sample_count, seed,
                               C.byref(rms), C.byref(mar), C.byref(cnt)))
        return rms.value, mar.value, cnt.value
    finally:
        if own:
            ix.close()


def sort_stress(graph, x: np.ndarray, sample_count: int, path_index: PathIndex | None = None, seed: int = 12345):
    """Sampled stress of a 1D sort (positions x[N])."""
    return layout_stress(graph, x, 1, sample_count, path_index, seed, layout_order=False)


class PinnedArray:
    """A numpy view of page-locked host memory from the library (gfs_host_alloc): the flatten target a host uses when
    it wants the index build to run at PCIe rate.  Keep the object alive while the view is in use."""

    def __init__(self, n: int, dtype):
        self._p = C.c_void_p()
        dt = np.dtype(dtype)
        check(lib().gfs_host_alloc(max(n, 1) * dt.itemsize, C.byref(self._p)))
        buf = (C.c_char * (max(n, 1) * dt.itemsize)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dt, count=n)

    def close(self):
        if self._p:
            self.array = None
            lib().gfs_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
