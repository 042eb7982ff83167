"""ctypes harness around oracle/libgfs_oracle.so — the CPU ORACLE (test infrastructure).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``gfasort_b200`` never does.

The oracle restates /root/reference/src/sgd.rs (see gfs_oracle.cpp for the file:line map).  This
module adds the graph container the restatement reads — the three fields of the reference's
``BidirectedGraph`` that the hot path touches (``nodes[*].sequence.len()``, ``paths[*].steps``,
``node_order``; sgd.rs:41-55, 276-294) — and a GFA reader that follows the CLI's ``parse_gfa``
(src/bin/gfasort.rs:88-167: numeric ids, S lines first through ``add_node`` which appends new ids
to ``node_order`` (graph_ops.rs:613-623), then P lines; L lines are irrelevant to this path).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgfs_oracle.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++).  Building the checker is not using it."""
    src = os.path.join(_HERE, "gfs_oracle.cpp")
    src2 = os.path.join(_HERE, "graph_oracle.cpp")
    newest = max(os.path.getmtime(src), os.path.getmtime(src2))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
    return _LIB_PATH


class _Graph(C.Structure):
    _fields_ = [("present", u8p), ("seq_len", u64p), ("nodes_len", C.c_uint64),
                ("node_order", u64p), ("node_order_len", C.c_uint64),
                ("steps", u64p), ("path_first", u64p), ("num_paths", C.c_uint64)]


class OracleParams(C.Structure):
    """PathSGDParams / LayoutSGDParams, field for field (sgd.rs:196-212, 676-707)."""
    _fields_ = [("iter_max", C.c_uint64), ("iter_with_max_learning_rate", C.c_uint64),
                ("min_term_updates", C.c_uint64), ("delta", C.c_double), ("eps", C.c_double),
                ("eta_max", C.c_double), ("theta", C.c_double), ("space", C.c_uint64),
                ("space_max", C.c_uint64), ("space_quantization_step", C.c_uint64),
                ("cooling_start", C.c_double), ("nthreads", C.c_uint64), ("progress", C.c_uint64),
                ("seed", C.c_uint64)]

    def copy(self) -> "OracleParams":
        o = OracleParams()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(self))
        return o

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class OracleStats(C.Structure):
    _fields_ = [("applied", C.c_uint64), ("attempts", C.c_uint64), ("seconds", C.c_double),
                ("epochs", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_fast_precise_pow.restype = C.c_double
        L.oracle_fast_precise_pow.argtypes = [C.c_double, C.c_double]
        L.oracle_dirty_zipf.restype = C.c_uint64
        L.oracle_dirty_zipf.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double]
        L.oracle_schedule.restype = None
        L.oracle_schedule.argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_double, f64p]
        L.oracle_zeta_size.restype = C.c_uint64
        L.oracle_zeta_size.argtypes = [C.c_uint64] * 3
        L.oracle_zetas.restype = None
        L.oracle_zetas.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, f64p, C.c_uint64, C.c_uint64]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.oracle_set_epoch_window.restype = None
        L.oracle_set_epoch_window.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_xoshiro_u64.restype = None
        L.oracle_xoshiro_u64.argtypes = [C.c_uint64, C.c_uint64, u64p]
        L.oracle_xoshiro_below.restype = None
        L.oracle_xoshiro_below.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u64p]
        L.oracle_splitmix64.restype = None
        L.oracle_splitmix64.argtypes = [C.c_uint64, C.c_uint64, u64p]
        L.oracle_xoshiro_from_state.restype = None
        L.oracle_xoshiro_from_state.argtypes = [u64p, C.c_uint64, u64p]
        L.oracle_xoshiro_f64.restype = None
        L.oracle_xoshiro_f64.argtypes = [C.c_uint64, C.c_uint64, f64p]
        L.oracle_path_index.restype = None
        L.oracle_path_index.argtypes = [C.POINTER(_Graph)] + [u64p] * 7
        L.oracle_params_from_graph.restype = None
        L.oracle_params_from_graph.argtypes = [C.POINTER(_Graph), C.c_int, C.c_uint64, C.POINTER(OracleParams)]
        L.oracle_init_x.restype = C.c_uint64
        L.oracle_init_x.argtypes = [C.POINTER(_Graph), f64p]
        L.oracle_init_layout.restype = None
        L.oracle_init_layout.argtypes = [C.POINTER(_Graph), C.c_uint64, C.c_uint64, f64p]
        L.oracle_path_linear_sgd.restype = C.c_int
        L.oracle_path_linear_sgd.argtypes = [C.POINTER(_Graph), C.POINTER(OracleParams), C.c_int, C.c_int,
                                             C.c_uint32, C.c_uint64, f64p, C.c_uint64, C.POINTER(OracleStats)]
        L.oracle_index_create.restype = C.c_void_p
        L.oracle_index_create.argtypes = [C.POINTER(_Graph)]
        L.oracle_index_free.restype = None
        L.oracle_index_free.argtypes = [C.c_void_p]
        L.oracle_path_linear_sgd_ix.restype = C.c_int
        L.oracle_path_linear_sgd_ix.argtypes = [C.c_void_p] + L.oracle_path_linear_sgd.argtypes
        L.oracle_path_linear_sgd_layout.restype = C.c_int
        L.oracle_path_linear_sgd_layout.argtypes = [C.POINTER(_Graph), C.POINTER(OracleParams), C.c_uint64,
                                                    C.c_int, C.c_int, C.c_uint32, C.c_uint64, f64p,
                                                    C.c_uint64, C.POINTER(OracleStats)]
        L.oracle_trace_terms.restype = None
        L.oracle_trace_terms.argtypes = [C.POINTER(_Graph), C.POINTER(OracleParams), C.c_int, C.c_int,
                                         C.c_double, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                         u8p, u64p, u64p, u8p, f64p]
        L.oracle_layout_stress.restype = C.c_double
        L.oracle_layout_stress.argtypes = [C.POINTER(_Graph), C.c_uint64, f64p, C.c_uint64, C.c_int,
                                           C.c_uint64, C.c_uint32, f64p, u64p]
        L.oracle_sort_by_position.restype = None
        L.oracle_sort_by_position.argtypes = [f64p, C.c_uint64, u64p]
        for name in ("oracle_find_head_nodes", "oracle_groom", "oracle_topological_order"):
            fn = getattr(L, name)
            fn.restype = C.c_uint64
            fn.argtypes = [u8p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


MODE_REFERENCE, MODE_EXACT = 0, 1
DRAW_XOSHIRO, DRAW_PHILOX = 0, 1
STREAM_SGD, STREAM_STRESS = 1, 2


@dataclass
class Graph:
    """The slice of the reference's BidirectedGraph that the hot path reads.

    ``present[id]`` / ``seq_len[id]`` mirror ``nodes: Vec<Option<BiNode>>`` (graph_ops.rs:10-16);
    ``steps`` is the concatenation of every path's ``Vec<Handle>`` with ``Handle = id<<1 | is_rev``
    (graph.rs:9-19); ``path_first[p]..path_first[p+1]`` delimits path ``p``.
    """
    present: np.ndarray
    seq_len: np.ndarray
    node_order: np.ndarray
    steps: np.ndarray
    path_first: np.ndarray
    path_names: list = field(default_factory=list)

    def __post_init__(self):
        self.present = np.ascontiguousarray(self.present, dtype=np.uint8)
        self.seq_len = np.ascontiguousarray(self.seq_len, dtype=np.uint64)
        self.node_order = np.ascontiguousarray(self.node_order, dtype=np.uint64)
        self.steps = np.ascontiguousarray(self.steps, dtype=np.uint64)
        self.path_first = np.ascontiguousarray(self.path_first, dtype=np.uint64)

    @property
    def num_paths(self) -> int:
        return len(self.path_first) - 1

    @property
    def total_steps(self) -> int:
        return int(self.path_first[-1])

    def node_count(self) -> int:
        return int(self.present.sum())

    def c(self) -> _Graph:
        g = _Graph()
        g.present = _p(self.present, u8p)
        g.seq_len = _p(self.seq_len, u64p)
        g.nodes_len = len(self.present)
        g.node_order = _p(self.node_order, u64p)
        g.node_order_len = len(self.node_order)
        g.steps = _p(self.steps, u64p)
        g.path_first = _p(self.path_first, u64p)
        g.num_paths = self.num_paths
        return g

    # --- what the Rust side of the C ABI would do before calling the library ---
    def node_ids(self) -> np.ndarray:
        """sgd.rs:276-284."""
        if len(self.node_order):
            return self.node_order
        return np.nonzero(self.present)[0].astype(np.uint64)

    def dense(self):
        """Flatten to the C-ABI inputs: (step_handles = dense_idx<<1|rev, path_first, node_len[N]).

        Dense idx follows sgd.rs:286-293 (live nodes of node_ids, in order).  Steps on nodes that are
        not in the map get dense idx == N (the library's "missing node" sentinel).
        """
        ids = self.node_ids()
        live = ids[self.present[ids.astype(np.int64)] != 0] if len(ids) else ids
        n = len(live)
        id2idx = np.full(len(self.present) + 1, n, dtype=np.uint64)
        id2idx[live.astype(np.int64)] = np.arange(n, dtype=np.uint64)
        node_id = (self.steps >> np.uint64(1)).astype(np.int64)
        node_id = np.minimum(node_id, len(self.present))
        handles = (id2idx[node_id] << np.uint64(1)) | (self.steps & np.uint64(1))
        node_len = self.seq_len[live.astype(np.int64)].astype(np.uint32)
        return handles.astype(np.uint64), self.path_first.copy(), node_len

    @staticmethod
    def from_dense(step_handles: np.ndarray, path_first: np.ndarray, node_len: np.ndarray) -> "Graph":
        """Inverse of dense() for synthetic graphs: node id = dense idx + 1, node_order = 1..N."""
        n = len(node_len)
        present = np.zeros(n + 1, dtype=np.uint8)
        present[1:] = 1
        seq_len = np.zeros(n + 1, dtype=np.uint64)
        seq_len[1:] = node_len
        steps = step_handles.astype(np.uint64) + np.uint64(2)
        return Graph(present, seq_len, np.arange(1, n + 1, dtype=np.uint64), steps, path_first)


def parse_gfa(path: str) -> Graph:
    """Follows src/bin/gfasort.rs:88-167 for S and P lines (numeric ids; extra S fields ignored)."""
    with open(path) as f:
        lines = f.read().split("\n")
    ids, lens, order, seen = [], {}, [], set()
    for line in lines:
        if line.startswith("S"):
            parts = line.split("\t")
            if len(parts) >= 3:
                nid = int(parts[1])
                if nid not in seen:          # add_node: push to node_order only when new
                    seen.add(nid)
                    order.append(nid)
                lens[nid] = len(parts[2].encode())
    nodes_len = (max(lens) + 1) if lens else 0
    present = np.zeros(nodes_len, dtype=np.uint8)
    seq_len = np.zeros(nodes_len, dtype=np.uint64)
    for nid, l in lens.items():
        present[nid] = 1
        seq_len[nid] = l
    steps, first, names = [], [0], []
    for line in lines:
        if line.startswith("P"):
            parts = line.split("\t")
            if len(parts) >= 3:
                names.append(parts[1])
                for s in parts[2].split(","):
                    s = s.strip()
                    if not s:
                        continue
                    nid = int(s[:-1])
                    steps.append((nid << 1) | (0 if s[-1] == "+" else 1))
                first.append(len(steps))
    return Graph(present, seq_len, np.array(order, dtype=np.uint64), np.array(steps, dtype=np.uint64),
                 np.array(first, dtype=np.uint64), names)


# ------------------------------------------------------------------------------------------------
# thin functional wrappers
# ------------------------------------------------------------------------------------------------
def fast_precise_pow(a: float, b: float) -> float:
    return lib().oracle_fast_precise_pow(a, b)


def dirty_zipf(zmin, zmax, theta, zeta, zeta2theta, u) -> int:
    return lib().oracle_dirty_zipf(zmin, zmax, theta, zeta, zeta2theta, u)


def schedule(w_min, w_max, iter_max, iter_with_max_lr, eps) -> np.ndarray:
    etas = np.zeros(iter_max + 1)
    lib().oracle_schedule(w_min, w_max, iter_max, iter_with_max_lr, eps, _p(etas, f64p))
    return etas


def zetas(space, space_max, q, theta, iter_cap=0, size=None) -> np.ndarray:
    n = lib().oracle_zeta_size(space, space_max, q) if size is None else size
    out = np.zeros(n)
    lib().oracle_zetas(space, space_max, q, theta, _p(out, f64p), n, iter_cap)
    return out


def philox(ctr, key) -> np.ndarray:
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c, u32p), _p(k, u32p), _p(o, u32p))
    return o


def splitmix64(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_splitmix64(seed, n, _p(out, u64p))
    return out


def xoshiro_from_state(state4, n: int) -> np.ndarray:
    """n outputs of xoshiro256+ started from the raw state (rand_xoshiro `from_seed`, little-endian words)."""
    st = np.asarray(state4, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_xoshiro_from_state(_p(st, u64p), n, _p(out, u64p))
    return out


def xoshiro_u64(seed: int, n: int) -> np.ndarray:
    """n outputs of Xoshiro256Plus::seed_from_u64(seed)."""
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_xoshiro_u64(seed, n, _p(out, u64p))
    return out


def xoshiro_f64(seed: int, n: int) -> np.ndarray:
    """n draws of rng.random::<f64>() = (next_u64 >> 11) * 2^-53."""
    out = np.zeros(n, dtype=np.float64)
    lib().oracle_xoshiro_f64(seed, n, _p(out, f64p))
    return out


def xoshiro_below(seed: int, bound: int, n: int) -> np.ndarray:
    """n draws of Uniform::new(0, bound).sample(&mut rng) for usize (rand 0.9 UniformUsize)."""
    out = np.zeros(n, dtype=np.uint64)
    lib().oracle_xoshiro_below(seed, bound, n, _p(out, u64p))
    return out


def path_index(g: Graph) -> dict:
    S, P = g.total_steps, g.num_paths
    out = {k: np.zeros(S, dtype=np.uint64) for k in ("step_to_handle", "step_to_position", "step_to_path", "step_to_rank")}
    per = {k: np.zeros(P, dtype=np.uint64) for k in ("step_count", "length", "first_step")}
    cg = g.c()
    lib().oracle_path_index(C.byref(cg), _p(out["step_to_handle"], u64p), _p(out["step_to_position"], u64p),
                            _p(out["step_to_path"], u64p), _p(out["step_to_rank"], u64p),
                            _p(per["step_count"], u64p), _p(per["length"], u64p), _p(per["first_step"], u64p))
    out.update(per)
    return out


def params_from_graph(g: Graph, layout: bool = False, nthreads: int = 1) -> OracleParams:
    p = OracleParams()
    cg = g.c()
    lib().oracle_params_from_graph(C.byref(cg), int(layout), nthreads, C.byref(p))
    return p


def init_x(g: Graph) -> np.ndarray:
    x = np.zeros(g.node_count())
    cg = g.c()
    n = lib().oracle_init_x(C.byref(cg), _p(x, f64p))
    return x[:n]


def init_layout(g: Graph, dims: int, seed: int = 9399220) -> np.ndarray:
    n = len(g.node_ids())
    coords = np.zeros(n * 2 * dims)
    cg = g.c()
    lib().oracle_init_layout(C.byref(cg), dims, seed, _p(coords, f64p))
    return coords


class PrebuiltIndex:
    """PathIndex + handle map built once (bench.py's baseline leg reuses it across steps)."""

    def __init__(self, g: Graph):
        self._g = g
        cg = g.c()
        self._h = lib().oracle_index_create(C.byref(cg))

    def close(self):
        if self._h:
            lib().oracle_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_epoch_window(epoch_begin: int = 0, epoch_end: int | None = None):
    """Run only epochs [epoch_begin, epoch_end) of the schedule in the following SGD calls (bench.py's bounded CPU
    samples); no arguments = the whole run, as the reference."""
    lib().oracle_set_epoch_window(epoch_begin, (1 << 64) - 1 if epoch_end is None else epoch_end)


def path_linear_sgd(g: Graph, p: OracleParams, mode=MODE_REFERENCE, draw=DRAW_XOSHIRO, x0=None,
                    philox_tid_base=0, index: "PrebuiltIndex | None" = None):
    x = init_x(g) if x0 is None else np.array(x0, dtype=np.float64)
    st = OracleStats()
    cg = g.c()
    if index is not None:
        rc = lib().oracle_path_linear_sgd_ix(index._h, C.byref(cg), C.byref(p), mode, draw, STREAM_SGD,
                                             philox_tid_base, _p(x, f64p), len(x), C.byref(st))
    else:
        rc = lib().oracle_path_linear_sgd(C.byref(cg), C.byref(p), mode, draw, STREAM_SGD, philox_tid_base,
                                          _p(x, f64p), len(x), C.byref(st))
    return x, st, rc


def path_linear_sgd_layout(g: Graph, p: OracleParams, dims=2, mode=MODE_REFERENCE, draw=DRAW_XOSHIRO,
                           coords0=None, philox_tid_base=0):
    coords = init_layout(g, dims, p.seed) if coords0 is None else np.array(coords0, dtype=np.float64)
    n = len(coords) // (2 * dims)
    st = OracleStats()
    cg = g.c()
    rc = lib().oracle_path_linear_sgd_layout(C.byref(cg), C.byref(p), dims, mode, draw, STREAM_SGD,
                                             philox_tid_base, _p(coords, f64p), n, C.byref(st))
    return coords, st, rc


def trace_terms(g: Graph, p: OracleParams, nd: bool, cooling: bool, theta_cur: float, tid: int,
                attempt0: int, count: int):
    valid = np.zeros(count, dtype=np.uint8)
    sa = np.zeros(count, dtype=np.uint64)
    sb = np.zeros(count, dtype=np.uint64)
    fl = np.zeros(count, dtype=np.uint8)
    dist = np.zeros(count)
    cg = g.c()
    lib().oracle_trace_terms(C.byref(cg), C.byref(p), int(nd), int(cooling), theta_cur, STREAM_SGD, tid,
                             attempt0, count, _p(valid, u8p), _p(sa, u64p), _p(sb, u64p), _p(fl, u8p),
                             _p(dist, f64p))
    return valid, sa, sb, fl, dist


def layout_stress(g: Graph, coords: np.ndarray, dims: int, samples: int, draw=DRAW_XOSHIRO, seed=12345):
    """Returns (rms_rel [the reference's sgd.rs:1279 value], mean_abs_rel, counted)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    mar = C.c_double()
    cnt = C.c_uint64()
    cg = g.c()
    r = lib().oracle_layout_stress(C.byref(cg), dims, _p(coords, f64p), samples, draw, seed, STREAM_STRESS,
                                   C.byref(mar), C.byref(cnt))
    return r, mar.value, cnt.value


def x_as_layout(x: np.ndarray) -> np.ndarray:
    """1D positions as a dims=1 Layout (both ends at the node position) for the stress functions."""
    return np.repeat(np.asarray(x, dtype=np.float64), 2)


def sort_by_position(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    order = np.zeros(len(x), dtype=np.uint64)
    lib().oracle_sort_by_position(_p(x, f64p), len(x), _p(order, u64p))
    return order


# ------------------------------------------------------------------------------------------------
# `g` and `s`: literal (O(N*E)) restatements of groom / find_head_nodes / exact_odgi_topological_order
# ------------------------------------------------------------------------------------------------
def _graph_args(present, edges, steps, path_first):
    present = np.ascontiguousarray(present, dtype=np.uint8)
    edges = np.ascontiguousarray(edges, dtype=np.uint64).reshape(-1, 2)
    ef = np.ascontiguousarray(edges[:, 0]); et = np.ascontiguousarray(edges[:, 1])
    steps = np.ascontiguousarray(steps, dtype=np.uint64)
    path_first = np.ascontiguousarray(path_first, dtype=np.uint64)
    keep = (present, ef, et, steps, path_first)
    return keep, (_p(present, u8p), len(present), _p(ef, u64p), _p(et, u64p), len(ef), _p(steps, u64p),
                  _p(path_first, u64p), len(path_first) - 1)


def find_head_nodes(present, edges, steps, path_first) -> np.ndarray:
    keep, args = _graph_args(present, edges, steps, path_first)
    out = np.zeros(int(keep[0].sum()) + 1, dtype=np.uint64)
    n = lib().oracle_find_head_nodes(*args, _p(out, u64p))
    return out[:n]


def groom(present, edges, steps, path_first):
    """(handles in increasing node id, reverse when flipped; number flipped) — groom.rs:49-199, BFS mode."""
    keep, args = _graph_args(present, edges, steps, path_first)
    out = np.zeros(int(keep[0].sum()), dtype=np.uint64)
    nf = lib().oracle_groom(*args, _p(out, u64p))
    return out, nf


def topological_order(present, edges, steps, path_first) -> np.ndarray:
    """exact_odgi_topological_order(use_heads=true, use_tails=false) — graph_ops.rs:1232-1485."""
    keep, args = _graph_args(present, edges, steps, path_first)
    out = np.zeros(int(keep[0].sum()), dtype=np.uint64)
    n = lib().oracle_topological_order(*args, _p(out, u64p))
    return out[:n]
