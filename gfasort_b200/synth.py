"""Synthetic pangenome graphs through the library's generator (gfs_synth_*).

Returns numpy views over the generator's storage — no copy of the (possibly multi-GB) step array.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._cabi import SynthSpec, check, lib, u32p, u64p

# the named shapes of BASELINE.json `configs`
SHAPES = {
    "tiny": (2_000, 4),
    "small": (50_000, 8),
    "config2_1M_32": (1_000_000, 32),
    "config3_10M_90": (10_000_000, 90),
    "config5_100M_90": (100_000_000, 90),
}


class SynthGraph:
    def __init__(self, num_nodes: int, num_paths: int, seed: int = 42, permute_ids: bool = True,
                 path_begin: int = 0, path_end: int | None = None, pinned: bool = False):
        spec = SynthSpec(num_nodes, num_paths, seed, int(permute_ids), int(pinned))
        if path_end is None:
            path_end = num_paths
        self._h = C.c_void_p()
        check(lib().gfs_synth_create_range(C.byref(spec), path_begin, path_end, C.byref(self._h)))
        S, P, N = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().gfs_synth_dims(self._h, C.byref(S), C.byref(P), C.byref(N)))
        self.S, self.P, self.N = S.value, P.value, N.value
        ph, pf, pl = u64p(), u64p(), u32p()
        check(lib().gfs_synth_arrays(self._h, C.byref(ph), C.byref(pf), C.byref(pl)))
        self.step_handles = np.ctypeslib.as_array(ph, shape=(max(self.S, 1),))[:self.S]
        self.path_first = np.ctypeslib.as_array(pf, shape=(self.P + 1,))
        self.node_len = np.ctypeslib.as_array(pl, shape=(self.N,))

    def close(self):
        if self._h:
            self.step_handles = self.path_first = self.node_len = None
            lib().gfs_synth_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initial_positions(self) -> np.ndarray:
        """X init of path_linear_sgd for node_order = 1..N (reference src/sgd.rs:286-293)."""
        x = np.zeros(self.N, dtype=np.float64)
        np.cumsum(self.node_len[:-1], dtype=np.float64, out=x[1:])
        return x

    def derived_params(self, layout: bool = False) -> dict:
        """YgsParams::from_graph / LayoutSGDParams::from_graph inputs that need no index
        (reference src/ygs.rs:60-79, src/sgd.rs:736-754); `space` for `Y` needs path lengths."""
        counts = np.diff(self.path_first)
        return {"sum_path_step_count": int(counts.sum()), "max_path_step_count": int(counts.max())}


def synth_path_counts(num_nodes: int, num_paths: int, seed: int = 42) -> np.ndarray:
    """Steps per path of the synthetic graph, without generating the steps."""
    spec = SynthSpec(num_nodes, num_paths, seed, 1, 0)
    counts = np.zeros(num_paths, dtype=np.uint64)
    check(lib().gfs_synth_path_counts(C.byref(spec), counts.ctypes.data_as(u64p)))
    return counts
