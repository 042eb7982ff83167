"""The synthetic-graph generator (gfasort_b200/csrc/gfs_synth.cpp) as a host-only library, oracle/libgfs_synth.so.

Test / bench infrastructure: lets the CPU legs (bench.py --impl reference, the cpu_baseline leg, tools/oracle_*.py)
build exactly the graphs the GPU arm runs on WITHOUT loading libgfasort_cuda.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libgfs_synth.so")
u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)


class _Spec(C.Structure):
    _fields_ = [("num_nodes", C.c_uint64), ("num_paths", C.c_uint64), ("seed", C.c_uint64),
                ("permute_ids", C.c_uint32), ("pinned", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-C", _HERE, "libgfs_synth.so"], check=True, capture_output=True)
        L = C.CDLL(_LIB)
        L.gfs_synth_create_range.argtypes = [C.POINTER(_Spec), C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
        L.gfs_synth_dims.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.gfs_synth_arrays.argtypes = [C.c_void_p, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u32p)]
        L.gfs_synth_free.argtypes = [C.c_void_p]
        L.gfs_synth_free.restype = None
        L.gfs_synth_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def synth_arrays(num_nodes: int, num_paths: int, seed: int = 42, permute_ids: bool = True):
    """(step_handles u64[S], path_first u64[P+1], node_len u32[N]) as owned numpy arrays."""
    L = lib()
    spec = _Spec(num_nodes, num_paths, seed, int(permute_ids), 0)
    h = C.c_void_p()
    if L.gfs_synth_create_range(C.byref(spec), 0, num_paths, C.byref(h)) != 0:
        raise RuntimeError("gfs_synth: " + L.gfs_synth_last_error().decode())
    S, P, N = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.gfs_synth_dims(h, C.byref(S), C.byref(P), C.byref(N))
    ph, pf, pl = u64p(), u64p(), u32p()
    L.gfs_synth_arrays(h, C.byref(ph), C.byref(pf), C.byref(pl))
    handles = np.ctypeslib.as_array(ph, shape=(max(S.value, 1),))[:S.value].copy()
    first = np.ctypeslib.as_array(pf, shape=(P.value + 1,)).copy()
    node_len = np.ctypeslib.as_array(pl, shape=(N.value,)).copy()
    L.gfs_synth_free(h)
    return handles, first, node_len
