"""ctypes binding of include/gfasort_cuda.h (libgfasort_cuda.so).

This is the Python twin of the Rust `extern "C"` block in INTEGRATION.md.  Loading fails loudly
when the library is missing; compute entry points fail loudly when there is no CUDA device.  There
is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GFASORT_LIB_PATH") or os.path.join(_HERE, "libgfasort_cuda.so")   # override: experiments only

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)

GFS_OK, GFS_ERR_INVALID, GFS_ERR_CUDA, GFS_ERR_NO_DEVICE, GFS_ERR_NO_VALID_PATH = 0, 1, 2, 3, 4
GFS_P2P_HANDLE_BYTES = 80


class GfsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgfasort_cuda error {code}: {msg}")
        self.code = code


class SgdParams(C.Structure):
    """gfs_sgd_params == PathSGDParams / LayoutSGDParams (reference src/sgd.rs:196-212, 676-707)."""
    _fields_ = [("iter_max", C.c_uint64), ("iter_with_max_learning_rate", C.c_uint64),
                ("min_term_updates", C.c_uint64), ("delta", C.c_double), ("eps", C.c_double),
                ("eta_max", C.c_double), ("theta", C.c_double), ("space", C.c_uint64),
                ("space_max", C.c_uint64), ("space_quantization_step", C.c_uint64),
                ("cooling_start", C.c_double), ("nthreads", C.c_uint64), ("progress", C.c_uint64),
                ("seed", C.c_uint64)]


class LaunchCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("total_threads", C.c_uint32), ("aggregate", C.c_int32),
                ("layout_f64", C.c_int32), ("rng_thread_base", C.c_uint64), ("stream", C.c_void_p),
                ("device_positions", C.c_void_p), ("sample_begin", C.c_uint64), ("sample_end", C.c_uint64)]

    @staticmethod
    def default() -> "LaunchCfg":
        return LaunchCfg(-1, 0, -1, -1, 0, None, None, 0, 0)


class Stats(C.Structure):
    _fields_ = [("applied_updates", C.c_uint64), ("attempts", C.c_uint64), ("epochs", C.c_uint64),
                ("launches", C.c_uint64), ("kernel_seconds", C.c_double), ("h2d_seconds", C.c_double),
                ("d2h_seconds", C.c_double), ("total_seconds", C.c_double), ("grid", C.c_uint32),
                ("block", C.c_uint32), ("coord_bytes", C.c_uint32), ("n_devices", C.c_uint32),
                ("window_steps", C.c_uint64), ("coherent", C.c_uint32), ("syncs_per_epoch", C.c_uint32)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class ShardPlan(C.Structure):
    """gfs_shard_plan: who samples what in a replicated multi-GPU run."""
    _fields_ = [("sample_begin", C.c_uint64), ("sample_end", C.c_uint64), ("path_begin", C.c_uint64),
                ("path_end", C.c_uint64), ("first_step", C.c_uint64)]


class SynthSpec(C.Structure):
    _fields_ = [("num_nodes", C.c_uint64), ("num_paths", C.c_uint64), ("seed", C.c_uint64),
                ("permute_ids", C.c_uint32), ("pinned", C.c_uint32)]


# every symbol include/gfasort_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gfs_last_error": (C.c_char_p, []),
    "gfs_device_info": (C.c_char_p, []),
    "gfs_index_build": (C.c_int, [u64p, u64p, u32p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_index_build_shard": (C.c_int, [u64p, u64p, u32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.c_uint64, C.c_int32, C.c_int32, u32p, C.POINTER(C.c_void_p)]),
    "gfs_index_build32": (C.c_int, [u32p, u64p, u32p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_index_build_shard32": (C.c_int, [u32p, u64p, u32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_int32, C.c_int32, u32p, C.POINTER(C.c_void_p)]),
    "gfs_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_host_free": (None, [C.c_void_p]),
    "gfs_index_build_info": (C.c_int, [C.c_void_p, f64p, f64p, f64p, f64p, f64p, u64p, u32p]),
    "gfs_index_export_relabel": (C.c_int, [C.c_void_p, u32p]),
    "gfs_index_apply_relabel": (C.c_int, [C.c_void_p, u32p]),
    "gfs_index_export": (C.c_int, [C.c_void_p, u64p, u64p]),
    "gfs_index_export_records": (C.c_int, [C.c_void_p, u64p, u32p]),
    "gfs_index_dims": (C.c_int, [C.c_void_p, u64p, u64p, u64p, u64p]),
    "gfs_index_free": (None, [C.c_void_p]),
    "gfs_sgd_1d": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), f64p, C.POINTER(Stats)]),
    "gfs_sgd_nd": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.c_uint32, f64p, C.POINTER(Stats)]),
    "gfs_sgd_1d_cfg": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.POINTER(LaunchCfg), f64p, C.POINTER(Stats)]),
    "gfs_sgd_nd_cfg": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.POINTER(LaunchCfg), C.c_uint32, f64p,
                                 C.POINTER(Stats)]),
    "gfs_stress": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, f64p, C.c_uint64, C.c_uint64, f64p, f64p, u64p]),
    "gfs_stress_partial": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, f64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_uint64, f64p]),
    "gfs_sgd_session_create": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.c_uint32, C.POINTER(LaunchCfg),
                                         C.POINTER(C.c_void_p)]),
    "gfs_sgd_session_upload": (C.c_int, [C.c_void_p, f64p]),
    "gfs_sgd_session_download": (C.c_int, [C.c_void_p, f64p]),
    "gfs_sgd_session_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]),
    "gfs_sgd_session_sync": (C.c_int, [C.c_void_p]),
    "gfs_sgd_session_save": (C.c_int, [C.c_void_p]),
    "gfs_sgd_session_restore": (C.c_int, [C.c_void_p]),
    "gfs_sgd_session_positions": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), u64p, u32p]),
    "gfs_sgd_session_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "gfs_sgd_session_destroy": (None, [C.c_void_p]),
    "gfs_sort_positions": (C.c_int, [f64p, C.c_uint64, u32p]),
    "gfs_sgd_session_sort": (C.c_int, [C.c_void_p, u32p]),
    "gfs_sgd_sort_1d": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), f64p, u32p, C.POINTER(Stats)]),
    "gfs_edges_from_paths": (C.c_int, [u64p, u64p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_edge_list_get": (C.c_int, [C.c_void_p, C.POINTER(u64p), C.POINTER(u64p), u64p]),
    "gfs_edge_list_free": (None, [C.c_void_p]),
    "gfs_remap_handles": (C.c_int, [u64p, C.c_uint64, u64p, C.c_uint64, u8p, C.c_uint64]),
    "gfs_find_head_nodes": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p]),
    "gfs_groom_order": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p]),
    "gfs_topological_order": (C.c_int, [u8p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p]),
    "gfs_gfa_parse_file": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gfs_gfa_parse_text": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_gfa_dims": (C.c_int, [C.c_void_p, u64p, u64p, u64p, u64p, u64p]),
    "gfs_gfa_arrays": (C.c_int, [C.c_void_p, C.POINTER(u8p), C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p),
                                 C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p)]),
    "gfs_gfa_text": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p)]),
    "gfs_gfa_free": (None, [C.c_void_p]),
    "gfs_layout_write_tsv": (C.c_int, [f64p, C.c_uint64, C.c_uint32, C.c_char_p, u64p]),
    "gfs_gfa_write": (C.c_int, [C.c_char_p, u8p, C.c_uint64, C.c_char_p, u64p, u64p, u64p, u64p, C.c_uint64, u64p, u64p,
                                C.c_uint64, C.c_char_p, u64p, u64p, u64p]),
    "gfs_reconcile_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gfs_reconcile_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gfs_synth_create": (C.c_int, [C.POINTER(SynthSpec), C.POINTER(C.c_void_p)]),
    "gfs_synth_create_range": (C.c_int, [C.POINTER(SynthSpec), C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gfs_synth_path_counts": (C.c_int, [C.POINTER(SynthSpec), u64p]),
    "gfs_synth_dims": (C.c_int, [C.c_void_p, u64p, u64p, u64p]),
    "gfs_synth_arrays": (C.c_int, [C.c_void_p, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u32p)]),
    "gfs_synth_free": (None, [C.c_void_p]),
    "gfs_debug_fast_precise_pow": (C.c_int, [f64p, f64p, f64p, C.c_uint64]),
    "gfs_debug_dirty_zipf": (C.c_int, [u64p, f64p, f64p, f64p, u64p, C.c_uint64]),
    "gfs_debug_philox": (C.c_int, [u32p, u32p, u32p, C.c_uint64]),
    "gfs_debug_trace_terms": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.c_int32, C.c_uint64, C.c_uint32,
                                        C.c_uint64, C.c_uint64, u8p, u64p, u64p, u8p, f64p]),
    "gfs_p2p_region_create": (C.c_int, [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "gfs_p2p_region_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p]),
    "gfs_p2p_region_ipc_handle": (C.c_int, [C.c_void_p, u8p]),
    "gfs_p2p_region_connect_ipc": (C.c_int, [C.c_void_p, u8p, C.c_uint32, C.c_uint32]),
    "gfs_p2p_region_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "gfs_p2p_region_snapshot": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gfs_p2p_reconcile": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gfs_p2p_reconcile_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p]),
    "gfs_p2p_region_snap_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gfs_p2p_region_snapshot_x": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gfs_p2p_reconcile_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gfs_p2p_reconcile_async_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p]),
    "gfs_p2p_region_check": (C.c_int, [C.c_void_p]),
    "gfs_p2p_region_free": (None, [C.c_void_p]),
    "gfs_shard_plan_make": (C.c_int, [u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(ShardPlan)]),
    "gfs_default_syncs_per_epoch": (C.c_uint32, [C.c_uint64, C.c_uint64]),
    "gfs_shard_epoch_quota": (C.c_uint64, [C.c_uint64, C.POINTER(ShardPlan), C.c_uint64]),
    "gfs_replica_create": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), C.c_uint32, C.POINTER(LaunchCfg), C.POINTER(ShardPlan),
                                     C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "gfs_replica_ipc_handle": (C.c_int, [C.c_void_p, u8p]),
    "gfs_replica_connect_ipc": (C.c_int, [C.c_void_p, u8p, C.c_uint32, C.c_uint32]),
    "gfs_replica_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "gfs_replica_upload": (C.c_int, [C.c_void_p, f64p]),
    "gfs_replica_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "gfs_replica_flush": (C.c_int, [C.c_void_p]),
    "gfs_replica_sync": (C.c_int, [C.c_void_p]),
    "gfs_replica_download": (C.c_int, [C.c_void_p, f64p]),
    "gfs_replica_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "gfs_replica_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p, u32p]),
    "gfs_replica_destroy": (None, [C.c_void_p]),
    "gfs_debug_schedule": (C.c_int, [C.POINTER(SgdParams), f64p]),
    "gfs_debug_zetas": (C.c_int, [C.c_void_p, C.POINTER(SgdParams), f64p, C.c_uint64, u64p]),
    "gfs_debug_zetas_host": (C.c_int, [C.POINTER(SgdParams), C.c_uint64, f64p, C.c_uint64, u64p]),
}

_lib = None


def lib():
    """Load libgfasort_cuda.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with gfasort_b200/csrc/build.sh "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). gfasort_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, allow=()):
    if rc != GFS_OK and rc not in allow:
        raise GfsError(rc, lib().gfs_last_error().decode(errors="replace"))
    return rc
