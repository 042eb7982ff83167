"""Multi-GPU `Y` / `L`: replicated positions, sharded term sampling.

The run itself lives behind the C ABI (gfasort_b200/csrc/gfs_multi.cu): `gfs_shard_plan_make` says which step
slice a rank samples and which paths' records it needs, `gfs_replica_*` is one rank (session + peer-memory
region + the epoch loop with its reconciles), and `gfs_sgd_1d` / `gfs_sgd_nd` on an index built under
GFASORT_GPUS=G drive all G GPUs from one process.  This module is the thin host side of the
one-process-per-GPU form: torch.distributed carries the 80-byte region handles, the shared node order and the
timing / stress reductions — plumbing, no data-path collective.

  * rank r samples only the steps of its slice [S*r/G, S*(r+1)/G) of the concatenated step array
    and holds the records of just the paths that slice overlaps (its partners never leave them);
  * every rank keeps a full replica of the positions and runs its share
    min_term_updates * |slice| / S of every epoch, which keeps the global sampling distribution
    uniform over steps (src/sgd.rs:435, 444);
  * replicas are reconciled `syncs_per_epoch` times per epoch.  "p2p" (default): ONE kernel per rank over
    NVLink peer memory forms the mean of the displacements over the replicas that moved the element
    (gfs_p2p.cu).  For comparison the same rule as pack + NCCL all-reduce + apply ("tavg"), the plain replica
    mean the north star names ("avg") and the displacement sum ("delta") remain, driven from here through
    torch.distributed.

`reconcile` runs unchanged on CPU tensors with the gloo backend, which is how tests/test_multi_gloo.py covers
the host logic without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world: int
    sample_begin: int      # global step range this rank samples from
    sample_end: int
    path_begin: int        # paths whose records this rank needs (those the slice overlaps)
    path_end: int
    first_step_of_path_begin: int = 0   # global step index of the first step of path_begin

    @property
    def steps(self) -> int:
        return self.sample_end - self.sample_begin


def shard_steps(path_first: np.ndarray, rank: int, world: int) -> Shard:
    """Step-balanced slice for `rank` and the path range covering it (gfs_shard_plan_make)."""
    import ctypes as C

    from ._cabi import ShardPlan, check, lib, u64p
    path_first = np.ascontiguousarray(path_first, dtype=np.uint64)
    pl = ShardPlan()
    check(lib().gfs_shard_plan_make(path_first.ctypes.data_as(u64p), len(path_first) - 1, rank, world, C.byref(pl)))
    return Shard(rank, world, pl.sample_begin, pl.sample_end, pl.path_begin, pl.path_end, pl.first_step)


def _plan(shard: Shard):
    from ._cabi import ShardPlan
    return ShardPlan(shard.sample_begin, shard.sample_end, shard.path_begin, shard.path_end, shard.first_step_of_path_begin)


def epoch_quota(min_term_updates: int, shard: Shard, total_steps: int) -> int:
    """This rank's share of one epoch: floor/ceil split of M in proportion to the slice, exact in sum
    (gfs_shard_epoch_quota)."""
    import ctypes as C

    from ._cabi import lib
    pl = _plan(shard)
    return int(lib().gfs_shard_epoch_quota(min_term_updates, C.byref(pl), total_steps))


def reconcile(x, x_sync, mode: str, group=None, scratch=None):
    """All-reduce the replicas in place.  x: this rank's positions (torch tensor).
    "avg": x <- mean over ranks.  "delta": x <- x_sync + sum over ranks of (x - x_sync); x_sync is
    then refreshed.  Returns x."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if x_sync is not None:
            x_sync.copy_(x)
        return x
    if mode == "avg":
        if x.is_cuda:
            dist.all_reduce(x, op=dist.ReduceOp.AVG, group=group)    # NCCL averages inside the collective
        else:
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)    # gloo has no AVG
            x.mul_(1.0 / world)
    elif mode == "delta":
        # sum_g x_g = G*x_sync + sum_g delta_g  =>  x_new = sum_g x_g - (G-1)*x_sync
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        x.add_(x_sync, alpha=-(world - 1))
    elif mode == "tavg":
        # mean of the displacements over the replicas that MOVED the element since the last sync:
        # x_new = x_sync + sum_g delta_g / max(1, #{g: delta_g != 0}).  Equal to "avg" where every replica
        # touched the node; a node only one rank's paths visit keeps its full displacement instead of 1/G of it.
        import torch
        if x.is_cuda:
            # the library's two fused kernels around one f32 all-reduce (displacements travel as f32: the
            # rounding is relative to the displacement, not to the position), on the current stream
            from ._cabi import check, lib
            n = x.numel()
            if scratch is None or scratch.numel() != 2 * n:
                scratch = torch.empty(2 * n, dtype=torch.float32, device=x.device)
            st = torch.cuda.current_stream(x.device).cuda_stream
            check(lib().gfs_reconcile_pack(x.data_ptr(), x_sync.data_ptr(), n, x.element_size(), scratch.data_ptr(), st))
            dist.all_reduce(scratch, op=dist.ReduceOp.SUM, group=group)
            check(lib().gfs_reconcile_apply(x.data_ptr(), x_sync.data_ptr(), n, x.element_size(), scratch.data_ptr(), st))
            return x                                   # apply refreshed x_sync already
        buf = torch.empty((2,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        torch.sub(x, x_sync, out=buf[0])
        buf[1].copy_(buf[0] != 0)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf[1].clamp_(min=1)
        torch.addcdiv(x_sync, buf[0], buf[1], out=x)
    else:
        raise ValueError(f"unknown reconcile mode {mode!r}")
    if x_sync is not None:
        x_sync.copy_(x)
    return x


# ------------------------------------------------------------------------------------------------
# peer-memory regions (mode "p2p")
# ------------------------------------------------------------------------------------------------
class _DeviceArray:
    """A raw device pointer as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerRegion:
    """One rank's replica + snapshot + barrier flags in one allocation that the other ranks map
    (gfs_p2p_region_*).  `x` / `x_sync` are torch views of the library's memory: valid until close()."""

    def __init__(self, device: int, n_elems: int, f64: bool, max_blocks: int = 0):
        import ctypes as C

        import torch

        from ._cabi import check, lib
        self._h = C.c_void_p()
        self.n, self.f64, self.device = int(n_elems), bool(f64), device
        check(lib().gfs_p2p_region_create(device, self.n, 8 if f64 else 4, max_blocks, C.byref(self._h)))
        px, pxs, nbytes = C.c_void_p(), C.c_void_p(), C.c_uint64()
        check(lib().gfs_p2p_region_ptrs(self._h, C.byref(px), C.byref(pxs), C.byref(nbytes)))
        self.x_ptr, self.x_sync_ptr, self.nbytes = px.value, pxs.value, nbytes.value
        ts = "<f8" if f64 else "<f4"
        dev = f"cuda:{device}"
        self.x = torch.as_tensor(_DeviceArray(self.x_ptr, self.n, ts), device=dev)
        self.x_sync = torch.as_tensor(_DeviceArray(self.x_sync_ptr, self.n, ts), device=dev)
        psn = C.c_void_p()
        check(lib().gfs_p2p_region_snap_ptr(self._h, C.byref(psn)))
        self.x_snap = torch.as_tensor(_DeviceArray(psn.value, self.n, ts), device=dev)     # the overlapped form's snapshot
        if self.n and (self.x.data_ptr() != self.x_ptr or self.x_sync.data_ptr() != self.x_sync_ptr):
            raise RuntimeError("PeerRegion: torch copied the region instead of viewing it")

    def ipc_handle(self) -> bytes:
        import ctypes as C

        from ._cabi import GFS_P2P_HANDLE_BYTES, check, lib, u8p
        buf = (C.c_uint8 * GFS_P2P_HANDLE_BYTES)()
        check(lib().gfs_p2p_region_ipc_handle(self._h, C.cast(buf, u8p)))
        return bytes(buf)

    def connect_ipc(self, handles: list, rank: int) -> None:
        """handles: the blobs of all ranks in rank order (one process per GPU)."""
        import ctypes as C

        from ._cabi import GFS_P2P_HANDLE_BYTES, check, lib, u8p
        assert all(len(h) == GFS_P2P_HANDLE_BYTES for h in handles)
        blob = (C.c_uint8 * (GFS_P2P_HANDLE_BYTES * len(handles))).from_buffer_copy(b"".join(handles))
        check(lib().gfs_p2p_region_connect_ipc(self._h, C.cast(blob, u8p), len(handles), rank))

    @staticmethod
    def connect_local(regions: list) -> None:
        """All replicas live in this process (several GPUs with peer access, or several replicas on one GPU)."""
        import ctypes as C

        from ._cabi import check, lib
        arr = (C.c_void_p * len(regions))(*[r._h for r in regions])
        check(lib().gfs_p2p_region_connect_local(arr, len(regions)))

    def reconcile(self, stream_ptr: int) -> None:
        """This rank's reconcile kernel (ranks on DIFFERENT devices)."""
        from ._cabi import check, lib
        check(lib().gfs_p2p_reconcile(self._h, stream_ptr))

    @staticmethod
    def reconcile_local(regions: list, stream_ptr: int) -> None:
        """All ranks of replicas that share ONE device, as one cooperative launch (tests)."""
        import ctypes as C

        from ._cabi import check, lib
        arr = (C.c_void_p * len(regions))(*[r._h for r in regions])
        check(lib().gfs_p2p_reconcile_local(arr, len(regions), stream_ptr))

    @staticmethod
    def reconcile_async_local(regions: list, stream_ptr: int) -> None:
        """The overlapped form's kernel (exchange over the snapshots, corrections added to the live replicas) for
        replicas that share ONE device, as one cooperative launch (tests)."""
        import ctypes as C

        from ._cabi import check, lib
        arr = (C.c_void_p * len(regions))(*[r._h for r in regions])
        check(lib().gfs_p2p_reconcile_async_local(arr, len(regions), stream_ptr))

    def check(self) -> None:
        from ._cabi import check, lib
        check(lib().gfs_p2p_region_check(self._h))

    def close(self) -> None:
        from ._cabi import lib
        if self._h:
            self.x = self.x_sync = self.x_snap = None
            lib().gfs_p2p_region_free(self._h)
            self._h = None


def _all_gather_bytes(blob: bytes, world: int, device: int, group=None) -> list:
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=f"cuda:{device}")
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    return [bytes(t.cpu().tolist()) for t in allh]


def connect_peer_regions(region: PeerRegion, rank: int, world: int, group=None) -> None:
    """One process per GPU: all-gather the IPC handles over torch.distributed and map the peers."""
    import torch.distributed as dist
    region.connect_ipc(_all_gather_bytes(region.ipc_handle(), world, region.device, group), rank)
    dist.barrier(group=group)            # nobody reconciles before everybody has mapped everybody


# ------------------------------------------------------------------------------------------------
# one rank of a replicated run (GPU only)
# ------------------------------------------------------------------------------------------------
_DS = {0: 1, 1: 1, 2: 2, 3: 4, 4: 4, 5: 8, 6: 8, 7: 8, 8: 8}     # coordinate stride per node end (gfs_internal.h coord_stride)


class ReplicaRun:
    """This rank's share of a `Y` (dims = 0) or `L` (dims >= 1) run.

    index: the PathIndex of paths [shard.path_begin, shard.path_end) only (build_shard_index) — a
    rank never needs the other ranks' records.

    mode "p2p" (default): the whole rank — session, peer-memory replica, epoch loop, reconciles — is one
    `gfs_replica` behind the C ABI; this class only exchanges the region handles.  The other modes keep the
    positions in a torch tensor that torch.distributed all-reduces between the library's SGD launches.
    """

    def __init__(self, index, n_nodes: int, shard: Shard, total_steps: int, params, dims: int = 0,
                 device: int = 0, syncs_per_epoch: int = 0, mode: str = "p2p", group=None,
                 layout_f64: bool = False):
        import ctypes as C

        import torch

        from ._cabi import LaunchCfg, check, lib
        # syncs_per_epoch = 0: the library's default, one reconcile per total_steps applied updates (1 for Y, 10 for L)
        syncs = int(syncs_per_epoch) if syncs_per_epoch else int(lib().gfs_default_syncs_per_epoch(params.min_term_updates, total_steps))
        self.shard, self.mode, self.group, self.syncs = shard, mode, group, max(1, syncs)
        self.dims, self.device = dims, device
        self.index = index
        self.N = int(n_nodes)
        self.n_epochs = params.iter_max + 1
        self.global_updates_per_epoch = params.min_term_updates
        self._h = self._rep = None
        f64 = dims == 0 or layout_f64
        n_elems = self.N if dims == 0 else self.N * 2 * _DS[dims]
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert world == shard.world, "the shard plan was made for another world size"
        cfg = LaunchCfg.default()
        cfg.device = device
        cfg.layout_f64 = int(layout_f64)
        if mode == "p2p":
            from ._cabi import u8p
            pl = _plan(shard)
            cp = params.c()
            self._rep = C.c_void_p()
            check(lib().gfs_replica_create(index.handle, C.byref(cp), dims, C.byref(cfg), C.byref(pl), total_steps,
                                           shard.rank, world, self.syncs, C.byref(self._rep)))
            if world > 1:
                from ._cabi import GFS_P2P_HANDLE_BYTES
                buf = (C.c_uint8 * GFS_P2P_HANDLE_BYTES)()
                check(lib().gfs_replica_ipc_handle(self._rep, C.cast(buf, u8p)))
                handles = _all_gather_bytes(bytes(buf), world, device, group)
                blob = (C.c_uint8 * (GFS_P2P_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
                check(lib().gfs_replica_connect_ipc(self._rep, C.cast(blob, u8p), world, shard.rank))
                dist.barrier(group=group)        # nobody reconciles before everybody has mapped everybody
            else:
                arr = (C.c_void_p * 1)(self._rep)
                check(lib().gfs_replica_connect_local(arr, 1))
            st, px, ne, eb = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint32()
            check(lib().gfs_replica_stream(self._rep, C.byref(st), C.byref(px), C.byref(ne), C.byref(eb)))
            self.stream = torch.cuda.ExternalStream(st.value, device=device)
            self.x = torch.as_tensor(_DeviceArray(px.value, ne.value, "<f8" if eb.value == 8 else "<f4"), device=f"cuda:{device}")
            self.x_sync = self.scratch = None
            return
        self.x = torch.zeros(n_elems, dtype=torch.float64 if f64 else torch.float32, device=f"cuda:{device}")
        self.x_sync = torch.empty_like(self.x) if mode in ("delta", "tavg") else None
        self.scratch = torch.empty(2 * n_elems, dtype=torch.float32, device=self.x.device) if mode == "tavg" else None
        self.stream = torch.cuda.Stream(device=device)
        from dataclasses import replace
        self.params = replace(params, min_term_updates=epoch_quota(params.min_term_updates, shard, total_steps))
        cfg.rng_thread_base = shard.rank << 24
        cfg.stream = self.stream.cuda_stream
        cfg.device_positions = self.x.data_ptr()
        cfg.sample_begin = shard.sample_begin - shard.first_step_of_path_begin    # index-local step range
        cfg.sample_end = cfg.sample_begin + shard.steps
        self._h = C.c_void_p()
        cp = self.params.c()
        check(lib().gfs_sgd_session_create(self.index.handle, C.byref(cp), dims, C.byref(cfg), C.byref(self._h)))
        torch.cuda.synchronize(device)        # allocations / zero fills on torch's stream are done before the session's stream runs

    def upload(self, positions):
        """Host positions (f64, the caller's node order / Layout order) -> this rank's replica."""
        import numpy as np

        from ._cabi import check, f64p, lib
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        import torch
        if self._rep is not None:
            check(lib().gfs_replica_upload(self._rep, positions.ctypes.data_as(f64p)))
            return
        check(lib().gfs_sgd_session_upload(self._h, positions.ctypes.data_as(f64p)))
        if self.x_sync is not None:
            # on the session's stream: torch's default stream does not order against it (non-blocking
            # streams), and a snapshot racing the first SGD launch would leave the ranks with different
            # x_sync — which "tavg" / "delta" never repair
            with torch.cuda.stream(self.stream):
                self.x_sync.copy_(self.x)
        torch.cuda.synchronize(self.device)

    def download(self):
        import numpy as np

        from ._cabi import check, f64p, lib
        ends, d = (1, 1) if self.dims == 0 else (2, self.dims)
        out = np.zeros(self.N * ends * d, dtype=np.float64)
        if self._rep is not None:
            check(lib().gfs_replica_download(self._rep, out.ctypes.data_as(f64p)))
        else:
            check(lib().gfs_sgd_session_download(self._h, out.ctypes.data_as(f64p)))
        return out

    def run_epoch(self, epoch: int):
        """One epoch of the schedule: `syncs` slices of this rank's quota, replicas reconciled after each.
        Asynchronous on self.stream."""
        import torch

        from ._cabi import check, lib
        if self._rep is not None:
            check(lib().gfs_replica_run(self._rep, epoch, epoch + 1))      # SGD slices + peer-memory reconciles, one C call
            return
        with torch.cuda.stream(self.stream):
            for k in range(self.syncs):
                check(lib().gfs_sgd_session_run(self._h, epoch, epoch + 1, k, self.syncs))
                if self.shard.world > 1:
                    reconcile(self.x, self.x_sync, self.mode, self.group, self.scratch)

    def flush(self):
        """Asynchronous: self.stream waits for the last overlapped reconcile (before a timing event on that stream)."""
        from ._cabi import check, lib
        if self._rep is not None:
            check(lib().gfs_replica_flush(self._rep))

    def stats(self) -> dict:
        import ctypes as C

        from ._cabi import Stats, check, lib
        st = Stats()
        if self._rep is not None:
            check(lib().gfs_replica_stats(self._rep, C.byref(st)))     # a timed-out peer barrier is an error here, never a silent skip
        else:
            check(lib().gfs_sgd_session_stats(self._h, C.byref(st)))
        return st.as_dict()

    def close(self):
        from ._cabi import lib
        self.x = self.x_sync = None
        if self._rep is not None:
            lib().gfs_replica_destroy(self._rep)
            self._rep = None
        if self._h:
            lib().gfs_sgd_session_destroy(self._h)
            self._h = None


def build_shard_index(shard_handles, shard_first, node_len, device: int = 0, rank: int = 0, world: int = 1, group=None):
    """PathIndex of this rank's paths, with ONE node relabelling for all ranks so that the position replicas
    line up element-wise: every rank runs K1 at the same time (rank 0 with the first-appearance relabelling,
    the others without), then rank 0's permutation is broadcast and the others adopt it
    (gfs_index_apply_relabel).  shard_first is local to the shard (starts at 0)."""
    import os

    import numpy as np

    from .sgd import PathIndex
    assert int(shard_first[0]) == 0, "shard_first must be local to the shard (start at 0)"
    if world == 1:
        return PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device)
    import torch
    import torch.distributed as dist
    relabel = os.environ.get("GFASORT_RELABEL", "1") != "0"
    ix = PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device, relabel=1 if (relabel and rank == 0) else 0)
    if not relabel:
        return ix
    perm = torch.empty(len(node_len), dtype=torch.int32, device=f"cuda:{device}")
    if rank == 0:
        perm.copy_(torch.from_numpy(ix.relabel_permutation().view(np.int32)))
    dist.broadcast(perm, src=0, group=group)
    if rank != 0:
        ix.apply_relabel(perm.cpu().numpy().view(np.uint32))
    return ix


def all_paths_stress(index, shard: Shard, total_steps: int, positions, dims: int, samples: int, layout_order: bool,
                     device: int = 0, group=None, seed: int = 12345):
    """Sampled path stress over ALL paths of a sharded graph: every rank evaluates the samples that land in its
    step slice (gfs_stress_partial), the three sums are all-reduced.  Same sample, same value as gfs_stress on
    one GPU holding the whole graph.  Returns (rms_rel, mean_abs_rel, counted)."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from ._cabi import check, f64p, lib
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    sums = (C.c_double * 3)(0.0, 0.0, 0.0)
    if shard.steps > 0:
        check(lib().gfs_stress_partial(index.handle, max(dims, 1), int(layout_order), positions.ctypes.data_as(f64p), samples, seed,
                                       total_steps, shard.first_step_of_path_begin, shard.sample_begin, shard.sample_end, sums))
    t = torch.tensor(list(sums), dtype=torch.float64, device=f"cuda:{device}")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    s0, s1, c = (float(v) for v in t.tolist())
    if c <= 0:
        return 0.0, 0.0, 0
    return float(np.sqrt(s0 / c)), s1 / c, int(c)
