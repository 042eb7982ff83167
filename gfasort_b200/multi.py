"""Multi-GPU `Y` / `L`: replicated positions, sharded term sampling.

One process per GPU (torch.distributed, NCCL over NVLink).  Terms shard naturally — every term
lives inside one path (reference src/sgd.rs:445, 502-503) — positions do not.  So:

  * rank r samples only the steps of its slice [S*r/G, S*(r+1)/G) of the concatenated step array
    and holds the records of just the paths that slice overlaps (its partners never leave them);
  * every rank keeps a full replica of the positions and runs its share
    min_term_updates * |slice| / S of every epoch, which keeps the global sampling distribution
    uniform over steps (src/sgd.rs:435, 444);
  * replicas are reconciled `syncs_per_epoch` times per epoch by an all-reduce over the position
    array: "avg" (north star: mean of the replicas), "tavg" (mean over the replicas that moved the
    node since the last sync) or "delta" (sum of the replicas' displacements, i.e. Hogwild with staleness);
    "p2p" is "tavg" done by ONE kernel per rank over NVLink peer memory instead of pack + NCCL all-reduce +
    apply (gfs_p2p_*; opt-in until measured).

Everything here is host logic over torch tensors; it runs unchanged on CPU tensors with the gloo
backend, which is how tests/test_multi_gloo.py covers it without GPUs.
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
    """Step-balanced slice for `rank` and the path range covering it."""
    path_first = np.asarray(path_first, dtype=np.uint64)
    S = int(path_first[-1])
    b = (S * rank) // world
    e = (S * (rank + 1)) // world
    if e <= b:
        return Shard(rank, world, b, b, 0, 0, 0)
    pb = int(np.searchsorted(path_first, b, side="right") - 1)
    pe = int(np.searchsorted(path_first, e - 1, side="right"))
    return Shard(rank, world, b, e, pb, pe, int(path_first[pb]))


def epoch_quota(min_term_updates: int, shard: Shard, total_steps: int) -> int:
    """This rank's share of one epoch: floor/ceil split of M in proportion to the slice, exact in sum."""
    lo = (min_term_updates * shard.sample_begin) // total_steps
    hi = (min_term_updates * shard.sample_end) // total_steps
    return hi - lo


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
        if self.n and (self.x.data_ptr() != self.x_ptr or self.x_sync.data_ptr() != self.x_sync_ptr):
            raise RuntimeError("PeerRegion: torch copied the region instead of viewing it")

    def ipc_handle(self) -> bytes:
        import ctypes as C

        from ._cabi import check, lib, u8p
        buf = (C.c_uint8 * 64)()
        check(lib().gfs_p2p_region_ipc_handle(self._h, C.cast(buf, u8p)))
        return bytes(buf)

    def connect_ipc(self, handles: list, rank: int) -> None:
        """handles: the 64-byte blobs of all ranks in rank order (one process per GPU)."""
        import ctypes as C

        from ._cabi import check, lib, u8p
        assert all(len(h) == 64 for h in handles)
        blob = (C.c_uint8 * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        check(lib().gfs_p2p_region_connect_ipc(self._h, C.cast(blob, u8p), len(handles), rank))

    @staticmethod
    def connect_local(regions: list) -> None:
        """All replicas live in this process (several GPUs with peer access, or several replicas on one GPU)."""
        import ctypes as C

        from ._cabi import check, lib
        arr = (C.c_void_p * len(regions))(*[r._h for r in regions])
        check(lib().gfs_p2p_region_connect_local(arr, len(regions)))

    def reconcile(self, stream_ptr: int) -> None:
        from ._cabi import check, lib
        check(lib().gfs_p2p_reconcile(self._h, stream_ptr))

    def check(self) -> None:
        from ._cabi import check, lib
        check(lib().gfs_p2p_region_check(self._h))

    def close(self) -> None:
        from ._cabi import lib
        if self._h:
            self.x = self.x_sync = None
            lib().gfs_p2p_region_free(self._h)
            self._h = None


def connect_peer_regions(region: PeerRegion, rank: int, world: int, group=None) -> None:
    """One process per GPU: all-gather the IPC handles over torch.distributed and map the peers."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(region.ipc_handle()), dtype=torch.uint8, device=f"cuda:{region.device}")
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    region.connect_ipc([bytes(t.cpu().tolist()) for t in allh], rank)
    dist.barrier(group=group)            # nobody reconciles before everybody has mapped everybody


# ------------------------------------------------------------------------------------------------
# one rank of a replicated run (GPU only: drives the C-ABI session API)
# ------------------------------------------------------------------------------------------------
_DS = {0: 1, 1: 1, 2: 2, 3: 4, 4: 4, 5: 8, 6: 8, 7: 8, 8: 8}     # coordinate stride per node end (gfs_lib.cu pick_nd)


class ReplicaRun:
    """This rank's share of a `Y` (dims = 0) or `L` (dims >= 1) run.

    index: the PathIndex of paths [shard.path_begin, shard.path_end) only (build_shard_index) — a
    rank never needs the other ranks' records.  The positions live in a torch tensor (so
    torch.distributed can all-reduce them in place) that the library's session uses as its position
    buffer; the session launches on `self.stream`.
    """

    def __init__(self, index, n_nodes: int, shard: Shard, total_steps: int, params, dims: int = 0,
                 device: int = 0, syncs_per_epoch: int = 1, mode: str = "avg", group=None,
                 layout_f64: bool = False):
        import ctypes as C

        import torch

        from ._cabi import LaunchCfg, check, lib
        self.shard, self.mode, self.group, self.syncs = shard, mode, group, max(1, int(syncs_per_epoch))
        self.dims, self.device = dims, device
        self.index = index
        self.N = int(n_nodes)
        f64 = dims == 0 or layout_f64
        n_elems = self.N if dims == 0 else self.N * 2 * _DS[dims]
        self.region = None
        if mode == "p2p":
            import torch.distributed as dist
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.region = PeerRegion(device, n_elems, f64)
            if world > 1:
                connect_peer_regions(self.region, shard.rank, world, group)
            else:
                PeerRegion.connect_local([self.region])
            self.x, self.x_sync, self.scratch = self.region.x, self.region.x_sync, None
        else:
            self.x = torch.zeros(n_elems, dtype=torch.float64 if f64 else torch.float32, device=f"cuda:{device}")
            self.x_sync = torch.empty_like(self.x) if mode in ("delta", "tavg") else None
            self.scratch = torch.empty(2 * n_elems, dtype=torch.float32, device=self.x.device) if mode == "tavg" else None
        self.stream = torch.cuda.Stream(device=device)
        from dataclasses import replace
        self.params = replace(params, min_term_updates=epoch_quota(params.min_term_updates, shard, total_steps))
        self.global_updates_per_epoch = params.min_term_updates
        cfg = LaunchCfg.default()
        cfg.device = device
        cfg.layout_f64 = int(layout_f64)
        cfg.rng_thread_base = shard.rank << 24
        cfg.stream = self.stream.cuda_stream
        cfg.device_positions = self.x.data_ptr()
        cfg.sample_begin = shard.sample_begin - shard.first_step_of_path_begin    # index-local step range
        cfg.sample_end = cfg.sample_begin + shard.steps
        self._h = C.c_void_p()
        cp = self.params.c()
        check(lib().gfs_sgd_session_create(self.index.handle, C.byref(cp), dims, C.byref(cfg), C.byref(self._h)))
        self.n_epochs = params.iter_max + 1
        torch.cuda.synchronize(device)        # allocations / zero fills on torch's stream are done before the session's stream runs

    def upload(self, positions):
        """Host positions (f64, the caller's node order / Layout order) -> this rank's replica."""
        import numpy as np

        from ._cabi import check, f64p, lib
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        import torch
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
        check(lib().gfs_sgd_session_download(self._h, out.ctypes.data_as(f64p)))
        return out

    def run_epoch(self, epoch: int):
        """One epoch of the schedule: `syncs` slices of this rank's quota, replicas reconciled after each."""
        import torch

        from ._cabi import check, lib
        with torch.cuda.stream(self.stream):
            for k in range(self.syncs):
                check(lib().gfs_sgd_session_run(self._h, epoch, epoch + 1, k, self.syncs))
                if self.region is not None:
                    self.region.reconcile(self.stream.cuda_stream)      # one kernel: barrier, reduce + scatter over NVLink, barrier
                else:
                    reconcile(self.x, self.x_sync, self.mode, self.group, self.scratch)

    def stats(self) -> dict:
        import ctypes as C

        from ._cabi import Stats, check, lib
        st = Stats()
        check(lib().gfs_sgd_session_stats(self._h, C.byref(st)))
        if self.region is not None:
            self.region.check()          # a timed-out peer barrier is an error, never a silent skip
        return st.as_dict()

    def close(self):
        from ._cabi import lib
        if self._h:
            lib().gfs_sgd_session_destroy(self._h)
            self._h = None
        if self.region is not None:
            self.x = self.x_sync = None
            self.region.close()
            self.region = None


def build_shard_index(shard_handles, shard_first, node_len, device: int = 0, rank: int = 0, world: int = 1, group=None):
    """PathIndex of this rank's paths, with ONE node relabelling for all ranks: rank 0 derives the
    first-appearance order from its own records and broadcasts it, so that the position replicas
    line up element-wise for the all-reduce.  shard_first is local to the shard (starts at 0)."""
    import os

    import numpy as np

    from .sgd import PathIndex
    assert int(shard_first[0]) == 0, "shard_first must be local to the shard (start at 0)"
    if world == 1:
        return PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device)
    import torch
    import torch.distributed as dist
    relabel = os.environ.get("GFASORT_RELABEL", "1") != "0"
    if not relabel:
        return PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device, relabel=0)
    perm = torch.empty(len(node_len), dtype=torch.int32, device=f"cuda:{device}")
    ix = None
    if rank == 0:
        ix = PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device, relabel=1)
        perm.copy_(torch.from_numpy(ix.relabel_permutation().view(np.int32)))
    dist.broadcast(perm, src=0, group=group)
    if rank != 0:
        ix = PathIndex.from_arrays(shard_handles, shard_first, node_len, device=device,
                                   new_of_old=perm.cpu().numpy().view(np.uint32))
    return ix
