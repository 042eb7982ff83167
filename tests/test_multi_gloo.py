"""CPU tests (-m "not gpu") of the multi-GPU host logic (gfasort_b200/multi.py) with world_size 2 on
the gloo backend: step-balanced sharding, exact epoch quotas, and the two replica reconcile rules.
The GPU side of the same path (ReplicaRun over NCCL) is exercised by bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gfasort_b200.multi import Shard, epoch_quota, reconcile, shard_steps


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_steps_partition():
    first = np.array([0, 10, 10, 45, 100, 101], dtype=np.uint64)       # 5 paths, one empty
    S = 101
    for world in (1, 2, 3, 4, 8):
        shards = [shard_steps(first, r, world) for r in range(world)]
        assert shards[0].sample_begin == 0 and shards[-1].sample_end == S
        for a, b in zip(shards, shards[1:]):
            assert a.sample_end == b.sample_begin                       # disjoint, covering
        assert max(s.steps for s in shards) - min(s.steps for s in shards) <= 1   # step-balanced
        for s in shards:
            # the path range covers the slice, so partners never leave the rank's records
            assert int(first[s.path_begin]) <= s.sample_begin and s.sample_end <= int(first[s.path_end])
            assert s.first_step_of_path_begin == int(first[s.path_begin])
        M = 1_000_003
        assert sum(epoch_quota(M, s, S) for s in shards) == M           # quotas are exact in sum
        for s in shards:
            assert abs(epoch_quota(M, s, S) - M * s.steps / S) <= 1     # proportional to the slice


def test_shard_more_ranks_than_steps():
    first = np.array([0, 3], dtype=np.uint64)
    shards = [shard_steps(first, r, 8) for r in range(8)]
    assert sum(s.steps for s in shards) == 3
    assert sum(epoch_quota(100, s, 3) for s in shards) == 100


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        x_sync = torch.randn(1000, dtype=torch.float64, generator=g)     # same on every rank
        deltas = [torch.randn(1000, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        # --- avg: mean of the replicas
        x = x_sync + deltas[rank]
        reconcile(x, None, "avg")
        want = x_sync + sum(deltas) / world
        assert torch.allclose(x, want, atol=1e-12), "avg"
        # --- delta: x_sync + sum of displacements, x_sync refreshed
        xs = x_sync.clone()
        x = xs + deltas[rank]
        reconcile(x, xs, "delta")
        want = x_sync + sum(deltas)
        assert torch.allclose(x, want, atol=1e-12), "delta"
        assert torch.equal(xs, x), "x_sync refreshed"
        # --- tavg: mean over the replicas that moved the element; untouched elements stay put
        xs = x_sync.clone()
        moved = [torch.rand(1000, generator=torch.Generator().manual_seed(200 + r)) < 0.5 for r in range(world)]
        x = xs + deltas[rank] * moved[rank]
        reconcile(x, xs, "tavg")
        cnt = sum(m.double() for m in moved).clamp(min=1)
        want = x_sync + sum(d * m for d, m in zip(deltas, moved)) / cnt
        assert torch.allclose(x, want, atol=1e-12), "tavg"
        none = ~(moved[0] | moved[1])
        assert torch.equal(x[none], x_sync[none])
        # replicas are identical after a reconcile (bitwise: same reduction on every rank)
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        # float32 replicas (nD layout) go through the same path
        xf = (x_sync + deltas[rank]).float()
        reconcile(xf, None, "avg")
        assert torch.allclose(xf.double(), x_sync + sum(deltas) / world, atol=1e-5)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_reconcile_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_reconcile_world1_is_identity():
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        x = torch.arange(5, dtype=torch.float64)
        xs = torch.zeros(5, dtype=torch.float64)
        reconcile(x, xs, "delta")
        assert torch.equal(x, torch.arange(5, dtype=torch.float64)) and torch.equal(xs, x)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
