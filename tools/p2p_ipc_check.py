"""Peer-memory reconcile between PROCESSES (CUDA IPC over NVLink): correctness against the moved-replica
mean and the time of one reconcile.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/p2p_ipc_check.py
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gfasort_b200.multi import PeerRegion

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)


def connect(region):
    handles = [None] * world
    dist.all_gather_object(handles, region.ipc_handle())
    region.connect_ipc(handles, rank)
    dist.barrier()


def expected(xs, xr):
    d = np.stack([x - xs for x in xr]); moved = np.stack([x != xs for x in xr]); cnt = moved.sum(0)
    out = xs + (d * moved).sum(0) / np.maximum(cnt, 1)
    one = cnt == 1
    out[one] = np.stack(xr)[moved.argmax(0)[one], np.nonzero(one)[0]]
    out[cnt == 0] = xs[cnt == 0]
    return out


stream = torch.cuda.Stream()
# ---- correctness, 3 consecutive reconciles -------------------------------------------------------
n = 1_000_003
reg = PeerRegion(rank, n, True)
connect(reg)
rng = np.random.default_rng(11)                      # same stream on every rank
xs = rng.standard_normal(n) * 1e6
reg.x_sync.copy_(torch.from_numpy(xs)); reg.x.copy_(torch.from_numpy(xs))
ok = True
for rnd in range(3):
    xr = []
    for g in range(world):
        mask = rng.random(n) < 0.5
        x = xs.copy(); x[mask] += rng.standard_normal(int(mask.sum())) * 100
        xr.append(x)
    reg.x.copy_(torch.from_numpy(xr[rank]))
    torch.cuda.synchronize(); dist.barrier()
    reg.reconcile(stream.cuda_stream)
    torch.cuda.synchronize(); reg.check()
    got = reg.x.cpu().numpy()
    want = expected(xs, xr)
    ok = ok and np.allclose(got, want, rtol=0, atol=1e-9) and np.array_equal(reg.x_sync.cpu().numpy(), got)
    xs = got
    dist.barrier()
flags = [None] * world
dist.all_gather_object(flags, bool(ok))
reg.close()
# ---- time of one reconcile at config 3's size (10M f64 = 80 MB per replica) ------------------------
n = 10_000_000
reg = PeerRegion(rank, n, True)
connect(reg)
reg.x_sync.zero_(); reg.x.copy_(torch.arange(n, dtype=torch.float64, device=f"cuda:{rank}") * (rank + 1) % 7)
torch.cuda.synchronize(); dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(3):
    reg.reconcile(stream.cuda_stream)
torch.cuda.synchronize(); dist.barrier()
K = 20
with torch.cuda.stream(stream):
    ev[0].record()
for it in range(K):
    reg.x.add_(1.0 + rank)                           # every element moves on every replica (on the default stream: keep it ordered)
    torch.cuda.synchronize()
    reg.reconcile(stream.cuda_stream)
    torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(K):
    reg.reconcile(stream.cuda_stream)                # nothing moved: same traffic, no host work in between
with torch.cuda.stream(stream):
    ev[1].record()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
reg.check()
if rank == 0:
    print(f"p2p_ipc_check: world {world}, correctness {'OK' if all(flags) else 'FAILED'} on {flags}; "
          f"one reconcile of {n} f64 ({n*8/1e6:.0f} MB per replica): {dt*1e3:.3f} ms back-to-back", flush=True)
reg.close()
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
