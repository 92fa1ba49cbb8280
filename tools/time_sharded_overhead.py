#!/usr/bin/env python
"""Under torchrun: per-step time of the local fused top-k alone vs the sharded step (local top-k +
all-gather of keys + merge), both replayed from CUDA graphs, on the bench shape split over the ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mdir_b200.search import GraphedSearch, Index, ShardedIndex, pack_bf16  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, D, NQ, K = 1001001, 2048, 70, 100
lo, hi = ShardedIndex.shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(1 + rank)
db = torch.randn((hi - lo, D), device=dev, generator=g)
db = db / db.norm(dim=1, keepdim=True)
index = Index.from_packed(pack_bf16(db), db32=db, idx_base=lo)
q = torch.randn((NQ, D), device=dev, generator=torch.Generator(device=dev).manual_seed(7))
q = q / q.norm(dim=1, keepdim=True)


def timeit(gs, n=300):
    gs.q.copy_(q)
    for _ in range(10):
        gs()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        gs()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


# the exchange + merge kernel alone, back to back (its intrinsic cost, without the per-step skew between ranks)
import ctypes as C  # noqa: E402
from mdir_b200 import _lib  # noqa: E402
sh0 = ShardedIndex.from_local(index)
_, _, keys0 = index.search(q, K, return_keys=True)
o_s = torch.empty((NQ, K), dtype=torch.float32, device=dev)
o_i = torch.empty((NQ, K), dtype=torch.int32, device=dev)


def exch():
    _lib.check(_lib.lib().mdir_shard_exchange_merge(_lib.ptr(keys0), NQ, K, rank, world, sh0.P2P_MAX_Q, sh0.P2P_MAX_K, 0, C.cast(sh0._mb, C.c_void_p),
                                                    _lib.ptr(o_s), _lib.ptr(o_i), None, None, _lib.stream()), "x")


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    exch()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
gx = torch.cuda.CUDAGraph()
with torch.cuda.graph(gx):
    for _ in range(20):
        exch()
dist.barrier()
torch.cuda.synchronize()
evx = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
gx.replay()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
evx[0].record()
for _ in range(10):
    gx.replay()
evx[1].record()
torch.cuda.synchronize()
if rank == 0:
    print("exchange+merge kernel alone: %.1f us per call (20 per graph, 10 replays)" % (evx[0].elapsed_time(evx[1]) / 200 * 1e3), flush=True)

t_local = timeit(GraphedSearch(index, NQ, K))
t_local16 = timeit(GraphedSearch(index, NQ, K, precision="bf16"))
t_shard = timeit(GraphedSearch(ShardedIndex.from_local(index), NQ, K))
ShardedIndex.p2p = False
t_nccl = timeit(GraphedSearch(ShardedIndex.from_local(index), NQ, K))
res = torch.tensor([t_local, t_local16, t_shard, t_nccl], device=dev, dtype=torch.float64)
allr = [torch.empty_like(res) for _ in range(world)]
dist.all_gather(allr, res)
if rank == 0:
    m = torch.stack(allr).max(0).values.tolist()
    print("world %d rows/rank %d: local fp32 step %.1f us, local bf16 step %.1f us; sharded step %.1f us with the fused NVLink exchange+merge "
          "(+%.1f us), %.1f us with ncclAllGather + merge (+%.1f us)"
          % (world, hi - lo, m[0] * 1e3, m[1] * 1e3, m[2] * 1e3, (m[2] - m[0]) * 1e3, m[3] * 1e3, (m[3] - m[0]) * 1e3), flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
