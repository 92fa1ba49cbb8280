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
