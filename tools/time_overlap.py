#!/usr/bin/env python
"""Under torchrun: per-step time of the sharded R1M search step in its three exchange schemes, every one replayed from
CUDA graphs over 8 rotating query batches: local step only (no exchange) / exchange + merge inside the graph, deferred
by a step / local-step graph + exchange kernel on a second stream (overlapped with the next scan)."""
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
N, D, NQ, K, NB = 1001001, 2048, 70, 100, 8
lo, hi = ShardedIndex.shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(1 + rank)
db = torch.randn((hi - lo, D), device=dev, generator=g)
db = db / db.norm(dim=1, keepdim=True)
index = Index.from_packed(pack_bf16(db), db32=db, idx_base=lo)
sharded = ShardedIndex.from_local(index)
qs = torch.randn((NB, NQ, D), device=dev, generator=torch.Generator(device=dev).manual_seed(7))
qs = qs / qs.norm(dim=2, keepdim=True)


def timed(run, n=400):
    run(24)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def graphs(**kw):
    out = []
    for b in range(NB):
        gb = GraphedSearch(kw.pop("target", sharded) if False else kw.get("target", sharded), NQ, K,
                           **{k_: v for k_, v in kw.items() if k_ != "target"})
        gb.q.copy_(qs[b])
        out.append(gb)
    return out


res = {}
gl = graphs(target=index)
res["local step only"] = timed(lambda n: [gl[t % NB].graph.replay() for t in range(n)])
gd = graphs(deferred=True)


def run_deferred(n):
    for t in range(n):
        gd[t % NB].graph.replay()
    gd[(n - 1) % NB].drain()


res["exchange inside the graph, merge deferred by a step"] = timed(run_deferred)
go = graphs(overlap=True)
exch = torch.cuda.Stream(device=dev)


def run_overlap(n, guard=True):
    cur = torch.cuda.current_stream(dev)
    for t in range(n):
        gb = go[t % NB]
        if guard and t >= NB:
            cur.wait_event(gb.done)
        gb.graph.replay()
        gb.local_done.record(cur)
        with torch.cuda.stream(exch):
            exch.wait_event(gb.local_done)
            gb.exchange()
            gb.done.record(exch)
    cur.wait_stream(exch)


res["local-step graph + exchange on a second stream"] = timed(run_overlap)
go = graphs(overlap=True, split=True)
assert all(g_.split for g_ in go)
res["scan graph; finalize graph + exchange on a second stream"] = timed(run_overlap)
for scan_ctas, fin_chunk, fin_stage in [tuple(int(v) for v in x.split(":")) for x in os.environ.get("MDIR_SM_SPLITS", "").split(",") if x]:
    del go
    go = graphs(overlap=True, split=True, scan_ctas=scan_ctas, fin_chunk=fin_chunk, fin_stage=fin_stage)
    res["same, scan <= %d CTAs, finalize %d q/launch on single CTAs staging %d keys" % (scan_ctas, fin_chunk, fin_stage)] = timed(run_overlap)
    flagged = sum(int(g_.status.ne(0).sum().item()) for g_ in go)
    if flagged:
        res["   (flagged queries in the last replays: %d)" % flagged] = 0.0
for scan_ctas, fin_chunk, fin_stage in [tuple(int(v) for v in x.split(":")) for x in os.environ.get("MDIR_SM_SPLITS", "").split(",") if x]:
    del go
    go = graphs(overlap=True, split=True, scan_ctas=scan_ctas, fin_chunk=fin_chunk, fin_stage=fin_stage)
    res["same, scan <= %d CTAs, finalize %d q/launch on single CTAs staging %d keys" % (scan_ctas, fin_chunk, fin_stage)] = timed(run_overlap)
    flagged = sum(int(g_.status.ne(0).sum().item()) for g_ in go)
    if flagged:
        res["   (flagged queries in the last replays: %d)" % flagged] = 0.0
if rank == 0:
    for k_, v in res.items():
        print("world %d, %d rows/rank: %-55s %.1f us per step (max over ranks)" % (world, hi - lo, k_, v))
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
