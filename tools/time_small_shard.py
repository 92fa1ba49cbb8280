#!/usr/bin/env python
"""One GPU, the per-rank work of the 1/2/4/8-GPU search step (no exchange): CUDA-graph replay time of the certified
top-100 search over a shard of 1,001,001 / N rows, and the share of the scan kernel.  python tools/time_small_shard.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index, GraphedSearch, pack_bf16  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
D, NQ, K = 2048, 70, 100
q = torch.randn((NQ, D), device=dev, generator=g)
q /= q.norm(dim=1, keepdim=True)


class Prof:
    def __init__(self):
        self.e0 = torch.cuda.Event(enable_timing=True, external=True)
        self.e1 = torch.cuda.Event(enable_timing=True, external=True)

    def begin(self):
        self.e0.record()

    def end(self, nbytes):
        self.e1.record()
        self.bytes = nbytes


for world in (1, 2, 4, 8):
    n = -(-1001001 // world)
    db = torch.randn((n, D), device=dev, generator=g)
    db /= db.norm(dim=1, keepdim=True)
    idx = Index.from_packed(pack_bf16(db), db32=db)
    gs = GraphedSearch(idx, NQ, K)
    gs.q.copy_(q)
    for _ in range(5):
        gs()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gs.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    prof = Prof()
    gp = GraphedSearch(idx, NQ, K, prof=prof)
    gp.q.copy_(q)
    sc = []
    for _ in range(10):
        gp()
        torch.cuda.synchronize()
        sc.append(prof.e0.elapsed_time(prof.e1))
    scan = sum(sc) / len(sc)
    print("shard 1/%d (%d rows): step %.1f us, scan %.1f us (%.2f TB/s), rest %.1f us, flagged=%s" %
          (world, n, ms * 1e3, scan * 1e3, n * D * 2 / scan / 1e9, (ms - scan) * 1e3, gs.check_overflow()))
    del gs, gp, idx, db
