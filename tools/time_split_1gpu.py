#!/usr/bin/env python
"""One GPU, R1M shape: per-step time of the search step as one CUDA graph vs split graphs (pack + scan on the compute
stream, finalize + certified re-score on a side stream beside the next scan) for several caps of the scan grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import GraphedSearch, Index, pack_bf16  # noqa: E402

dev = torch.device("cuda", 0)
N, D, NQ, K, NB = int(sys.argv[1]) if len(sys.argv) > 1 else 1001001, 2048, 70, 100, 8
g = torch.Generator(device=dev).manual_seed(1)
db = torch.randn((N, D), device=dev, generator=g)
db = db / db.norm(dim=1, keepdim=True)
index = Index.from_packed(pack_bf16(db), db32=db)
qs = torch.randn((NB, NQ, D), device=dev, generator=g)
qs = qs / qs.norm(dim=2, keepdim=True)
side = torch.cuda.Stream(device=dev)


def timed(run, n=200):
    run(16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def graphs(**kw):
    out = []
    for b in range(NB):
        gb = GraphedSearch(index, NQ, K, **kw)
        gb.q.copy_(qs[b])
        out.append(gb)
    return out


gl = graphs()
print("one graph per step: %.1f us" % timed(lambda n: [gl[t % NB].graph.replay() for t in range(n)]))
ref = [(gl[b]()[0].clone(), gl[b]()[1].clone()) for b in range(NB)]
for cap in (0, 144, 140, 136, 132, 124):
    go = graphs(overlap=True, split=True, scan_ctas=cap)

    def run(n):
        cur = torch.cuda.current_stream(dev)
        for t in range(n):
            gb = go[t % NB]
            if t >= NB:
                cur.wait_event(gb.done)
            gb.graph.replay()
            gb.local_done.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(gb.local_done)
                gb.exchange()
                gb.done.record(side)
        cur.wait_stream(side)

    us = timed(run)
    same = all(torch.equal(go[b].out[1], ref[b][1]) and torch.equal(go[b].out[0], ref[b][0]) for b in range(NB))
    print("split graphs, scan grid cap %3d: %.1f us per step, results == one-graph results: %s" % (cap, us, same))
    del go
