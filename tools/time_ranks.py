#!/usr/bin/env python
"""Time Index.ranks() (scores + full per-query ranking) on the C1 / mining / C3-slice shapes, with the score pass timed
alone beside it; run under ncu for the launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ranks.csv python tools/time_ranks.py 1 c3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index, RANK_STATS  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
which = sys.argv[2:] or ["c1", "mining", "c3"]
SHAPES = {"c1": (4993, 2048, 70, "fp32"), "mining": (20000, 2048, 2000, "fp32"), "c3": (100000, 512, 1024, "bf16")}
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for tag in which:
    n_db, D, nq, prec = SHAPES[tag]
    db = torch.randn((n_db, D), device=dev, generator=g)
    db /= db.norm(dim=1, keepdim=True)
    q = torch.randn((nq, D), device=dev, generator=g)
    q /= q.norm(dim=1, keepdim=True)
    idx = Index(db, device=dev, keep_fp32=(prec != "bf16"))
    ms = timed(lambda: idx.ranks(q, precision=prec))
    ms_sc = timed(lambda: idx.scores(q, precision=prec))
    print("%s: ranks %d q x %d db x %d-D (%s): %.3f ms (scores alone %.3f ms), %.2f G pairs/s, floor frac %.4f, %s" %
          (tag, nq, n_db, D, prec, ms, ms_sc, n_db * nq / ms / 1e6, n_db * nq * 12 / 1e9 / (ms * 1e-3) / 6537.3, RANK_STATS))
    del idx, db, q
