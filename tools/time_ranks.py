#!/usr/bin/env python
"""Time Index.ranks() (scores + full per-query ranking) on the C3 slice shape; run under ncu for the launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ranks.csv python tools/time_ranks.py 1"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index, RANK_STATS  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
n_db, D, nq = 100000, 512, 1024
db = torch.randn((n_db, D), device=dev, generator=g)
db /= db.norm(dim=1, keepdim=True)
q = torch.randn((nq, D), device=dev, generator=g)
q /= q.norm(dim=1, keepdim=True)
idx = Index(db, device=dev, keep_fp32=False)
r = idx.ranks(q, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    r = idx.ranks(q, precision="bf16")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("ranks %d q x %d db x %d-D: %.3f ms, %.2f G pairs/s, floor frac %.4f, %s" % (nq, n_db, D, ms, n_db * nq / ms / 1e6, n_db * nq * 12 / 1e9 / (ms * 1e-3) / 6537.3, RANK_STATS))
