#!/usr/bin/env python
"""A few eager (un-graphed) search steps at a given shard size, for ncu launch lists / full captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py [rows] [n_q]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index, pack_bf16  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1001001
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 70
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
db = torch.empty((rows, 2048), device=dev)
for r0 in range(0, rows, 65536):
    blk = torch.randn((min(65536, rows - r0), 2048), device=dev, generator=g)
    db[r0:r0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
q = torch.randn((n_q, 2048), device=dev, generator=g)
q /= q.norm(dim=1, keepdim=True)
index = Index.from_packed(pack_bf16(db), db32=db)
index.stats()
for _ in range(4):
    s, i = index.search(q, 100, precision="fp32", check=False)
torch.cuda.synchronize()
print("flagged:", index.check_overflow(), "best:", s[0, :3].tolist())
