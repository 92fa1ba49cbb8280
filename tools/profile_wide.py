#!/usr/bin/env python
"""Two wide top-10 searches (1,024 database rows as queries) on the 1,001,001 x 2048 index, for ncu:
    ncu --set full --clock-control none -k regex:sim_scan -s 2 -c 2 python tools/profile_wide.py   (SAMPLE + FILTER of the second search)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
db = torch.randn((1001001, 2048), device=dev, generator=g)
db /= db.norm(dim=1, keepdim=True)
idx = Index(db, device=dev, keep_fp32=False)
for b in range(2):
    s, i = idx.search(db[b * 1024:(b + 1) * 1024], 10, precision="bf16", block_q=1024)
torch.cuda.synchronize()
print("self first:", bool((i[:, 0] == torch.arange(1024, 2048, device=dev)).all()))
