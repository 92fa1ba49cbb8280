#!/usr/bin/env python
"""The bench step in isolation, for ncu: builds the R1M-shaped index once and runs a few search
steps (optionally the head / CLAHE kernels).  Numbers printed under a profiler are not bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import mdir_b200  # noqa: E402
from mdir_b200.search import Index, pack_bf16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--rows", type=int, default=1001001)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--nq", type=int, default=70)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--what", default="search", choices=["search", "head", "clahe", "ranks"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
if a.what == "search":
    db = torch.empty((a.rows, a.dim), dtype=torch.float32, device=dev)
    for r0 in range(0, a.rows, 65536):
        blk = torch.randn((min(65536, a.rows - r0), a.dim), device=dev, generator=g)
        db[r0:r0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    q = torch.randn((a.nq, a.dim), device=dev, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    index = Index.from_packed(pack_bf16(db), db32=db)
    torch.cuda.synchronize()
    for _ in range(a.steps):
        s, i = index.search(q, 100, precision=a.precision, check=False)
    torch.cuda.synchronize()
    print("overflow:", index.check_overflow(), "best:", s[0, :3].tolist())
elif a.what == "head":
    C, hws = 2048, [(32, 24), (23, 17), (16, 12)]
    fm = [torch.randn((1, C, h, w), device=dev, generator=g).clamp_(min=0) for _ in range(192) for (h, w) in hws]
    P = (torch.randn((C, C), generator=g, device=dev) / C ** 0.5).cpu().numpy()
    m = torch.zeros((C, 1)).numpy()
    head = mdir_b200.RetrievalHead("gem", p=2.9137, whitening={"P": P, "m": m}, nscales=3, device=dev)
    packed = head.pack(fm)
    for _ in range(a.steps):
        out = head(packed)
    torch.cuda.synchronize()
elif a.what == "clahe":
    imgs = (torch.rand((256, 768, 1024), device=dev, generator=g) ** 4 * 255).to(torch.uint8)
    for _ in range(a.steps):
        out = mdir_b200.clahe_u8(imgs, 4, (8, 8))
    torch.cuda.synchronize()
elif a.what == "ranks":
    db = torch.randn((100000, 512), device=dev, generator=g)
    db = db / db.norm(dim=1, keepdim=True)
    q = db[:2048] + 0.01
    index = Index(db, device=dev, keep_fp32=False)
    for _ in range(a.steps):
        r = index.ranks(q)
    torch.cuda.synchronize()
