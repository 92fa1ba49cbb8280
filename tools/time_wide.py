#!/usr/bin/env python
"""Tensor-bound shapes of the similarity scan: 128 queries per pass against WIDE launches (256 / 1,024 queries per pass as
(tile, 128-query block) work items).
   python tools/time_wide.py            # top-10 search at 1M x 2048 (DBA inner loop) and C3 (100k x 512)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)


def unit(n, d):
    x = torch.randn((n, d), device=dev, generator=g)
    return x / x.norm(dim=1, keepdim=True)


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n_db, D, nq_total, reps in ((1001001, 2048, 4096, 3), (100000, 512, 10240, 3)):
    db = unit(n_db, D)
    idx = Index(db, device=dev, keep_fp32=False)
    q = db[:nq_total].clone()
    for blk in (128, 256, 1024):
        s0, i0 = idx.search(q[:512], 10, precision="bf16", block_q=blk)
        ms = timeit(lambda: idx.search(q, 10, precision="bf16", block_q=blk, check=False), reps)
        fl = 2.0 * nq_total * n_db * D
        print("top-10 search %d q x %d x %d, %d queries/pass: %.2f ms, %.0f TFLOP/s (%.2f of 1397 sustained), overflow=%s" %
              (nq_total, n_db, D, blk, ms, fl / ms / 1e9, fl / ms / 1e9 / 1397.3, idx.check_overflow()))
        if blk == 128:
            ref = (s0.clone(), i0.clone())
        else:
            print("  %d/pass == 128/pass:" % blk, bool(torch.equal(ref[1], i0) and torch.equal(ref[0], s0)))
    if D == 512:
        sc = torch.empty((nq_total, n_db), dtype=torch.float32, device=dev)
        ms = timeit(lambda: idx.scores(q, out=sc), reps)
        print("dense scores %d x %d x %d: %.2f ms, %.0f TFLOP/s, %.2f TB/s written" % (nq_total, n_db, D, ms, 2.0 * nq_total * n_db * D / ms / 1e9, nq_total * n_db * 4 / ms / 1e9))
        del sc
    del idx, db, q
