#!/usr/bin/env python
"""Dense bf16 similarity GEMM alone (mdir_sim_scan_dense_bf16: tcgen05 / TMEM / TMA), as TFLOP/s against the sustained
bf16 peak: the C3 shape (1,024 of the 10,000 queries x 100,000 x 512) and a 2048-D variant.
    python tools/time_gemm.py [shape index]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mdir_b200.search import Index  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
SHAPES = ((100000, 512, 1024), (100000, 2048, 1024), (100000, 2048, 4096))
for n_db, D, nq in (SHAPES if len(sys.argv) < 2 else [SHAPES[int(sys.argv[1])]]):
    db = torch.randn((n_db, D), device=dev, generator=g)
    db /= db.norm(dim=1, keepdim=True)
    q = torch.randn((nq, D), device=dev, generator=g)
    q /= q.norm(dim=1, keepdim=True)
    idx = Index(db, device=dev, keep_fp32=False)
    out = torch.empty((nq, n_db), device=dev)
    for _ in range(3):
        idx.scores(q, out=out, precision="bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        idx.scores(q, out=out, precision="bf16")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = 2.0 * nq * n_db * D / (ms * 1e-3) / 1e12
    print("dense GEMM %d q x %d db x %d-D: %.3f ms (incl. the query pack), %.0f TFLOP/s = %.3f of 1397.3 sustained; output %.2f TB/s" %
          (nq, n_db, D, ms, tf, tf / 1397.3, nq * n_db * 4 / (ms * 1e-3) / 1e12))
    del idx, db, q, out
