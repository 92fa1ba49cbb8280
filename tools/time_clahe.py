#!/usr/bin/env python
"""Times the two CLAHE kernels on the bench shape (256 images of 768 x 1024) with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mdir_b200  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
for name, imgs in (("gamma4", (torch.rand((256, 768, 1024), device=dev, generator=g) ** 4 * 255).to(torch.uint8)),
                   ("uniform", (torch.rand((256, 768, 1024), device=dev, generator=g) * 255).to(torch.uint8)),
                   ("flat", torch.full((256, 768, 1024), 17, device=dev, dtype=torch.uint8))):
    for _ in range(3):
        mdir_b200.clahe_u8(imgs, 4, (8, 8))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        mdir_b200.clahe_u8(imgs, 4, (8, 8))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print("%-8s %.3f ms per 256 images  (%.0f images/s, %.3f of 6537 GB/s)" % (name, ms, 256 / ms * 1e3, 256 * 2 * 768 * 1024 / 1e9 / (ms * 1e-3) / 6537))
