"""CPU baselines of SURVEY.md 8d that are NOT numpy: the torch-CPU head (GeM + L2N + multi-scale aggregation + Lw
whitening) and OpenCV CLAHE, timed by bench.py beside the `head` / `clahe` side measurements.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product package).  When the reference checkout is importable
(/root/reference, or baseline/_ref on the GPU box) the reference's OWN functions are timed (kind "reference"):
  cirtorch.layers.pooling.GeM + normalization.L2N            (layers/functional.py:21-22,130-131)
  CirMultiscaleAggregation.aggregate_tensor                   (mdir/components/data/wrapper.py:109-119)
  CirtorchWhiten.postprocess                                  (wrapper.py:193-195)
otherwise a line-for-line torch restatement of the same calls (kind "port").  Two thread settings, both reported:
as shipped (torch 3 threads, cv2 1 thread: mdir/stages/validate.py:10-12, augmentation_transforms.py:6) and all cores."""
import os
import time

import numpy as np
import torch


def _cores():
    return len(os.sched_getaffinity(0))


def _ref_head():
    try:
        from oracle import ref_import
        if not ref_import.available():
            return None
        ref_import.import_reference()
        from cirtorch.layers.pooling import GeM
        from cirtorch.layers.normalization import L2N
        from mdir.components.data.wrapper import CirMultiscaleAggregation, CirtorchWhiten
        return GeM, L2N, CirMultiscaleAggregation, CirtorchWhiten
    except Exception:  # noqa: BLE001
        return None


def head_cpu(n_img, hws, C, p, P, m, threads):
    """Seconds per image of the reference head on torch-CPU for images of len(hws) scales (per-image loop, batch 1, as
    the reference runs it).  P (C, C), m (C, 1) numpy float32."""
    ref = _ref_head()
    g = torch.Generator().manual_seed(3)
    maps = [[torch.randn((1, C, h, w), generator=g).clamp_(min=0) for (h, w) in hws] for _ in range(n_img)]
    Pt, mt = torch.from_numpy(P).float(), torch.from_numpy(m).float().reshape(-1, 1)
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        if ref is not None:
            GeM, L2N, Agg, Whiten = ref
            pool, norm = GeM(p=p), L2N()
            wh = Whiten.__new__(Whiten)
            wh.P, wh.m, wh.dimensions = Pt, mt, C

            def one(fm):
                outs = [norm(pool(x)).squeeze(-1).squeeze(-1).permute(1, 0) for x in fm]          # imageretrievalnet.py:107-115
                v = Agg.aggregate_tensor(outs, len(fm), C, p)
                return wh.postprocess(v.reshape(-1), None, None)
            kind = "reference"
        else:
            def one(fm):
                outs = []
                for x in fm:
                    o = torch.nn.functional.avg_pool2d(x.clamp(min=1e-6).pow(p), (x.size(-2), x.size(-1))).pow(1. / p)   # functional.py:21-22
                    o = o / (torch.norm(o, p=2, dim=1, keepdim=True) + 1e-6).expand_as(o)                              # functional.py:130-131
                    outs.append(o.squeeze(-1).squeeze(-1).permute(1, 0))
                v = torch.zeros(C)
                for o in outs:                                                                                          # wrapper.py:109-119
                    v += o.pow(p).squeeze()
                v = (v / len(outs)).pow(1. / p)
                v = v / v.norm()
                X = torch.mm(Pt, v.unsqueeze(1) - mt)                                                                   # wrapper.py:193-195
                return (X / (torch.norm(X, p=2, dim=0, keepdim=True) + 1e-6)).squeeze()
            kind = "port"
        with torch.no_grad():
            one(maps[0])
            t0 = time.perf_counter()
            for fm in maps:
                one(fm)
            dt = (time.perf_counter() - t0) / n_img
    finally:
        torch.set_num_threads(old)
    return dt, kind


def clahe_cpu(n_img, H, W, clip, grid, threads):
    """Seconds per image of cv2.createCLAHE(clip, (grid, grid)).apply on uint8 (functional.py:114-117)."""
    import cv2
    rs = np.random.RandomState(2)
    imgs = [(rs.rand(H, W) ** 4 * 255).astype(np.uint8) for _ in range(n_img)]
    old = cv2.getNumThreads()
    cv2.setNumThreads(threads)
    try:
        cl = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid))
        cl.apply(imgs[0])
        t0 = time.perf_counter()
        for im in imgs:
            cl.apply(im)
        dt = (time.perf_counter() - t0) / n_img
    finally:
        cv2.setNumThreads(old)
    return dt


def head_and_clahe_baselines(C=2048, hws=((32, 24), (23, 17), (16, 12)), p=2.9137):
    cores = _cores()
    rs = np.random.RandomState(1)
    P = (rs.randn(C, C) / np.sqrt(C)).astype(np.float32)
    m = (rs.randn(C, 1) * 0.01).astype(np.float32)
    out = {"cores": cores}
    for tag, th, n in (("as_shipped_3_threads", 3, 6), ("all_cores", cores, 12)):
        dt, kind = head_cpu(n, list(hws), C, p, P, m, th)
        out["head_" + tag] = {"descriptors_per_s": 1.0 / dt, "ms_per_image": dt * 1e3, "threads": th, "kind": kind, "images": n}
    for tag, th, n in (("as_shipped_1_thread", 1, 8), ("all_cores", cores, 16)):
        dt = clahe_cpu(n, 768, 1024, 4, 8, th)
        out["clahe_" + tag] = {"images_per_s": 1.0 / dt, "ms_per_image": dt * 1e3, "threads": th, "kind": "reference (cv2.createCLAHE, the call functional.py:114-117 makes)", "images": n}
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(head_and_clahe_baselines(), indent=1))
