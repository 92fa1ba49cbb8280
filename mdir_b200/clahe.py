"""CLAHE illumination normalisation (the N/D preprocessing of mdir's CLAHE scenario).

* ``clahe_u8`` -- the kernel boundary: batches of ragged uint8 planes on the device,
  bit-exact against ``cv2.createCLAHE(clip, (gx, gy)).apply`` (csrc/clahe.cu).
* ``ChannelClahe`` / ``ImageClahe`` and the transform classes ``ApplyClahe`` /
  ``AddClaheFromRgb`` / ``CreateClahedImage``: same names and arguments as
  mdir/components/data/transform/{functional.py:109-129, photometric_transforms.py:10-43}.
  The RGB<->Lab conversion around the L channel stays stock OpenCV (SURVEY.md 8f row f1
  marks it "next"); only the CLAHE itself is replaced.  These transforms touch the GPU, so
  run the DataLoader with num_workers=0 on this path (threading).
"""
import ctypes

import numpy as np
import torch

from . import _lib


_DESC_DTYPE = np.dtype([("src_off", np.int64), ("dst_off", np.int64), ("H", np.int32), ("W", np.int32),
                        ("src_pitch", np.int32), ("dst_pitch", np.int32)])


def _launch(sbase, dbase, descs_np, dev, max_h, max_w, clip_limit, grid):
    lib = _lib.lib()
    n = descs_np.shape[0]
    descs_d = torch.from_numpy(descs_np.view(np.uint8).reshape(-1)).to(dev)
    tiles_x, tiles_y = int(grid[0]), int(grid[1])
    ws = torch.empty(lib.mdir_clahe_workspace_bytes(n, tiles_x, tiles_y), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdir_clahe_u8(ctypes.c_void_p(sbase), ctypes.c_void_p(dbase), _lib.ptr(descs_d), n, int(max_h), int(max_w),
                                     float(clip_limit), tiles_x, tiles_y, _lib.ptr(ws), _lib.stream()), "mdir_clahe_u8")


def clahe_u8(images, clip_limit=4.0, grid=(8, 8)):
    """images: one (H,W) / (B,H,W) uint8 cuda tensor, or a list of (H,W) uint8 cuda tensors of
    different sizes (one launch for the whole ragged batch).  Returns the same structure.
    grid = (tiles_x, tiles_y) like cv2's tileGridSize."""
    if isinstance(images, torch.Tensor):
        _lib.require_cuda(images, "images")
        if images.dtype != torch.uint8 or images.dim() not in (2, 3):
            raise _lib.MdirError("clahe_u8 expects a uint8 (H,W) or (B,H,W) tensor")
        x = images if images.dim() == 3 else images.unsqueeze(0)
        if x.stride(2) != 1 or x.numel() == 0:
            if x.numel() == 0:
                raise _lib.MdirError("empty image")
            x = x.contiguous()
        B, H, W = x.shape
        out = torch.empty((B, H, W), dtype=torch.uint8, device=x.device)
        descs = np.empty(B, dtype=_DESC_DTYPE)
        descs["src_off"] = np.arange(B, dtype=np.int64) * x.stride(0)
        descs["dst_off"] = np.arange(B, dtype=np.int64) * (H * W)
        descs["H"], descs["W"], descs["src_pitch"], descs["dst_pitch"] = H, W, x.stride(1), W
        _launch(x.data_ptr(), out.data_ptr(), descs, x.device, H, W, clip_limit, grid)
        return out if images.dim() == 3 else out[0]
    planes = list(images)
    if not planes:
        return []
    for p in planes:
        _lib.require_cuda(p, "image")
        if p.dtype != torch.uint8 or p.dim() != 2 or p.numel() == 0:
            raise _lib.MdirError("clahe_u8 expects a list of non-empty (H,W) uint8 tensors")
    planes = [p if p.stride(1) == 1 else p.contiguous() for p in planes]
    dev = planes[0].device
    outs = [torch.empty((p.shape[0], p.shape[1]), dtype=torch.uint8, device=dev) for p in planes]
    sbase = min(p.data_ptr() for p in planes)
    dbase = min(o.data_ptr() for o in outs)
    descs = np.empty(len(planes), dtype=_DESC_DTYPE)
    for i, (p, o) in enumerate(zip(planes, outs)):
        descs[i] = (p.data_ptr() - sbase, o.data_ptr() - dbase, p.shape[0], p.shape[1], p.stride(0), o.stride(0))
    _launch(sbase, dbase, descs, dev, max(p.shape[0] for p in planes), max(p.shape[1] for p in planes), clip_limit, grid)
    return outs


class ChannelClahe:
    """transform/functional.py:109-117"""

    def __init__(self, clip_limit, grid_size, device="cuda"):
        if not isinstance(grid_size, tuple):
            grid_size = (int(grid_size), int(grid_size))
        self.clip_limit = int(clip_limit)
        self.grid_size = grid_size
        self.device = device

    def apply(self, chan):
        q = (np.asarray(chan) * 255).astype(np.uint8)                 # C truncation, functional.py:117
        out = clahe_u8(torch.from_numpy(np.ascontiguousarray(q)).to(self.device), self.clip_limit, self.grid_size)
        return out.cpu().numpy().astype(np.float32) / 255.0


def _cv2():
    import cv2
    return cv2


def rgb2normspace(img, colorspace):
    """transform/functional.py:24-27 (lab only: the colourspace of the CLAHE scenario)."""
    if colorspace.lower() != "lab":
        raise NotImplementedError("Colorspace %s is not supported" % colorspace)
    cv2 = _cv2()
    return (cv2.cvtColor(img, cv2.COLOR_RGB2LAB) + np.array([0, 128, 128], dtype=np.float32)) / np.array([100.0, 255.0, 255.0], dtype=np.float32)


def normspace2rgb(img, colorspace):
    """transform/functional.py:38-41"""
    if colorspace.lower() != "lab":
        raise NotImplementedError("Colorspace %s is not supported" % colorspace)
    cv2 = _cv2()
    return cv2.cvtColor((img * np.array([100.0, 255.0, 255.0], dtype=np.float32)) - np.array([0, 128, 128], dtype=np.float32), cv2.COLOR_LAB2RGB)


class ImageClahe(ChannelClahe):
    """transform/functional.py:120-129"""

    def __init__(self, clip_limit, grid_size, colorspace, device="cuda"):
        super().__init__(clip_limit, grid_size, device)
        self.colorspace = colorspace

    def apply(self, img):
        spc = rgb2normspace(img, self.colorspace)
        spc[:, :, 0] = super().apply(spc[:, :, 0])
        return normspace2rgb(spc, self.colorspace)


class ApplyClahe:
    """photometric_transforms.py:25-36; arguments arrive as strings from "apply_clahe:4:lab:8"."""

    def __init__(self, clip_limit=4, colorspace="lab", grid_size=8):
        self.params = {"clip_limit": clip_limit, "colorspace": colorspace, "grid_size": grid_size}
        self.clahe = ImageClahe(**self.params)

    def __call__(self, pic):
        return [self.clahe.apply(pic)]

    def __repr__(self):
        return "%s(%s)" % (self.__class__.__name__, ", ".join("%s=%s" % kv for kv in self.params.items()))


class CreateClahedImage(ApplyClahe):
    """photometric_transforms.py:39-43"""

    def __call__(self, pic):
        return [pic, self.clahe.apply(pic[:, :, :3])]


class AddClaheFromRgb:
    """photometric_transforms.py:10-23"""

    def __init__(self, clip_limit=4, grid_size=8, colorspace="lab"):
        self.params = {"clip_limit": int(clip_limit), "grid_size": grid_size, "colorspace": colorspace}
        self.clahe = ChannelClahe(clip_limit=int(clip_limit), grid_size=grid_size)

    def __call__(self, *pics):
        acc = []
        for pic in pics:
            assert isinstance(pic, np.ndarray)
            spc = rgb2normspace(pic[:, :, :3], self.params["colorspace"])
            chan = self.clahe.apply(spc[:, :, 0])
            acc.append(np.concatenate((pic, np.expand_dims(chan, axis=2)), axis=2))
        return acc
