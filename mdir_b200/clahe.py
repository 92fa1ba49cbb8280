"""CLAHE illumination normalisation (the N/D preprocessing of mdir's CLAHE scenario).

* ``clahe_u8`` -- the kernel boundary: batches of ragged uint8 planes on the device,
  bit-exact against ``cv2.createCLAHE(clip, (gx, gy)).apply`` (csrc/clahe.cu).
* ``ChannelClahe`` / ``ImageClahe`` and the transform classes ``ApplyClahe`` /
  ``AddClaheFromRgb`` / ``CreateClahedImage``: same names and arguments as
  mdir/components/data/transform/{functional.py:109-129, photometric_transforms.py:10-43}.
  ``image_clahe`` runs the whole ImageClahe transform on the device: RGB->Lab follows OpenCV's
  float code path bit for bit (trilinear interpolation of its 33^3 fixed-point lattice), CLAHE on
  the uint8 L plane, Lab->RGB within ~1e-5.  No host OpenCV anywhere.  These transforms
  touch the GPU, so run the DataLoader with num_workers=0 on this path (threading): called inside
  a worker process they raise with that remedy instead of crashing in CUDA's fork check.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


_RGB_DESC_DTYPE = np.dtype([("rgb_off", np.int64), ("out_off", np.int64), ("l_off", np.int64), ("H", np.int32), ("W", np.int32)])
_DESC_DTYPE = np.dtype([("src_off", np.int64), ("dst_off", np.int64), ("H", np.int32), ("W", np.int32),
                        ("src_pitch", np.int32), ("dst_pitch", np.int32)])


_DESC_CACHE = {}          # (device, descriptor bytes) -> device copy: a regular batch re-uses its table call after call


def _launch(sbase, dbase, descs_np, dev, max_h, max_w, clip_limit, grid):
    lib = _lib.lib()
    n = descs_np.shape[0]
    key = (str(dev), descs_np.tobytes())
    descs_d = _DESC_CACHE.get(key)
    if descs_d is None:
        if len(_DESC_CACHE) >= 16:
            _DESC_CACHE.clear()
        descs_d = _DESC_CACHE[key] = torch.from_numpy(descs_np.view(np.uint8).reshape(-1).copy()).to(dev)
    tiles_x, tiles_y = int(grid[0]), int(grid[1])
    ws = torch.empty(lib.mdir_clahe_workspace_bytes(n, tiles_x, tiles_y), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdir_clahe_u8(ctypes.c_void_p(sbase), ctypes.c_void_p(dbase), _lib.ptr(descs_d), n, int(max_h), int(max_w),
                                     float(clip_limit), tiles_x, tiles_y, _lib.ptr(ws), _lib.stream()), "mdir_clahe_u8")


def clahe_u8(images, clip_limit=4.0, grid=(8, 8), out=None):
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
    outs = list(out) if out is not None else [torch.empty((p.shape[0], p.shape[1]), dtype=torch.uint8, device=dev) for p in planes]
    sbase = min(p.data_ptr() for p in planes)
    dbase = min(o.data_ptr() for o in outs)
    descs = np.empty(len(planes), dtype=_DESC_DTYPE)
    for i, (p, o) in enumerate(zip(planes, outs)):
        descs[i] = (p.data_ptr() - sbase, o.data_ptr() - dbase, p.shape[0], p.shape[1], p.stride(0), o.stride(0))
    _launch(sbase, dbase, descs, dev, max(p.shape[0] for p in planes), max(p.shape[1] for p in planes), clip_limit, grid)
    return outs


# ---------------------------------------------------------------- RGB <-> Lab on the device
_HERE = os.path.dirname(os.path.abspath(__file__))
_LAB_CACHE = {}


def _spline_build(fv):
    """Natural cubic spline through len(fv) unit-spaced points -> (n, 4) coefficients (OpenCV's splineBuild)."""
    n = len(fv) - 1
    tab = np.zeros(n * 4)
    cn = 0.0
    for i in range(1, n):
        t = (fv[i + 1] - fv[i] * 2 + fv[i - 1]) * 3
        l = 1.0 / (4 - tab[(i - 1) * 4])
        tab[i * 4] = l
        tab[i * 4 + 1] = (t - tab[(i - 1) * 4 + 1]) * l
    for i in range(n - 1, -1, -1):
        c = tab[i * 4 + 1] - tab[i * 4] * cn
        b = fv[i + 1] - fv[i] - (cn + c * 2) / 3
        d = (cn - c) / 3
        tab[i * 4:i * 4 + 4] = (fv[i], b, c, d)
        cn = c
    return tab.reshape(n, 4)


def lab_tables(device):
    """(lut (33,33,33,4) int16, gamma_tab (1024,4) fp32) on `device`, cached."""
    key = str(device)
    if key not in _LAB_CACHE:
        lut = np.load(os.path.join(_HERE, "data", "rgb2lab_lut_s16.npy"))
        lut4 = np.zeros(lut.shape[:3] + (4,), np.int16)
        lut4[..., :3] = lut
        x = np.arange(1025) / 1024.0
        g = np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1 / 2.4) - 0.055)       # sRGB gamma (Lab2RGBfloat)
        tab = _spline_build(g).astype(np.float32)
        _LAB_CACHE[key] = (torch.from_numpy(lut4).to(device), torch.from_numpy(tab).to(device))
    return _LAB_CACHE[key]


def _check_rgb(imgs):
    for im in imgs:
        _lib.require_cuda(im, "image")
        if im.dtype != torch.float32 or im.dim() != 3 or im.shape[2] != 3 or im.numel() == 0:
            raise _lib.MdirError("expected non-empty (H,W,3) float32 RGB tensors")
    return [im.contiguous() for im in imgs]


def _rgb_descs(imgs, compact_out):
    """Descriptor table of a ragged RGB batch.  Inputs are addressed relative to the lowest input pointer (they stay
    where they are); outputs go to ONE compact buffer, image after image (so separately allocated inputs that sit
    gigabytes apart in the caching allocator cost nothing).  -> (sbase, descs_np, l_off, out_off, npx)"""
    npx = [im.shape[0] * im.shape[1] for im in imgs]
    l_off = np.concatenate([[0], np.cumsum([(n + 15) // 16 * 16 for n in npx])]).astype(np.int64)
    out_off = np.concatenate([[0], np.cumsum([(3 * n + 3) // 4 * 4 for n in npx])]).astype(np.int64)
    sbase = min(im.data_ptr() for im in imgs)
    descs = np.zeros(len(imgs), dtype=_RGB_DESC_DTYPE)
    for i, im in enumerate(imgs):
        descs[i] = ((im.data_ptr() - sbase) // 4, int(out_off[i]) if compact_out else 0, int(l_off[i]), im.shape[0], im.shape[1])
    return sbase, descs, l_off, out_off, npx


def rgb_l_clahe_u8(images, clip_limit=4, grid=(8, 8)):
    """The L channel of the reference's normalised Lab space, quantised and equalised as ChannelClahe does
    (transform/functional.py:24-27,109-117): RGB float32 HWC in [0,1] -> trunc(L/100*255) uint8 (OpenCV's
    interpolated float RGB2Lab, bit for bit) -> CLAHE.  -> list of (H,W) uint8 cuda tensors."""
    lib = _lib.lib()
    imgs = _check_rgb(list(images))
    if not imgs:
        return []
    dev = imgs[0].device
    lut, _ = lab_tables(dev)
    sbase, descs, l_off, _, npx = _rgb_descs(imgs, False)
    l_in = torch.empty((int(l_off[-1]),), dtype=torch.uint8, device=dev)
    descs_d = torch.from_numpy(descs.view(np.uint8).reshape(-1)).to(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdir_rgb_to_l_u8(ctypes.c_void_p(sbase), _lib.ptr(descs_d), len(imgs), max(npx), _lib.ptr(lut), _lib.ptr(l_in),
                                        _lib.stream()), "mdir_rgb_to_l_u8")
    planes = [l_in[int(l_off[i]):int(l_off[i]) + npx[i]].view(im.shape[0], im.shape[1]) for i, im in enumerate(imgs)]
    return clahe_u8(planes, clip_limit, grid)


def image_clahe(images, clip_limit=4, grid=(8, 8)):
    """ImageClahe.apply (transform/functional.py:120-129, colorspace 'lab') entirely on the device:
    RGB float32 HWC in [0,1] -> Lab (OpenCV's interpolated float path) -> CLAHE on L as uint8 -> RGB.
    images: one (H,W,3) cuda tensor or a list of them (ragged sizes, one launch per stage).  The results are views
    into one compact output buffer."""
    lib = _lib.lib()
    single = isinstance(images, torch.Tensor)
    imgs = [images] if single else list(images)
    if not imgs:
        return []
    imgs = _check_rgb(imgs)
    dev = imgs[0].device
    lut, gamma = lab_tables(dev)
    sbase, descs, l_off, out_off, npx = _rgb_descs(imgs, True)
    l_in = torch.empty((int(l_off[-1]),), dtype=torch.uint8, device=dev)
    arena = torch.empty((int(out_off[-1]),), dtype=torch.float32, device=dev)
    descs_d = torch.from_numpy(descs.view(np.uint8).reshape(-1)).to(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mdir_rgb_to_l_u8(ctypes.c_void_p(sbase), _lib.ptr(descs_d), len(imgs), max(npx), _lib.ptr(lut), _lib.ptr(l_in),
                                        _lib.stream()), "mdir_rgb_to_l_u8")
        l_out = torch.empty_like(l_in)
        planes = [l_in[int(l_off[i]):int(l_off[i]) + npx[i]].view(im.shape[0], im.shape[1]) for i, im in enumerate(imgs)]
        oplanes = [l_out[int(l_off[i]):int(l_off[i]) + npx[i]].view(im.shape[0], im.shape[1]) for i, im in enumerate(imgs)]
        clahe_u8(planes, clip_limit, grid, out=oplanes)
        _lib.check(lib.mdir_lab_clahe_to_rgb(ctypes.c_void_p(sbase), _lib.ptr(descs_d), len(imgs), max(npx), _lib.ptr(lut),
                                             _lib.ptr(gamma), _lib.ptr(l_out), _lib.ptr(arena), _lib.stream()), "mdir_lab_clahe_to_rgb")
    outs = [arena[int(out_off[i]):int(out_off[i]) + im.numel()].view(im.shape) for i, im in enumerate(imgs)]
    return outs[0] if single else outs


def _not_in_worker(what):
    """The transform classes launch CUDA kernels: inside a forked DataLoader worker that cannot work ("Cannot
    re-initialize CUDA in forked subprocess").  Fail with the remedy instead."""
    info = torch.utils.data.get_worker_info()
    if info is not None:
        raise _lib.MdirError("%s runs on the GPU and was called inside DataLoader worker %d: use num_workers=0 on this path "
                             "(mdir_b200.extract_vectors does), or mdir_b200.install(transforms=False) to keep the reference's "
                             "CPU transforms for loaders with workers" % (what, info.id))


def _require_lab(colorspace):
    if str(colorspace).lower() != "lab":
        raise NotImplementedError("colorspace %r: only 'lab' (the CLAHE scenario's) runs on the device; mdir_b200.install() "
                                  "leaves other colourspaces to the reference's own transform classes" % (colorspace,))


class ChannelClahe:
    """transform/functional.py:109-117"""

    def __init__(self, clip_limit, grid_size, device="cuda"):
        if not isinstance(grid_size, tuple):
            grid_size = (int(grid_size), int(grid_size))
        self.clip_limit = int(clip_limit)
        self.grid_size = grid_size
        self.device = device

    def apply(self, chan):
        _not_in_worker("ChannelClahe")
        q = (np.asarray(chan) * 255).astype(np.uint8)                 # C truncation, functional.py:117
        out = clahe_u8(torch.from_numpy(np.ascontiguousarray(q)).to(self.device), self.clip_limit, self.grid_size)
        return out.cpu().numpy().astype(np.float32) / 255.0


class ImageClahe(ChannelClahe):
    """transform/functional.py:120-129 -- colorspace 'lab', float32 RGB in [0,1], the whole transform on the device
    (no host OpenCV anywhere on this path)."""

    def __init__(self, clip_limit, grid_size, colorspace, device="cuda"):
        super().__init__(clip_limit, grid_size, device)
        _require_lab(colorspace)
        self.colorspace = colorspace

    def apply(self, img):
        _not_in_worker("ImageClahe")
        img = np.asarray(img)
        if img.dtype != np.float32 or img.ndim != 3 or img.shape[2] != 3:
            raise _lib.MdirError("ImageClahe.apply expects an (H,W,3) float32 RGB image in [0,1] (what pil2np produces); got %s %s"
                                 % (img.dtype, img.shape))
        x = torch.from_numpy(np.ascontiguousarray(img)).to(self.device)
        return image_clahe(x, self.clip_limit, self.grid_size).cpu().numpy()


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
    """photometric_transforms.py:10-23: append the CLAHE'd L channel of each picture as a fourth channel.
    RGB -> L (OpenCV's float RGB2Lab, bit for bit), the uint8 quantisation and CLAHE all run on the device."""

    def __init__(self, clip_limit=4, grid_size=8, colorspace="lab", device="cuda"):
        _require_lab(colorspace)
        self.params = {"clip_limit": int(clip_limit), "grid_size": grid_size, "colorspace": colorspace}
        self.clahe = ChannelClahe(clip_limit=int(clip_limit), grid_size=grid_size, device=device)

    def __call__(self, *pics):
        _not_in_worker("AddClaheFromRgb")
        for pic in pics:
            assert isinstance(pic, np.ndarray)
        dev = self.clahe.device
        rgb = [torch.from_numpy(np.ascontiguousarray(pic[:, :, :3], dtype=np.float32)).to(dev) for pic in pics]
        chans = rgb_l_clahe_u8(rgb, self.clahe.clip_limit, self.clahe.grid_size)
        acc = []
        for pic, ch in zip(pics, chans):
            chan = ch.cpu().numpy().astype(np.float32) / 255.0
            acc.append(np.concatenate((pic, np.expand_dims(chan, axis=2)), axis=2))
        return acc
