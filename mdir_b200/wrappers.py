"""Drop-in eval-stage wrappers (mdir/components/data/wrapper.py) and the batched head.

* ``CirMultiscaleAggregation`` / ``CirtorchWhiten``: same constructor arguments, same
  preprocess/postprocess contract as wrapper.py:84-136 and :181-195, so mdir's ``Compose``
  (wrapper.py:17-37) drives them unchanged.
* ``RetrievalHead``: the B200-first form of the same arithmetic -- all images x scales of a
  batch pooled in ONE kernel pass over the feature maps, then aggregation + Lw projection.
* ``whitenapply``: cirtorch/utils/whiten.py:4-12.
"""
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .layers import POOL_GEM, POOL_MAC, POOL_SPOC

_KIND = {"gem": POOL_GEM, "mac": POOL_MAC, "spoc": POOL_SPOC}


def _load_lw(whitening):
    """A {'m','P'} dict, or a path to the .pkl the reference loads with tools/utils.py:44-50."""
    if isinstance(whitening, dict):
        return whitening
    if isinstance(whitening, str) and whitening.endswith(".pkl"):
        with open(whitening, "rb") as handle:
            return pickle.load(handle)
    raise ValueError("Unsupported whitening source %r (expected dict or .pkl path)" % (whitening,))


def ms_aggregate(vectors, nscales, msp, m=None, normalized=True, l2n_eps=1e-6):
    """vectors: (n_img, S, C) fp32 cuda.  -> (n_img, C).  See mdir_ms_aggregate."""
    _lib.require_cuda(vectors, "vectors")
    v = vectors.contiguous()
    n_img, S, Cc = v.shape
    assert S == nscales, "%s != %s" % (S, nscales)
    out = torch.empty((n_img, Cc), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().mdir_ms_aggregate(_lib.ptr(v), n_img, S, Cc, -1.0 if normalized else float(l2n_eps),
                                                float(msp), _lib.ptr(m), _lib.ptr(out), _lib.stream()), "mdir_ms_aggregate")
    return out


def split_lw(P):
    """P (dims, D) fp32 cuda -> (dims, 3D) [hi|hi|lo] for the tensor-core (3xTF32) projection."""
    P = P.contiguous()
    out = torch.empty((P.shape[0], 3 * P.shape[1]), dtype=torch.float32, device=P.device)
    with torch.cuda.device(P.device):
        _lib.check(_lib.lib().mdir_split_tf32x3(_lib.ptr(P), P.shape[0], P.shape[1], 0, _lib.ptr(out), _lib.stream()), "mdir_split_tf32x3")
    return out


def whiten_project(v, m, P, dims, renorm_eps=1e-6, Px3=None):
    """v (n, D) fp32 cuda, m (D,) or None, P (>=dims, D) -> (n, dims).  Batches of more than 4
    vectors go to the tensor cores when the pre-split Px3 = split_lw(P) is supplied."""
    _lib.require_cuda(v, "v")
    v = v.contiguous()
    n, D = v.shape
    out = torch.empty((n, dims), dtype=torch.float32, device=v.device)
    if Px3 is not None and n > 4 and D % 4 == 0:
        lib = _lib.lib()
        with torch.cuda.device(v.device):
            ws = torch.empty(lib.mdir_whiten_tc_workspace_bytes(n, D, int(dims)), dtype=torch.uint8, device=v.device)
            _lib.check(lib.mdir_whiten_project_tc(_lib.ptr(v), _lib.ptr(m), n, D, _lib.ptr(Px3), int(dims), float(renorm_eps),
                                                  _lib.ptr(out), _lib.ptr(ws), _lib.stream()), "mdir_whiten_project_tc")
        return out
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().mdir_whiten_project(_lib.ptr(v), _lib.ptr(m), n, D, _lib.ptr(P), int(dims), float(renorm_eps),
                                                  _lib.ptr(out), _lib.stream()), "mdir_whiten_project")
    return out


class Wrapper:
    """wrapper.py:40-60 -- the (pre, post) protocol mdir's Compose expects."""

    def __init__(self, device):
        self.device = device

    def preprocess(self, tensor, _model):
        return tensor, None

    def postprocess(self, tensor, _model, _meta):
        return tensor


class CirMultiscaleAggregation(Wrapper):
    """wrapper.py:84-136.  preprocess stays stock torch (it feeds the backbone)."""

    def __init__(self, scales, device):
        super().__init__(device)
        if isinstance(scales, str):
            scales = {"True": True, "False": False}[scales]
        if isinstance(scales, bool):
            scales = [1, 1. / np.sqrt(2), 1. / 2] if scales else [1]
        self.scales = scales

    def preprocess(self, tensor, _model):
        if len(self.scales) == 1:
            return tensor if isinstance(tensor, list) else [tensor], isinstance(tensor, list)
        acc = []
        if isinstance(tensor, list):
            for single in tensor:
                for scale in self.scales:
                    acc.append(F.interpolate(single, scale_factor=scale, mode='bilinear', align_corners=False))
            return acc, True
        return [F.interpolate(tensor, scale_factor=scale, mode='bilinear', align_corners=False) for scale in self.scales], False

    @staticmethod
    def aggregate_tensor(tensor, nscales, outputdim, msp):
        assert len(tensor) == nscales, "%s != %s" % (len(tensor), nscales)
        stacked = torch.stack([t.reshape(-1) for t in tensor]).unsqueeze(0)       # (1, S, C)
        assert stacked.shape[2] == outputdim
        return ms_aggregate(stacked, nscales, msp)[0]

    def postprocess(self, tensor, model, waslist):
        msp = 1
        if len(self.scales) > 1 and model.meta['pooling'] == 'gem' and not model.meta['regional'] and not model.meta['whitening']:
            msp = model.pool.p.item()
        S = len(self.scales)
        if not waslist:
            return self.aggregate_tensor(tensor, S, model.meta['out_channels'], msp)
        assert len(tensor) % S == 0, "%s %% %s != 0" % (len(tensor), S)
        stacked = torch.stack([t.reshape(-1) for t in tensor]).view(len(tensor) // S, S, -1)
        return list(ms_aggregate(stacked, S, msp))

    def __repr__(self):
        return f"{self.__class__.__name__}(scales={self.scales})"


class CirtorchWhiten(Wrapper):
    """wrapper.py:181-195 -- Lw {m, P} applied with optional dimensionality reduction."""

    def __init__(self, whitening, dimensions, device):
        super().__init__(device)
        whitening = _load_lw(whitening)
        self.P = torch.tensor(whitening['P'], dtype=torch.float32, device=device).contiguous()
        self.m = torch.tensor(whitening['m'], dtype=torch.float32, device=device).reshape(-1).contiguous()
        self.dimensions = dimensions or self.P.shape[0]

    def postprocess(self, tensor, model, _meta):
        return whiten_project(tensor.reshape(1, -1), self.m, self.P, self.dimensions)[0]


def whitenapply(X, m, P, dimensions=None, device="cuda"):
    """cirtorch/utils/whiten.py:4-12: X (D,N), m (D,1), P (D,D) ndarrays -> (dims,N) ndarray of
    X's dtype.  Arithmetic is fp32 on the device (the reference's fp32 wrapper and fp64 numpy
    forms agree to 2.2e-7)."""
    if not dimensions:
        dimensions = P.shape[0]
    Xt = torch.as_tensor(np.ascontiguousarray(np.asarray(X).T), dtype=torch.float32).to(device)
    mt = torch.as_tensor(np.asarray(m).reshape(-1), dtype=torch.float32).to(device)
    Pt = torch.as_tensor(np.ascontiguousarray(np.asarray(P)[:dimensions]), dtype=torch.float32).to(device)
    out = whiten_project(Xt, mt, Pt, dimensions, Px3=split_lw(Pt) if Pt.shape[1] % 4 == 0 else None)
    return np.ascontiguousarray(out.cpu().numpy().T).astype(np.asarray(X).dtype, copy=False)


class PackedMaps:
    """Device-side description of a ragged batch of feature maps (see RetrievalHead.pack)."""

    def __init__(self, base, off, hw, n_maps, C, device, keepalive=None):
        self.base, self.off, self.hw, self.n_maps, self.C, self.device = base, off, hw, n_maps, C, device
        self._keepalive = keepalive


class RetrievalHead:
    """Batched post-backbone head for B images x S scales:
    pool (GeM/MAC/SPoC) -> L2N -> multi-scale aggregation (msp rule) -> [Lw centre + project +
    renorm].  Feature maps are read exactly once, by one kernel launch, wherever they live."""

    def __init__(self, pooling="gem", p=3.0, eps=1e-6, whitening=None, dimensions=None, nscales=1,
                 regional=False, model_whitening=False, device="cuda"):
        self.kind = _KIND[pooling]
        self.pooling = pooling
        self.p = float(p)
        self.eps = float(eps)
        self.nscales = int(nscales)
        self.device = torch.device(device)
        # wrapper.py:122-124
        self.msp = self.p if (self.nscales > 1 and pooling == "gem" and not regional and not model_whitening) else 1.0
        self.P = self.m = None
        self.dimensions = None
        if whitening is not None:
            lw = _load_lw(whitening)
            self.P = torch.tensor(lw['P'], dtype=torch.float32, device=self.device).contiguous()
            self.m = torch.tensor(lw['m'], dtype=torch.float32, device=self.device).reshape(-1).contiguous()
            self.dimensions = dimensions or self.P.shape[0]
            self.Px3 = split_lw(self.P) if self.P.shape[1] % 4 == 0 else None

    def pack(self, fmaps):
        """Describe a ragged set of feature maps once: flat list of (1,C,h,w)/(C,h,w) fp32 cuda
        tensors (image-major, scale-minor) -> PackedMaps (device offset / size tables).  The maps
        are NOT copied; the table stays valid as long as the tensors keep their storage (e.g. a
        pre-allocated arena the backbone writes into)."""
        maps = []
        for f in fmaps:
            _lib.require_cuda(f, "fmap")
            if f.dtype != torch.float32:
                raise _lib.MdirError("feature maps must be float32")
            f = f.contiguous()
            maps.append(f[0] if f.dim() == 4 else f)
        Cc = maps[0].shape[0]
        base = min(f.data_ptr() for f in maps)
        off = torch.tensor([(f.data_ptr() - base) // 4 for f in maps], dtype=torch.int64)
        hw = torch.tensor([f.shape[1] * f.shape[2] for f in maps], dtype=torch.int32)
        dev = maps[0].device
        return PackedMaps(base, off.to(dev), hw.to(dev), len(maps), Cc, dev, maps)

    def pool_maps(self, fmaps):
        """fmaps: PackedMaps, a flat list of tensors (packed on the fly), or one uniform
        (n_maps, C, h, w) tensor.  -> (n_maps, C) pooled (not yet normalised)."""
        lib = _lib.lib()
        if isinstance(fmaps, torch.Tensor):
            x = _lib.require_cuda(fmaps, "fmaps").contiguous()
            n_maps, Cc = x.shape[0], x.shape[1]
            out = torch.empty((n_maps, Cc), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.check(lib.mdir_pool(self.kind, _lib.ptr(x), None, None, n_maps, Cc, x.shape[2] * x.shape[3], self.p, self.eps,
                                         _lib.ptr(out), _lib.stream()), "mdir_pool")
            return out
        pm = fmaps if isinstance(fmaps, PackedMaps) else self.pack(fmaps)
        out = torch.empty((pm.n_maps, pm.C), dtype=torch.float32, device=pm.device)
        with torch.cuda.device(pm.device):
            import ctypes
            _lib.check(lib.mdir_pool(self.kind, ctypes.c_void_p(pm.base), _lib.ptr(pm.off), _lib.ptr(pm.hw), pm.n_maps, pm.C, 0, self.p,
                                     self.eps, _lib.ptr(out), _lib.stream()), "mdir_pool")
        return out

    def capture(self, packed):
        """Static shapes (a PackedMaps over a pre-allocated arena, or one uniform tensor): record the whole head once
        into a CUDA graph.  Returns replay() -> the (n_img, dims) output tensor of the graph; the nine launches of the
        head then cost one graph launch instead of nine host round trips (the small kernels are latency-bound)."""
        dev = packed.device
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self(packed)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self(packed)

        def replay():
            graph.replay()
            return out

        replay.graph, replay.out = graph, out
        return replay

    def __call__(self, fmaps):
        pooled = self.pool_maps(fmaps)
        n_maps, Cc = pooled.shape
        assert n_maps % self.nscales == 0
        v = ms_aggregate(pooled.view(n_maps // self.nscales, self.nscales, Cc), self.nscales, self.msp, normalized=False)
        if self.P is not None:
            v = whiten_project(v, self.m, self.P, self.dimensions, Px3=self.Px3)
        return v
