"""Drop-in pooling / normalisation layers (same names, arguments, state_dict keys and
__repr__ as mdir/external/cirtorch/layers/{pooling,normalization,functional}.py), backed
by the sm_100a kernels in csrc/pool_head.cu through the C ABI.  Inference only: the
reference's eval stages run under torch.no_grad() (mdir/stages/validate.py:32)."""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib

POOL_GEM, POOL_MAC, POOL_SPOC = 0, 1, 2


def _prep(x):
    _lib.require_cuda(x, "feature map")
    if x.dtype != torch.float32:
        raise _lib.MdirError("feature maps must be float32 (reference dtype), got %s" % x.dtype)
    if x.dim() != 4:
        raise _lib.MdirError("expected (N,C,h,w), got shape %s" % (tuple(x.shape),))
    if torch.is_grad_enabled() and x.requires_grad:
        raise _lib.MdirError("mdir_b200 pooling is forward-only (eval stages); wrap the call in torch.no_grad()")
    return x.contiguous()


def _pool(kind, x, p=3.0, eps=1e-6):
    x = _prep(x)
    N, Cc, h, w = x.shape
    out = torch.empty((N, Cc, 1, 1), dtype=torch.float32, device=x.device)
    if N * Cc:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mdir_pool(kind, _lib.ptr(x), None, None, N, Cc, h * w, float(p), float(eps),
                                            _lib.ptr(out), _lib.stream()), "mdir_pool")
    return out


def mac(x):
    """cirtorch/layers/functional.py:11-12"""
    return _pool(POOL_MAC, x)


def spoc(x):
    """cirtorch/layers/functional.py:16-17"""
    return _pool(POOL_SPOC, x)


def gem(x, p=3, eps=1e-6):
    """cirtorch/layers/functional.py:21-22"""
    if isinstance(p, torch.Tensor):
        p = p.item()
    return _pool(POOL_GEM, x, p, eps)


def l2n(x, eps=1e-6):
    """cirtorch/layers/functional.py:130-131 -- x / (||x||_2 over dim 1 + eps)"""
    _lib.require_cuda(x, "l2n input")
    if x.dtype != torch.float32 or x.dim() < 2:
        raise _lib.MdirError("l2n expects a float32 tensor with >= 2 dims")
    xc = x.contiguous()
    N, Cc = xc.shape[0], xc.shape[1]
    inner = 1
    for s in xc.shape[2:]:
        inner *= s
    out = torch.empty_like(xc)
    if xc.numel():
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mdir_l2n(_lib.ptr(xc), N, Cc, inner, float(eps), _lib.ptr(out), _lib.stream()), "mdir_l2n")
    return out


class MAC(nn.Module):
    def forward(self, x):
        return mac(x)

    def __repr__(self):
        return self.__class__.__name__ + '()'


class SPoC(nn.Module):
    def forward(self, x):
        return spoc(x)

    def __repr__(self):
        return self.__class__.__name__ + '()'


class GeM(nn.Module):
    """cirtorch/layers/pooling.py:36-47: p is a learnable Parameter of shape [1] (state_dict key 'p')."""

    def __init__(self, p=3, eps=1e-6):
        super().__init__()
        self.p = Parameter(torch.ones(1) * p)
        self.eps = eps

    def forward(self, x):
        return gem(x, p=self.p.data, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + '(' + 'p=' + '{:.4f}'.format(self.p.data.tolist()[0]) + ', ' + 'eps=' + str(self.eps) + ')'


class L2N(nn.Module):
    """cirtorch/layers/normalization.py:10-20"""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return l2n(x, eps=self.eps)

    def __repr__(self):
        return self.__class__.__name__ + '(' + 'eps=' + str(self.eps) + ')'


# the registry the reference selects from (cirtorch/networks/imageretrievalnet.py:32-37);
# 'rmac' is outside the hot path and is left to the reference.
POOLING = {"mac": MAC, "spoc": SPoC, "gem": GeM}
