"""Learning the whitening {m, P} on the device.

Reference: mdir/external/cirtorch/utils/whiten.py:14-53 (``pcawhitenlearn``, ``whitenlearn``): numpy on
(D, N) matrices whose columns are images.  The O(D^2 N) contractions (``np.dot(df, df.T)``,
``np.dot(P, X - m)`` / ``np.dot(Xc, Xc.T)``) run in this library's fp64 kernels (``mdir_gemm_f64`` with
the mean subtraction fused into the operand loads); the D x D factorisations (Cholesky, triangular
inverse, symmetric eigendecomposition) are cuSOLVER calls through torch.linalg, as LAPACK is for the
reference.  Everything is fp64 on the device; eigenvector signs are arbitrary, exactly as with
``np.linalg.eig``.
"""
import os

import numpy as np
import torch

from . import _lib


def _dev_f64(X, device):
    """(D, N) fp32/fp64 numpy or torch -> contiguous fp64 device tensor (fp32 widened on the device)."""
    t = torch.as_tensor(X)
    if t.dim() != 2:
        raise _lib.MdirError("expected a (D, N) matrix")
    if t.dtype == torch.float64:
        return t.to(device).contiguous()
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    t = t.to(device).contiguous()
    out = torch.empty(t.shape, dtype=torch.float64, device=device)
    _lib.check(_lib.lib().mdir_f32_to_f64(_lib.ptr(t), t.numel(), _lib.ptr(out), _lib.stream()), "mdir_f32_to_f64")
    return out


def gemm_f64(A, B, b_is_kxn, alpha=1.0, a_sub=None, b_sub=None):
    """alpha * (A - a_sub[:, None]) @ (B - b_sub[:, None]).T  (b_is_kxn False; B is (N, K))
    or alpha * (A - a_sub[:, None]) @ (B - b_sub[:, None])    (b_is_kxn True;  B is (K, N)); fp64 device tensors."""
    lib = _lib.lib()
    A, B = A.contiguous(), B.contiguous()              # cuSOLVER results come back column-major
    M, K = A.shape
    N = B.shape[1] if b_is_kxn else B.shape[0]
    if (B.shape[0] if b_is_kxn else B.shape[1]) != K:
        raise _lib.MdirError("inner dimensions differ")
    C = torch.empty((M, N), dtype=torch.float64, device=A.device)
    nws = lib.mdir_gemm_f64_workspace_bytes(M, N, K)
    ws = torch.empty((nws,), dtype=torch.uint8, device=A.device) if nws else None
    _lib.check(lib.mdir_gemm_f64(_lib.ptr(A), A.stride(0), _lib.ptr(a_sub), _lib.ptr(B), B.stride(0), _lib.ptr(b_sub), 1 if b_is_kxn else 0,
                                 M, N, K, float(alpha), _lib.ptr(C), N, _lib.ptr(ws), _lib.stream()), "mdir_gemm_f64")
    return C


def _cols_mean(X, idx=None):
    D, N = X.shape
    mean = torch.empty((D,), dtype=torch.float64, device=X.device)
    n_idx = N if idx is None else idx.numel()
    _lib.check(_lib.lib().mdir_cols_mean_f64(_lib.ptr(X), X.stride(0), D, N, _lib.ptr(idx), n_idx, _lib.ptr(mean), _lib.stream()),
               "mdir_cols_mean_f64")
    return mean


def _sorted_eigh(S):
    """Eigenpairs of the symmetric S by descending eigenvalue (whiten.py:25-28 / 46-49)."""
    S = (S + S.t()) * 0.5
    eigval, eigvec = torch.linalg.eigh(S)
    return torch.flip(eigval, dims=(0,)), torch.flip(eigvec, dims=(1,))


def _cholesky(S):
    """whiten.py:55-70: add 1e-10, 1e-9, ... to the diagonal until the factorisation succeeds."""
    alpha = 0.0
    eye = torch.eye(S.shape[0], dtype=S.dtype, device=S.device)
    while True:
        L, info = torch.linalg.cholesky_ex(S + alpha * eye)
        if int(info.item()) == 0:
            return L
        alpha = 1e-10 if alpha == 0 else alpha * 10
        print(">>>> {}::cholesky: Matrix is not positive definite, adding {:.0e} on the diagonal".format(os.path.basename(__file__), alpha))


def whitenlearn(X, qidxs, pidxs, device="cuda"):
    """Lw whitening from matching pairs (whiten.py:37-53).  X (D, N); qidxs/pidxs: column indices of
    matching pairs.  Returns (m (D, 1), P (D, D)) fp64 numpy, the dict the Lw head consumes."""
    dev = torch.device(device)
    with torch.cuda.device(dev):
        X = _dev_f64(X, dev)
        D, N = X.shape
        q = torch.as_tensor(np.asarray(qidxs, dtype=np.int64)).to(dev)
        p = torch.as_tensor(np.asarray(pidxs, dtype=np.int64)).to(dev)
        if q.numel() != p.numel() or q.numel() == 0:
            raise _lib.MdirError("qidxs and pidxs must be equally long and non-empty")
        if int(torch.min(torch.minimum(q, p)).item()) < -N or int(torch.max(torch.maximum(q, p)).item()) >= N:
            raise IndexError("pair index out of range for %d columns" % N)
        q, p = q % N, p % N                                        # numpy-style negative indices
        n_pairs = q.numel()
        m = _cols_mean(X, q)
        df = torch.empty((D, n_pairs), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mdir_pair_diff_f64(_lib.ptr(X), X.stride(0), D, N, _lib.ptr(q), _lib.ptr(p), n_pairs, _lib.ptr(df),
                                                 _lib.stream()), "mdir_pair_diff_f64")
        S = gemm_f64(df, df, False, alpha=1.0 / n_pairs)                       # np.dot(df, df.T) / n
        L = _cholesky(S)
        Pc = torch.linalg.solve_triangular(L, torch.eye(D, dtype=torch.float64, device=dev), upper=False)   # inv(cholesky(S))
        # np.dot(df, df.T) with df = Pc (X - m)  ==  Pc [(X - m)(X - m)^T] Pc^T: one N-long contraction, then D^3 work
        Cx = gemm_f64(X, X, False, a_sub=m, b_sub=m)
        T = gemm_f64(Pc, Cx, True)
        Dm = gemm_f64(T, Pc, False)
        _, eigvec = _sorted_eigh(Dm)
        P = gemm_f64(eigvec.t().contiguous(), Pc, True)                        # np.dot(eigvec.T, P)
        return m.reshape(D, 1).cpu().numpy(), P.cpu().numpy()


def pcawhitenlearn(X, shrink=None, device="cuda"):
    """PCA whitening without annotations (whiten.py:14-35)."""
    dev = torch.device(device)
    with torch.cuda.device(dev):
        X = _dev_f64(X, dev)
        D, N = X.shape
        m = _cols_mean(X)
        Xcov = gemm_f64(X, X, False, alpha=1.0 / N, a_sub=m, b_sub=m)          # (Xcov + Xcov.T) / (2N) after symmetrisation
        eigval, eigvec = _sorted_eigh(Xcov)
        if shrink:
            b = eigval[shrink - 1]
            eigval = (1 - b) * eigval + b
        P = eigvec.t() * torch.rsqrt(eigval).reshape(D, 1)                     # inv(sqrt(diag(eigval))) @ eigvec.T
        return m.reshape(D, 1).cpu().numpy(), P.contiguous().cpu().numpy()
