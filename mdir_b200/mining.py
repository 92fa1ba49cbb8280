"""Hard-negative mining for training tuples on the device.

Reference: ``TuplesDataset.create_epoch_tuples`` -- mdir/external/cirtorch/datasets/traindataset.py:242-267:
``torch.mm`` + ``torch.sort(dim=0, descending=True)`` over (pool x queries), then a Python walk down
every query's ranking that keeps at most one image per cluster and never the query's own cluster.
Here: 3xTF32 tensor-core scores + segmented radix sort (``Index.ranks``) + ``mdir_mine_negatives``
(one warp per query) + ``mdir_pair_l2dist``; nothing leaves the device until the final lists.
"""
import numpy as np
import torch

from . import _lib
from .search import Index, _as_dev_f32


def mine_hard_negatives(qvecs, poolvecs, qclusters, poolclusters, nnum, precision="fp32", device="cuda"):
    """qvecs (D, n_q), poolvecs (D, n_pool): fp32 descriptor matrices, one COLUMN per image, as the
    reference holds them; qclusters (n_q,), poolclusters (n_pool,): integer cluster ids.
    Returns (pos (n_q, nnum) int64 pool positions in selection order, ndist (n_q, nnum) fp32) on
    the device.  Raises IndexError when some query cannot find nnum negatives (the reference's
    ``ranks[r, q]`` runs off the end of the pool the same way, traindataset.py:258)."""
    lib = _lib.lib()
    dev = torch.device(device)
    nnum = int(nnum)
    index = Index(poolvecs, dxn=True, device=dev, keep_fp32=True)
    q32 = _as_dev_f32(qvecs, dev).t().contiguous()
    n_q, n_pool = q32.shape[0], index.n
    with torch.cuda.device(dev):
        pos = torch.empty((n_q, nnum), dtype=torch.int64, device=dev)
        dist = torch.empty((n_q, nnum), dtype=torch.float32, device=dev)
        if n_q == 0 or nnum == 0:
            return pos, dist
        qc = torch.as_tensor(np.asarray(qclusters, dtype=np.int32)).to(dev)
        pc = torch.as_tensor(np.asarray(poolclusters, dtype=np.int32)).to(dev)
        if qc.numel() != n_q or pc.numel() != n_pool:
            raise _lib.MdirError("cluster id arrays must have one entry per query / pool image")
        ranks = index.ranks(q32, precision=precision)                       # (n_pool, n_q) int64
        found = torch.empty((n_q,), dtype=torch.int32, device=dev)
        _lib.check(lib.mdir_mine_negatives(_lib.ptr(ranks), n_q, n_pool, n_q, _lib.ptr(pc), _lib.ptr(qc), nnum, _lib.ptr(pos),
                                           _lib.ptr(found), _lib.stream()), "mdir_mine_negatives")
        _lib.check(lib.mdir_pair_l2dist(_lib.ptr(q32), _lib.ptr(index.db32), _lib.ptr(pos), n_q, nnum, index.D, 1e-6, _lib.ptr(dist),
                                        _lib.stream()), "mdir_pair_l2dist")
        if int(found.min().item()) < nnum:
            raise IndexError("negative pool exhausted: a query found fewer than %d clusters to draw negatives from" % nnum)
    return pos, dist


def search_hard_negatives(qvecs, poolvecs, qidxs, idxs2images, clusters, nnum, precision="fp32", device="cuda"):
    """The block traindataset.py:242-267 as one call, in the reference's own variables: returns
    (nidxs, ndist_acc) with nidxs[q] = list of nnum IMAGE indices (idxs2images[pool position]) and
    ndist_acc the flat list of their L2 distances in selection order."""
    idxs2images = np.asarray(torch.as_tensor(idxs2images).cpu().numpy(), dtype=np.int64)
    clusters = np.asarray(clusters)
    qc = clusters[np.asarray(qidxs, dtype=np.int64)]
    pc = clusters[idxs2images]
    pos, dist = mine_hard_negatives(qvecs, poolvecs, qc, pc, nnum, precision=precision, device=device)
    pos = pos.cpu().numpy()
    nidxs = [[int(idxs2images[p]) for p in row] for row in pos]
    return nidxs, [float(x) for x in dist.cpu().numpy().reshape(-1)]
