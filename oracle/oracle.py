"""CPU oracle for mdir's post-backbone retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the reference algorithm for every row
of SURVEY.md section 8a.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``mdir_b200/`` imports it, and the
product path raises when the CUDA library is missing instead of falling back here.

Pinning status
--------------
* The reference ships no tests, fixtures or golden vectors.
  The oracle is pinned against **outputs of the reference itself run in the build
  container** (``oracle/make_golden.py`` imports /root/reference live and writes
  ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` re-checks every function
  here against those fixtures, and CLAHE additionally against the installed
  ``cv2`` wheel at run time (cv2 is the arbiter named by the reference,
  ``mdir/components/data/transform/functional.py:114``).
* alpha-QE and DBA are **not in the reference** (SURVEY.md App. E): ``alpha_qe``
  and ``dba`` below are restatements of the published definitions and are
  "parity unpinned".

All ``file:line`` citations are relative to /root/reference.
"""
import numpy as np

# ----------------------------------------------------------------------------
# pooling / normalisation     mdir/external/cirtorch/layers/functional.py
# ----------------------------------------------------------------------------


def mac(x):
    """functional.py:11-12 -- global max over (h, w).  x: (N,C,h,w) -> (N,C,1,1)."""
    x = np.asarray(x, dtype=np.float32)
    return x.max(axis=(2, 3), keepdims=True)


def spoc(x):
    """functional.py:16-17 -- global mean over (h, w)."""
    x = np.asarray(x, dtype=np.float32)
    return x.mean(axis=(2, 3), keepdims=True, dtype=np.float64).astype(np.float32)


def gem(x, p=3.0, eps=1e-6, dtype=np.float64):
    """functional.py:21-22 -- avg_pool2d(x.clamp(min=eps).pow(p), (h,w)).pow(1/p).

    Evaluated in float64 by default and rounded to fp32 once: the reference is
    fp32 ATen whose distance from this is 3.9e-7 max rel, well
    inside the 1e-5 gate, and fp64 gives an order-independent target.
    """
    x = np.asarray(x, dtype=np.float32).astype(dtype)
    p = dtype(p)
    xc = np.maximum(x, dtype(np.float32(eps)))
    m = np.power(xc, p).mean(axis=(2, 3), keepdims=True)
    return np.power(m, dtype(1.0) / p).astype(np.float32)


def l2n(x, eps=1e-6):
    """functional.py:130-131 -- x / (||x||_2 over dim=1 + eps); eps is ADDED to the norm."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    return (x / (n + eps)).astype(np.float32)


POOLING = {"mac": mac, "spoc": spoc, "gem": gem}


def net_tail(fmap, pooling="gem", p=3.0, eps=1e-6):
    """ImageRetrievalNet.forward tail, networks/imageretrievalnet.py:107-115 with
    ``whiten is None``: norm(pool(o)).squeeze(-1).squeeze(-1).permute(1,0).
    fmap (N,C,h,w) -> (C,N)."""
    if pooling == "gem":
        o = gem(fmap, p, eps)
    else:
        o = POOLING[pooling](fmap)
    o = l2n(o)[:, :, 0, 0]
    return np.ascontiguousarray(o.T)


# ----------------------------------------------------------------------------
# multi-scale aggregation / Lw whitening      mdir/components/data/wrapper.py
# ----------------------------------------------------------------------------


def aggregate_tensor(tensors, nscales, outputdim, msp):
    """CirMultiscaleAggregation.aggregate_tensor, wrapper.py:109-119.
    v = sum_s o_s^msp ; v = (v/S)^(1/msp) ; v /= ||v||   (NO eps)."""
    assert len(tensors) == nscales, "%s != %s" % (len(tensors), nscales)
    v = np.zeros(outputdim, dtype=np.float64)
    for t in tensors:
        v += np.power(np.asarray(t, dtype=np.float32).astype(np.float64).reshape(-1), float(msp))
    v = np.power(v / nscales, 1.0 / float(msp))
    v = v / np.sqrt((v * v).sum())
    return v.astype(np.float32)


def multiscale_msp(nscales, pooling, regional, whitening, p):
    """The msp rule, wrapper.py:121-124."""
    if nscales > 1 and pooling == "gem" and not regional and not whitening:
        return float(p)
    return 1.0


def cirwhiten_postprocess(v, m, P, dimensions=None):
    """CirtorchWhiten.postprocess, wrapper.py:193-195: P,m are cast to fp32 at
    construction (wrapper.py:188-189); X = P[:dims] @ (v - m); X / (||X|| + 1e-6).
    v (D,) -> (dims,)."""
    P = np.asarray(P, dtype=np.float32)
    m = np.asarray(m, dtype=np.float32).reshape(-1)
    dimensions = dimensions or P.shape[0]
    x = (np.asarray(v, dtype=np.float32).reshape(-1) - m).astype(np.float64)
    X = P[:dimensions].astype(np.float64) @ x
    X = X / (np.sqrt((X * X).sum()) + 1e-6)
    return X.astype(np.float32)


def whitenapply(X, m, P, dimensions=None):
    """cirtorch/utils/whiten.py:4-12 (numpy, dtype follows the inputs; fp64 Lw)."""
    if not dimensions:
        dimensions = P.shape[0]
    X = np.dot(P[:dimensions, :], X - m)
    X = X / (np.linalg.norm(X, ord=2, axis=0, keepdims=True) + 1e-6)
    return X


def gem_head(fmaps, p, eps, m=None, P=None, dimensions=None, pooling="gem",
             regional=False, whitening=False):
    """The whole eval-stage head for ONE image: per scale
    net_tail -> aggregate_tensor (msp rule) -> CirtorchWhiten.postprocess.
    fmaps: list of S arrays (1,C,h_s,w_s).  Returns (dims,) fp32."""
    outs = [net_tail(f, pooling, p, eps)[:, 0] for f in fmaps]
    S = len(outs)
    C = outs[0].shape[0]
    if S > 1:
        v = aggregate_tensor(outs, S, C, multiscale_msp(S, pooling, regional, whitening, p))
    else:
        # wrapper.py:96-98,126-127: a single scale still goes through aggregate_tensor
        v = aggregate_tensor(outs, 1, C, 1.0)
    if P is not None:
        v = cirwhiten_postprocess(v, m, P, dimensions)
    return v


# ----------------------------------------------------------------------------
# CLAHE (8UC1)   cv2.createCLAHE as called at transform/functional.py:109-117
# ----------------------------------------------------------------------------


def _reflect101(idx, n):
    """cv2 BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    idx = np.asarray(idx).copy()
    if n == 1:
        return np.zeros_like(idx)
    for _ in range(64):
        lo = idx < 0
        hi = idx >= n
        if not (lo.any() or hi.any()):
            break
        idx[lo] = -idx[lo]
        idx[hi] = 2 * (n - 1) - idx[hi]
    return idx


def _round_half_even_u8(x):
    """cvRound (round-half-to-even) followed by saturate_cast<uchar>."""
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def clahe_luts(src, clip=4.0, tiles_x=8, tiles_y=8):
    """Steps 1-3 of SURVEY.md App. A (OpenCV modules/imgproc/src/clahe.cpp,
    CLAHE_CalcLut_Body; the source is a third-party dependency not vendored in
    /root/reference -- opencv-python is unpinned in requirements.txt:3, the
    installed wheel 4.13.0 is the arbiter).  Returns (lut[ty,tx,256] u8, tw, th)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape
    if W % tiles_x == 0 and H % tiles_y == 0:
        ext = src
    else:
        pb = tiles_y - (H % tiles_y)
        pr = tiles_x - (W % tiles_x)
        ys = _reflect101(np.arange(H + pb), H)
        xs = _reflect101(np.arange(W + pr), W)
        ext = src[np.ix_(ys, xs)]
    tw = ext.shape[1] // tiles_x
    th = ext.shape[0] // tiles_y
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip_limit = 0
    if clip > 0.0:
        clip_limit = max(int(float(clip) * area / 256.0), 1)
    luts = np.zeros((tiles_y, tiles_x, 256), dtype=np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            tile = ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist = np.bincount(tile.reshape(-1), minlength=256).astype(np.int64)
            if clip_limit > 0:
                clipped = int(np.maximum(hist - clip_limit, 0).sum())
                hist = np.minimum(hist, clip_limit)
                redist = clipped // 256
                residual = clipped - redist * 256
                hist = hist + redist
                if residual != 0:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            csum = np.cumsum(hist).astype(np.float32)
            luts[ty, tx] = _round_half_even_u8(csum * lut_scale)
    return luts, tw, th


def clahe_u8(src, clip=4.0, tiles_x=8, tiles_y=8):
    """cv2.createCLAHE(clipLimit=clip, tileGridSize=(tiles_x, tiles_y)).apply(src),
    bit-exact (CLAHE_Interpolation_Body; every fp32
    multiply and add rounded individually, this exact association)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape
    luts, tw, th = clahe_luts(src, clip, tiles_x, tiles_y)
    f32 = np.float32
    inv_tw = f32(1.0) / f32(tw)
    inv_th = f32(1.0) / f32(th)

    def axis(n, inv, ntiles):
        tf = (np.arange(n, dtype=np.float32) * inv).astype(np.float32) - f32(0.5)
        t1 = np.floor(tf).astype(np.int32)
        a = (tf - t1.astype(np.float32)).astype(np.float32)
        a1 = (f32(1.0) - a).astype(np.float32)
        t2 = np.minimum(t1 + 1, ntiles - 1)
        t1 = np.maximum(t1, 0)
        return t1, t2, a, a1

    tx1, tx2, xa, xa1 = axis(W, inv_tw, tiles_x)
    ty1, ty2, ya, ya1 = axis(H, inv_th, tiles_y)
    v = src.astype(np.intp)
    Y1 = ty1[:, None]
    Y2 = ty2[:, None]
    X1 = tx1[None, :]
    X2 = tx2[None, :]
    l11 = luts[Y1, X1, v].astype(np.float32)
    l12 = luts[Y1, X2, v].astype(np.float32)
    l21 = luts[Y2, X1, v].astype(np.float32)
    l22 = luts[Y2, X2, v].astype(np.float32)
    xa_ = xa[None, :]
    xa1_ = xa1[None, :]
    top = ((l11 * xa1_).astype(np.float32) + (l12 * xa_).astype(np.float32)).astype(np.float32)
    bot = ((l21 * xa1_).astype(np.float32) + (l22 * xa_).astype(np.float32)).astype(np.float32)
    res = ((top * ya1[:, None]).astype(np.float32) + (bot * ya[:, None]).astype(np.float32)).astype(np.float32)
    return _round_half_even_u8(res)


def channel_clahe(chan, clip_limit=4, grid_size=8):
    """ChannelClahe.apply, transform/functional.py:109-117:
    u8 = (chan*255).astype(uint8) [C truncation]; CLAHE; .astype(float32)/255.0"""
    g = (int(grid_size), int(grid_size)) if not isinstance(grid_size, tuple) else grid_size
    q = (np.asarray(chan, dtype=np.float32) * 255).astype(np.uint8)
    return clahe_u8(q, float(int(clip_limit)), g[0], g[1]).astype(np.float32) / 255.0


def cv2_lab_lattice():
    """The 33^3 table OpenCV's float RGB->Lab interpolates (color_lab.cpp, RGB2Lab_f with
    useInterpolation; third-party, opencv-python unpinned in requirements.txt:3 -- the installed wheel is
    the arbiter): feeding cv2 the lattice points returns the entries exactly."""
    import cv2
    g = (np.arange(33) / 32.0).astype(np.float32)
    R, G, B = np.meshgrid(g, g, g, indexing="ij")
    lab = cv2.cvtColor(np.stack([R, G, B], -1).reshape(-1, 1, 3), cv2.COLOR_RGB2LAB).reshape(33, 33, 33, 3).astype(np.float64)
    return np.stack([np.rint(lab[..., 0] / 100 * 16384), np.rint((lab[..., 1] + 128) / 256 * 16384),
                     np.rint((lab[..., 2] + 128) / 256 * 16384)], -1).astype(np.int64)


def rgb2lab_cv(img, lattice):
    """cv2.cvtColor(float32 RGB, COLOR_RGB2LAB) restated (RGB2Lab_f::operator() + trilinearInterpolate):
    iR = cvRound(clip(R)*2^14); cell = iR >> 9; 4-bit weights (iR & 511) >> 5; CV_DESCALE(sum, 12)."""
    f32 = np.float32
    x = np.clip(np.asarray(img, dtype=f32), 0, 1)
    ii = np.rint(x * f32(16384)).astype(np.int64)
    t = ii >> 9
    fr = (ii & 511) >> 5
    t1 = np.minimum(t + 1, 32)
    acc = np.zeros(x.shape[:-1] + (3,), np.int64)
    for dx in (0, 1):
        wx = fr[..., 0] if dx else 16 - fr[..., 0]
        tx = t1[..., 0] if dx else t[..., 0]
        for dy in (0, 1):
            wy = fr[..., 1] if dy else 16 - fr[..., 1]
            ty = t1[..., 1] if dy else t[..., 1]
            for dz in (0, 1):
                wz = fr[..., 2] if dz else 16 - fr[..., 2]
                tz = t1[..., 2] if dz else t[..., 2]
                acc += lattice[tx, ty, tz] * (wx * wy * wz)[..., None]
    acc = (acc + 2048) >> 12
    o = acc.astype(f32) / f32(16384)
    return np.stack([o[..., 0] * f32(100), o[..., 1] * f32(256) - f32(128), o[..., 2] * f32(256) - f32(128)], -1).astype(f32)


def _spline_build(fv):
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


def lab2rgb_cv(lab):
    """cv2.cvtColor(float32 Lab, COLOR_LAB2RGB) restated (Lab2RGBfloat + splineInterpolate of the sRGB
    gamma table); agrees with the wheel to ~6e-6 (FMA contraction / table rounding)."""
    f32 = np.float32
    lab = np.asarray(lab, dtype=f32)
    L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
    lT = f32(0.008856) * f32(903.3)
    fT = f32(7.787) * f32(0.008856) + f32(16.0) / f32(116.0)
    y_lo = (L / f32(903.3)).astype(f32)
    fy_lo = (f32(7.787) * y_lo + f32(16.0) / f32(116.0)).astype(f32)
    fy_hi = ((L + f32(16.0)) / f32(116.0)).astype(f32)
    y_hi = (fy_hi * fy_hi * fy_hi).astype(f32)
    lo = L <= lT
    y = np.where(lo, y_lo, y_hi)
    fy = np.where(lo, fy_lo, fy_hi)

    def inv(f):
        return np.where(f <= fT, ((f - f32(16.0) / f32(116.0)) / f32(7.787)).astype(f32), (f * f * f).astype(f32))

    X = inv((a / f32(500.0) + fy).astype(f32))
    Z = inv((fy - b / f32(200.0)).astype(f32))
    Mi = np.array([[3.240479, -1.53715, -0.498535], [-0.969256, 1.875991, 0.041556], [0.055648, -0.204043, 1.057311]])
    C = (Mi * np.array([0.950456, 1.0, 1.088754])[None, :]).astype(f32)
    rgb = np.stack([C[i, 0] * X + C[i, 1] * y + C[i, 2] * Z for i in range(3)], -1).astype(f32)
    rgb = np.clip(rgb, 0, 1).astype(f32)
    xg = np.arange(1025) / 1024.0
    tab = _spline_build(np.where(xg <= 0.0031308, xg * 12.92, 1.055 * np.power(xg, 1 / 2.4) - 0.055)).astype(f32)
    xs = (rgb * f32(1024)).astype(f32)
    ix = np.clip(xs.astype(np.int32), 0, 1023)
    tt = (xs - ix.astype(f32)).astype(f32)
    c = tab[ix]
    return (((c[..., 3] * tt + c[..., 2]) * tt + c[..., 1]) * tt + c[..., 0]).astype(f32)


def image_clahe(img, clip_limit=4, grid_size=8, lattice=None):
    """ImageClahe.apply with colorspace 'lab' (transform/functional.py:120-129 over :24-48):
    spc = (RGB2LAB(img) + [0,128,128]) / [100,255,255]; spc[...,0] = ChannelClahe(spc[...,0]);
    LAB2RGB(spc * [100,255,255] - [0,128,128])."""
    f32 = np.float32
    lattice = cv2_lab_lattice() if lattice is None else lattice
    lab = rgb2lab_cv(img, lattice)
    spc = ((lab + np.array([0, 128, 128], f32)) / np.array([100.0, 255.0, 255.0], f32)).astype(f32)
    spc[..., 0] = channel_clahe(spc[..., 0], clip_limit, grid_size)
    return lab2rgb_cv((spc * np.array([100.0, 255.0, 255.0], f32) - np.array([0, 128, 128], f32)).astype(f32))


# ----------------------------------------------------------------------------
# similarity / ranks     mdir/components/optim/score/cirscore.py:65-70
# ----------------------------------------------------------------------------


def scores(vecs, qvecs):
    """cirscore.py:69 -- np.dot(vecs.T, qvecs); vecs (D,N_db), qvecs (D,N_q) fp32
    -> (N_db, N_q) fp32."""
    return np.dot(np.asarray(vecs).T, np.asarray(qvecs))


def ranks_from_scores(sc):
    """cirscore.py:70 -- np.argsort(-scores, axis=0), made canonical with
    kind='stable' => ties broken by ascending db index (numpy's
    default introsort leaves tie order unspecified; the stable order is one valid
    reference output and the one north_star names)."""
    return np.argsort(-np.asarray(sc), axis=0, kind="stable")


def ranks(vecs, qvecs):
    return ranks_from_scores(scores(vecs, qvecs))


def topk_from_scores(sc, k):
    """First k rows of ranks_from_scores plus the scores at those ranks.
    -> (idx (k,N_q) int64, val (k,N_q) fp32)."""
    r = ranks_from_scores(sc)[:k]
    return r, np.take_along_axis(np.asarray(sc), r, axis=0)


# ----------------------------------------------------------------------------
# mAP     mdir/external/cirtorch/utils/evaluate.py
# ----------------------------------------------------------------------------


def compute_ap(pos_ranks, nres):
    """evaluate.py:3-37 -- trapezoidal AP over zero-based ranks of positives."""
    ap = 0.0
    recall_step = 1.0 / nres
    for j in range(len(pos_ranks)):
        rank = int(pos_ranks[j])
        precision_0 = 1.0 if rank == 0 else float(j) / rank
        precision_1 = float(j + 1) / (rank + 1)
        ap += (precision_0 + precision_1) * recall_step / 2.0
    return ap


def compute_map(rk, gnd, kappas=()):
    """evaluate.py:39-111.  rk: (N_db, N_q) integer ranks; gnd: list of
    {'ok': [...], 'junk': [...]}.  Returns (map, aps, pr, prs)."""
    rk = np.asarray(rk)
    nq = len(gnd)
    aps = np.zeros(nq)
    pr = np.zeros(len(kappas))
    prs = np.zeros((nq, len(kappas)))
    nempty = 0
    mp = 0.0
    for i in range(nq):
        qgnd = np.array(gnd[i]["ok"])
        if qgnd.shape[0] == 0:
            aps[i] = float("nan")
            prs[i, :] = float("nan")
            nempty += 1
            continue
        qgndj = np.array(gnd[i]["junk"]) if "junk" in gnd[i] else np.empty(0)
        col = rk[:, i]
        pos = np.arange(rk.shape[0])[np.isin(col, qgnd)]
        junk = np.arange(rk.shape[0])[np.isin(col, qgndj)]
        k = 0
        ij = 0
        if len(junk):
            ip = 0
            while ip < len(pos):
                while ij < len(junk) and pos[ip] > junk[ij]:
                    k += 1
                    ij += 1
                pos[ip] = pos[ip] - k
                ip += 1
        ap = compute_ap(pos, len(qgnd))
        mp += ap
        aps[i] = ap
        pos = pos + 1
        for j in range(len(kappas)):
            kq = min(max(pos), kappas[j])
            prs[i, j] = (pos <= kq).sum() / kq
        pr = pr + prs[i, :]
    mp = mp / (nq - nempty)
    pr = pr / (nq - nempty)
    return mp, aps, pr, prs


def compute_map_emh(rk, gnd, kappas=(1, 5, 10)):
    """The roxford5k/rparis6k branch of compute_map_and_print, evaluate.py:123-152:
    Easy / Medium / Hard regrouping of easy/hard/junk."""
    def regroup(ok_keys, junk_keys):
        out = []
        for g in gnd:
            out.append({"ok": np.concatenate([np.asarray(g[k], dtype=np.int64) for k in ok_keys]),
                        "junk": np.concatenate([np.asarray(g[k], dtype=np.int64) for k in junk_keys])})
        return out
    mE, aE, pE, _ = compute_map(rk, regroup(["easy"], ["junk", "hard"]), kappas)
    mM, aM, pM, _ = compute_map(rk, regroup(["easy", "hard"], ["junk"]), kappas)
    mH, aH, pH, _ = compute_map(rk, regroup(["hard"], ["junk", "easy"]), kappas)
    return ({"map_easy": mE, "map_medium": mM, "map_hard": mH},
            {"ap_easy": aE, "ap_medium": aM, "ap_hard": aH},
            {"mpr_easy": pE, "mpr_medium": pM, "mpr_hard": pH})


# ----------------------------------------------------------------------------
# alpha-QE / DBA -- NOT in the reference; parity unpinned
# ----------------------------------------------------------------------------


def alpha_qe(db, q, alpha=3.0, n_qe=10):
    """Radenovic et al. TPAMI'18 alpha query expansion.  db (N,D), q (Nq,D) fp32
    rows L2-normalised.  q' = q + sum_{i<=n_qe} max(s_i,0)^alpha x_i ; q' /= ||q'||."""
    db = np.asarray(db, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    sc = db @ q.T                                   # (N, Nq)
    idx, val = topk_from_scores(sc, n_qe)           # (n_qe, Nq)
    out = np.empty_like(q, dtype=np.float64)
    for j in range(q.shape[0]):
        w = np.power(np.maximum(val[:, j].astype(np.float64), 0.0), alpha)
        out[j] = q[j].astype(np.float64) + (w[:, None] * db[idx[:, j]].astype(np.float64)).sum(0)
        out[j] /= np.sqrt((out[j] ** 2).sum())
    return out.astype(np.float32)


def dba(db, alpha=3.0, k_dba=10, chunk=4096):
    """Database-side augmentation: every db vector replaced by the alpha-weighted
    sum of its own top-k_dba neighbours in the db (self included, weight s^alpha),
    then re-normalised."""
    db = np.asarray(db, dtype=np.float32)
    out = np.empty(db.shape, dtype=np.float64)
    for s in range(0, db.shape[0], chunk):
        blk = db[s:s + chunk]
        sc = db @ blk.T
        idx, val = topk_from_scores(sc, k_dba)
        w = np.power(np.maximum(val.astype(np.float64), 0.0), alpha)   # (k, B)
        acc = np.einsum("kb,kbd->bd", w, db[idx].astype(np.float64))
        out[s:s + chunk] = acc / np.sqrt((acc ** 2).sum(1, keepdims=True))
    return out.astype(np.float32)


# ----------------------------------------------------------------------------
# f4: hard-negative mining and whitening learning
# ----------------------------------------------------------------------------


def mine_negatives(qvecs, poolvecs, qclusters, poolclusters, nnum):
    """cirtorch/datasets/traindataset.py:242-267.  qvecs (D,Nq), poolvecs (D,Np) fp32 (columns =
    images); cluster ids per query / pool image.  Walk every query's ranking (best score first),
    keep the first nnum pool positions whose cluster is neither the query's nor already kept.
    Returns (pos (Nq,nnum) int64 pool positions, ndist (Nq,nnum) fp32) with
    ndist = sqrt(sum((q - p + 1e-6)**2)) (traindataset.py:263)."""
    qvecs = np.asarray(qvecs, dtype=np.float32)
    poolvecs = np.asarray(poolvecs, dtype=np.float32)
    rk = ranks_from_scores(np.dot(poolvecs.T, qvecs))
    nq = qvecs.shape[1]
    pos = np.full((nq, nnum), -1, dtype=np.int64)
    ndist = np.full((nq, nnum), np.nan, dtype=np.float32)
    for q in range(nq):
        used = [qclusters[q]]
        r = 0
        n = 0
        while n < nnum:
            cand = rk[r, q]                      # IndexError when the pool runs out, like the reference
            if poolclusters[cand] not in used:
                used.append(poolclusters[cand])
                pos[q, n] = cand
                d = qvecs[:, q] - poolvecs[:, cand] + np.float32(1e-6)
                ndist[q, n] = np.sqrt(np.sum(d * d, dtype=np.float32))
                n += 1
            r += 1
    return pos, ndist


def _eig_desc(S):
    """Eigenpairs by descending eigenvalue (whiten.py:25-28, 46-49; the reference calls the general
    np.linalg.eig on a symmetric matrix -- same pairs up to sign and rounding)."""
    w, v = np.linalg.eigh((S + S.T) * 0.5)
    return w[::-1], v[:, ::-1]


def whitenlearn(X, qidxs, pidxs):
    """cirtorch/utils/whiten.py:37-53 (Lw whitening from matching pairs), fp64.  X (D,N)."""
    X = np.asarray(X, dtype=np.float64)
    m = X[:, qidxs].mean(axis=1, keepdims=True)
    df = X[:, qidxs] - X[:, pidxs]
    S = df @ df.T / df.shape[1]
    alpha = 0.0                                   # whiten.py:55-70: jitter the diagonal until positive definite
    while True:
        try:
            L = np.linalg.cholesky(S + alpha * np.eye(S.shape[0]))
            break
        except np.linalg.LinAlgError:
            alpha = 1e-10 if alpha == 0 else alpha * 10
    Pc = np.linalg.inv(L)
    Y = Pc @ (X - m)
    _, vec = _eig_desc(Y @ Y.T)
    return m, vec.T @ Pc


def pcawhitenlearn(X, shrink=None):
    """cirtorch/utils/whiten.py:14-35, fp64."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[1]
    m = X.mean(axis=1, keepdims=True)
    Xc = X - m
    cov = Xc @ Xc.T
    w, v = _eig_desc((cov + cov.T) / (2 * N))
    if shrink:
        b = w[shrink - 1]
        w = (1 - b) * w + b
    return m, (v / np.sqrt(w)[None, :]).T


def whitening_rows_aligned(P, P_ref):
    """Rows of P with the sign of each flipped to agree with P_ref (eigenvector signs are arbitrary)."""
    s = np.sign(np.sum(P * P_ref, axis=1, keepdims=True))
    s[s == 0] = 1
    return P * s
