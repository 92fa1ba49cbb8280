"""Seeded synthetic inputs shared by oracle/make_golden.py and tests/ (TEST
INFRASTRUCTURE ONLY).  Everything is np.random.RandomState (legacy, bit-stable
across numpy versions) so the GPU box regenerates exactly what the golden
outputs were computed from."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- feature maps -----------------------------------------------------------

POOL_SHAPES = [(1, 64, 7, 5), (2, 128, 23, 17), (1, 256, 16, 12), (1, 32, 1, 1), (1, 96, 32, 24), (3, 40, 3, 11)]
POOL_PS = [3.0, 2.9137, 1.0, 5.5]


def fmap(shape, seed, kind="relu"):
    """Backbone-like activations: relu(randn) (half zeros), with a few negatives
    and exact zeros kept so that clamp(min=eps) matters."""
    rs = np.random.RandomState(seed)
    x = rs.randn(*shape).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0.0) * np.float32(2.0)
    elif kind == "signed":
        pass
    elif kind == "zeros":
        x[:] = 0.0
    return x


# ---- CLAHE images ------------------------------------------------------------

CLAHE_SMALL = [(64, 64), (9, 9), (17, 23), (127, 93), (200, 150), (8, 8), (40, 333)]
CLAHE_LARGE = [(767, 1023), (1024, 725), (768, 1024), (683, 1024)]
CLAHE_DISTS = ["uniform", "gamma", "flat", "gradient", "bimodal"]


def image_u8(hw, dist, seed):
    H, W = hw
    rs = np.random.RandomState(seed)
    if dist == "uniform":
        img = rs.randint(0, 256, size=(H, W))
    elif dist == "gamma":      # night-like: most mass near 0
        img = (rs.rand(H, W) ** 4.0) * 255.0
    elif dist == "flat":
        img = np.full((H, W), 37)
    elif dist == "gradient":
        yy, xx = np.mgrid[0:H, 0:W]
        img = (yy * 255.0 / max(H - 1, 1) * 0.5 + xx * 255.0 / max(W - 1, 1) * 0.5) + rs.randint(0, 3, size=(H, W))
    elif dist == "bimodal":
        m = rs.rand(H, W) < 0.5
        img = np.where(m, rs.normal(40, 10, (H, W)), rs.normal(200, 15, (H, W)))
    else:
        raise KeyError(dist)
    return np.clip(img, 0, 255).astype(np.uint8)


def clahe_cases():
    """(key, hw, dist, clip, seed) for the golden CLAHE set."""
    out = []
    seed = 1000
    for hw in CLAHE_SMALL + CLAHE_LARGE:
        for dist in CLAHE_DISTS:
            clips = [4] if dist in ("flat", "gradient", "bimodal") else [4, 2, 40]
            for clip in clips:
                seed += 1
                out.append(("%dx%d_%s_c%d" % (hw[0], hw[1], dist, clip), hw, dist, clip, seed))
    return out


# ---- descriptors / search ----------------------------------------------------

def descriptors(n, d, seed, clusters=0):
    """L2-normalised rows (n, d) fp32; optional Gaussian clusters so neighbours mean something."""
    rs = np.random.RandomState(seed)
    if clusters:
        cent = rs.randn(clusters, d).astype(np.float32)
        x = cent[rs.randint(0, clusters, size=n)] + 0.7 * rs.randn(n, d).astype(np.float32)
    else:
        x = rs.randn(n, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def planted_queries(db, nq, seed, noise=0.3):
    """Queries = perturbed db rows, so that mAP is non-trivial.  Returns (q, src_idx)."""
    rs = np.random.RandomState(seed)
    src = rs.choice(db.shape[0], size=nq, replace=False)
    q = db[src] + noise * rs.randn(nq, db.shape[1]).astype(np.float32) / np.sqrt(db.shape[1])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32), src


def gnd_okjunk(n_db, nq, seed, n_ok=12, n_junk=6, empty_every=0):
    rs = np.random.RandomState(seed)
    gnd = []
    for i in range(nq):
        perm = rs.permutation(n_db)
        ok = perm[:n_ok]
        if empty_every and i % empty_every == empty_every - 1:
            ok = perm[:0]
        gnd.append({"ok": np.sort(ok).tolist(), "junk": np.sort(perm[n_ok:n_ok + n_junk]).tolist()})
    return gnd


def gnd_emh(n_db, nq, seed, n_easy=8, n_hard=10, n_junk=6):
    rs = np.random.RandomState(seed)
    gnd = []
    for i in range(nq):
        perm = rs.permutation(n_db)
        e = perm[:n_easy] if i % 7 != 3 else perm[:0]          # some queries have no easy positives
        gnd.append({"easy": np.sort(e).tolist(),
                    "hard": np.sort(perm[n_easy:n_easy + n_hard]).tolist(),
                    "junk": np.sort(perm[n_easy + n_hard:n_easy + n_hard + n_junk]).tolist()})
    return gnd


def lw(d, seed):
    """Random Lw {m (D,1), P (D,D)} float64 shaped like whitenlearn's output: rows of
    P scaled by decreasing 1/sqrt(eigenvalue)-like factors."""
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.randn(d, d))
    scale = 1.0 / np.sqrt(np.linspace(1.0, 0.05, d))
    P = (scale[:, None] * q)
    m = 0.02 * rs.randn(d, 1)
    return {"m": m.astype(np.float64), "P": P.astype(np.float64)}
