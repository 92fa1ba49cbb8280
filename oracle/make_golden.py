#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Usage:  python oracle/make_golden.py [section ...]   (needs /root/reference; cv2 4.13)
        sections: base (pooling/head/clahe/search), mining, whitenlearn, extract, transforms; default: all

Every array written here is an OUTPUT OF THE REFERENCE'S OWN CODE (mdir/cirtorch
functions, or the cv2/numpy calls at the reference's call sites) on the seeded
inputs of oracle/synth.py.  tests/test_oracle_golden.py checks oracle/oracle.py
against them; the -m gpu tests check the CUDA path against them.  The reference
tree cannot travel to the GPU box, these fixtures do.
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_import, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_mining():
    """Hard-negative mining: the LIVE TuplesDataset.create_epoch_tuples (cirtorch/datasets/traindataset.py:178-272)
    run on CPU over a synthetic image folder with a recording descriptor network."""
    import tempfile
    import torch
    import torch.nn as nn
    from PIL import Image
    from cirtorch.datasets.traindataset import TuplesDataset

    D, n_img, n_clu = 32, 400, 40
    rs = np.random.RandomState(31)
    tmp = tempfile.mkdtemp(prefix="mdir_mining_")
    base = rs.randint(0, 256, size=(n_clu, 8, 8, 3))
    clusters = [i % n_clu for i in range(n_img)]
    images = []
    for i in range(n_img):
        px = np.clip(base[clusters[i]] + rs.randint(-60, 61, size=(8, 8, 3)), 0, 255).astype(np.uint8)
        fn = os.path.join(tmp, "im%04d.png" % i)
        Image.fromarray(px).save(fn)
        images.append(fn)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(5)
            self.W = torch.randn(D, 192, generator=g)
            self.meta = {"out_channels": D}
            self.seen = []

        def forward(self, x):
            v = self.W @ (x.reshape(-1) - 0.5)
            v = v / v.norm()
            self.seen.append(v.clone())
            return v.reshape(1, D, 1, 1)

    def to_tensor(img):
        return torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1).contiguous()

    out = {}
    for case, (qsize, poolsize, nnum) in enumerate([(60, 300, 5), (25, 400, 8), (40, 120, 3)]):
        ds = TuplesDataset.__new__(TuplesDataset)
        ds.name, ds.mode, ds.imsize, ds.transform, ds.print_freq = "synthetic", "train", None, to_tensor, 1000
        ds.images, ds.clusters = images, clusters
        qp = list(range(0, 240, 2))
        ds.qpool = qp
        ds.ppool = [(q + n_clu) % n_img for q in qp]            # same cluster, other image
        ds.qsize, ds.poolsize, ds.nnum = qsize, poolsize, nnum
        net = Net()
        torch.manual_seed(100 + case)
        with contextlib.redirect_stdout(io.StringIO()):
            ret = ds.create_epoch_tuples(net, device=torch.device("cpu"))
        torch.manual_seed(100 + case)                           # replay the method's two randperm draws
        idxs2qpool = torch.randperm(len(ds.qpool))[:qsize]
        idxs2images = torch.randperm(len(images))[:poolsize]
        assert [ds.qpool[i] for i in idxs2qpool] == ds.qidxs
        vec = torch.stack(net.seen, 1).numpy()                  # (D, qsize + poolsize) in call order
        tag = "c%d_" % case
        out[tag + "qvecs"] = vec[:, :qsize].copy()
        out[tag + "poolvecs"] = vec[:, qsize:].copy()
        out[tag + "qclusters"] = np.array([clusters[i] for i in ds.qidxs], np.int32)
        out[tag + "poolclusters"] = np.array([clusters[int(i)] for i in idxs2images], np.int32)
        out[tag + "idxs2images"] = idxs2images.numpy().astype(np.int64)
        out[tag + "nidxs"] = np.array([[int(x) for x in row] for row in ds.nidxs], np.int64)
        out[tag + "ndist"] = np.array(ret["average_negative_distance"], np.float32)
        out[tag + "nnum"] = np.array(nnum)
    np.savez_compressed(os.path.join(OUT, "mining.npz"), **out)


def make_extract():
    """Descriptor extraction: the LIVE cirtorch extract_vectors (networks/imageretrievalnet.py:277-324: DataLoader over
    image files, extract_ss / extract_ms) on CPU with a small seeded ImageRetrievalNet.  The fixture keeps the network's
    weights, the exact tensors the loader fed to it, and the (D, N) results."""
    import tempfile
    import torch
    import torch.nn as nn
    from PIL import Image
    from cirtorch.layers.pooling import GeM
    from cirtorch.networks.imageretrievalnet import ImageRetrievalNet, extract_vectors
    from mdir.components.data.wrapper import CirMultiscaleAggregation, CirtorchWhiten

    torch.manual_seed(17)
    feats = [nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.ReLU(), nn.Conv2d(16, 48, 3, stride=2, padding=1), nn.ReLU()]
    meta = {"architecture": "tiny", "local_whitening": False, "pooling": "gem", "regional": False, "whitening": False,
            "mean": [0.485, 0.456, 0.406], "std": [0.229, 0.224, 0.225], "outputdim": 48, "out_channels": 48}
    p = 2.9137
    net = ImageRetrievalNet(feats, None, GeM(p=p), None, meta).eval()
    rs = np.random.RandomState(23)
    tmp = tempfile.mkdtemp(prefix="mdir_extract_")
    files = []
    for i, (h, w) in enumerate(((64, 48), (57, 91), (120, 96), (40, 40), (33, 77), (96, 128))):
        fn = os.path.join(tmp, "im%d.png" % i)
        Image.fromarray(rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)).save(fn)
        files.append(fn)
    mean, std = torch.tensor(meta["mean"]).view(3, 1, 1), torch.tensor(meta["std"]).view(3, 1, 1)

    def transform(img):
        x = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1).contiguous()
        return (x - mean) / std

    out = {"p": np.array(p)}
    for k, v in net.features.state_dict().items():
        out["w_" + k] = v.numpy()
    for i, fn in enumerate(files):
        out["input_%d" % i] = transform(Image.open(fn).convert("RGB")).numpy()
    ms = [1, 1 / np.sqrt(2), 1 / 2]
    with contextlib.redirect_stdout(io.StringIO()):
        out["vecs_ss"] = extract_vectors(net, files, None, transform, device=torch.device("cpu")).numpy()
        out["vecs_ms"] = extract_vectors(net, files, None, transform, ms=ms, msp=p, device=torch.device("cpu")).numpy()
        out["vecs_ms_msp1"] = extract_vectors(net, files, None, transform, ms=ms, msp=1, device=torch.device("cpu")).numpy()
    out["ms"] = np.array(ms)
    # mdir's own inference pattern: Compose([CirtorchWhiten, CirMultiscaleAggregation]) around the same network
    # (mdir/components/data/wrapper.py:8-36 -- preprocess in order, postprocess in reverse)
    from mdir.components.data.wrapper import Compose
    lwd = synth.lw(48, 29)
    out["lw_m"], out["lw_P"] = lwd["m"], lwd["P"]
    wh = CirtorchWhiten.__new__(CirtorchWhiten)
    wh.device = torch.device("cpu")
    wh.P = torch.tensor(lwd["P"], dtype=torch.float32)
    wh.m = torch.tensor(lwd["m"], dtype=torch.float32)
    wh.dimensions = 32
    comp = Compose([wh, CirMultiscaleAggregation(True, torch.device("cpu"))], torch.device("cpu"))
    with torch.no_grad():
        out["compose_wh32"] = np.stack([comp(torch.from_numpy(out["input_%d" % i]).unsqueeze(0), net, net).numpy() for i in range(len(files))])
    np.savez_compressed(os.path.join(OUT, "extract.npz"), **out)


def make_transforms():
    """The three CLAHE transform classes of the reference's TRANSFORMS registry (photometric_transforms.py:10-43), built
    from the string arguments of the "apply_clahe:4:lab:8" mini-language, on a float RGB picture whose sides are not
    multiples of the 8 x 8 grid."""
    from mdir.components.data.transform import TRANSFORMS
    rs = np.random.RandomState(91)
    pic = (rs.rand(61, 83, 3) ** 2.2).astype(np.float32)
    pic[:9, :11] = 0.0
    pic[20:30, 40:60] = 1.0
    out = {"pic": pic}
    out["apply_clahe_4_lab_8"] = TRANSFORMS["apply_clahe"]("4", "lab", "8")(pic.copy())[0]
    out["apply_clahe_2_lab_4"] = TRANSFORMS["apply_clahe"]("2", "lab", "4")(pic.copy())[0]
    two = TRANSFORMS["create_clahed"]()(pic.copy())
    assert np.array_equal(two[0], pic)
    out["create_clahed_1"] = two[1]
    out["add_clahe_fromrgb"] = TRANSFORMS["add_clahe_fromrgb"]()(pic.copy())[0]
    out["add_clahe_fromrgb_2_4"] = TRANSFORMS["add_clahe_fromrgb"]("2", "4")(pic.copy())[0]
    np.savez_compressed(os.path.join(OUT, "transforms.npz"), **out)


def make_whitenlearn():
    """Lw / PCA whitening learning: cirtorch/utils/whiten.py:14-53 (pure numpy, fp64)."""
    from cirtorch.utils.whiten import whitenlearn, pcawhitenlearn, whitenapply
    out = {}
    for case, (D, N, npair) in enumerate([(32, 600, 200), (128, 3000, 1200)]):
        X = synth.descriptors(N, D, 40 + case, clusters=30).T.astype(np.float64)      # (D, N), columns = images
        rs = np.random.RandomState(50 + case)
        qidxs = rs.randint(0, N, npair)
        pidxs = (qidxs + 30 * rs.randint(1, 5, npair)) % N       # arbitrary 'matching' pairs: only the algebra is under test
        m, P = whitenlearn(X, qidxs, pidxs)
        mp, Pp = pcawhitenlearn(X)
        tag = "c%d_" % case
        out[tag + "X_recipe"] = np.array([N, D, 40 + case, 30])    # X = synth.descriptors(N, D, seed, clusters).T as fp64
        out[tag + "qidxs"], out[tag + "pidxs"] = qidxs, pidxs
        out[tag + "m"], out[tag + "P"] = m, np.real(P)
        out[tag + "applied"] = whitenapply(X[:, :50], m, np.real(P))
        out[tag + "pca_m"], out[tag + "pca_P"] = mp, np.real(Pp)
    np.savez_compressed(os.path.join(OUT, "whitenlearn.npz"), **out)


def main():
    ref_import.import_reference()
    os.makedirs(OUT, exist_ok=True)
    sections = sys.argv[1:] or ["base", "mining", "whitenlearn", "extract", "transforms"]
    if "transforms" in sections:
        make_transforms()
    if "extract" in sections:
        make_extract()
    if "mining" in sections:
        make_mining()
    if "whitenlearn" in sections:
        make_whitenlearn()
    if "base" in sections:
        make_base()
    for f in sorted(os.listdir(OUT)):
        print("%-14s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


def make_base():
    import cv2
    import torch
    import torch.nn as nn
    from cirtorch.layers.pooling import GeM, MAC, SPoC
    from cirtorch.layers.normalization import L2N
    from cirtorch.networks.imageretrievalnet import ImageRetrievalNet
    from cirtorch.utils.whiten import whitenapply
    from cirtorch.utils.evaluate import compute_map, compute_map_and_print
    from mdir.components.data.wrapper import CirMultiscaleAggregation, CirtorchWhiten
    from mdir.components.data.transform.functional import ChannelClahe, ImageClahe

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # ---- 1. pooling + L2N (cirtorch/layers/pooling.py, normalization.py) -------
    pool = {}
    with torch.no_grad():
        for si, shape in enumerate(synth.POOL_SHAPES):
            for kind in ("relu", "signed", "zeros"):
                x = torch.from_numpy(synth.fmap(shape, 100 + si, kind))
                tag = "s%d_%s" % (si, kind)
                pool[tag + "_mac"] = MAC()(x).numpy()
                pool[tag + "_spoc"] = SPoC()(x).numpy()
                pool[tag + "_l2n_mac"] = L2N()(MAC()(x)).numpy()
                for p in synth.POOL_PS:
                    pool[tag + "_gem_p%g" % p] = GeM(p=p)(x).numpy()
                    pool[tag + "_l2n_gem_p%g" % p] = L2N()(GeM(p=p)(x)).numpy()
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), **pool)

    # ---- 2. ImageRetrievalNet tail + multiscale + Lw head --------------------------
    head = {}
    C = 128
    lwd = synth.lw(C, 7)
    head["lw_m"] = lwd["m"]
    head["lw_P"] = lwd["P"]
    scales_hw = [(32, 24), (23, 17), (16, 12)]
    with torch.no_grad():
        for pooling, mod in (("gem", lambda p: GeM(p=p)), ("mac", lambda p: MAC()), ("spoc", lambda p: SPoC())):
            for p in (3.0, 2.9137):
                if pooling != "gem" and p != 3.0:
                    continue
                meta = {"architecture": "identity", "local_whitening": False, "pooling": pooling, "regional": False,
                        "whitening": False, "mean": [0, 0, 0], "std": [1, 1, 1], "outputdim": C, "out_channels": C}
                net = ImageRetrievalNet([nn.Identity()], None, mod(p), None, meta).eval()
                for img in range(3):
                    outs = []
                    for s, (h, w) in enumerate(scales_hw):
                        x = torch.from_numpy(synth.fmap((1, C, h, w), 200 + 10 * img + s, "relu"))
                        o = net(x)                                            # (C,1)
                        head["tail_%s_p%g_i%d_s%d" % (pooling, p, img, s)] = o.numpy()
                        outs.append(o)
                    ms = CirMultiscaleAggregation(True, torch.device("cpu"))
                    v = ms.postprocess([o.clone() for o in outs], net, False)  # msp rule + aggregate_tensor
                    head["agg_%s_p%g_i%d" % (pooling, p, img)] = v.numpy()
                    v1 = CirMultiscaleAggregation([1], torch.device("cpu")).postprocess([outs[0].clone()], net, False)
                    head["agg1_%s_p%g_i%d" % (pooling, p, img)] = v1.numpy()
                    for dims in (None, 64, 32):
                        wh = CirtorchWhiten.__new__(CirtorchWhiten)
                        wh.device = torch.device("cpu")
                        wh.P = torch.tensor(lwd["P"], dtype=torch.float32)
                        wh.m = torch.tensor(lwd["m"], dtype=torch.float32)
                        wh.dimensions = dims or wh.P.shape[0]
                        head["wh_%s_p%g_i%d_d%s" % (pooling, p, img, dims)] = wh.postprocess(v.clone(), net, None).numpy()
    # whitenapply (cirtorch/utils/whiten.py:4-12), fp64 batch
    X = synth.descriptors(40, C, 9).T.astype(np.float64)
    head["whitenapply_X"] = X
    head["whitenapply_full"] = whitenapply(X, lwd["m"], lwd["P"])
    head["whitenapply_d48"] = whitenapply(X, lwd["m"], lwd["P"], 48)
    np.savez_compressed(os.path.join(OUT, "head.npz"), **head)

    # ---- 3. CLAHE: cv2 at the reference call site (transform/functional.py:114-117) -
    clahe = {}
    for key, hw, dist, clip, seed in synth.clahe_cases():
        img = synth.image_u8(hw, dist, seed)
        out = cv2.createCLAHE(clipLimit=int(clip), tileGridSize=(8, 8)).apply(img)
        clahe["in_sha_" + key] = np.array(synth.sha(img))
        if hw in synth.CLAHE_SMALL:
            clahe["out_" + key] = out
        clahe["out_sha_" + key] = np.array(synth.sha(out))
    # non-square grid and the float wrapper ChannelClahe.apply
    img = synth.image_u8((127, 93), "gamma", 77)
    clahe["grid4x6_127x93"] = cv2.createCLAHE(clipLimit=3, tileGridSize=(4, 6)).apply(img)
    chan = (synth.image_u8((200, 150), "gamma", 78).astype(np.float32) + np.float32(0.37)) / np.float32(255.3)
    clahe["channelclahe_200x150"] = ChannelClahe(4, 8).apply(chan)
    # the whole transform: RGB -> Lab -> CLAHE(L) -> RGB  (ImageClahe.apply, functional.py:120-129)
    rgb = (np.random.RandomState(79).rand(90, 122, 3) ** 2.2).astype(np.float32)
    clahe["imageclahe_in_sha"] = np.array(synth.sha(rgb))
    clahe["imageclahe_90x122"] = ImageClahe(4, 8, "lab").apply(rgb.copy())
    clahe["rgb2lab_90x122"] = cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)
    np.savez_compressed(os.path.join(OUT, "clahe.npz"), **clahe)

    # ---- 4. similarity / ranks / mAP (cirscore.py:69-71, evaluate.py) ---------------
    srch = {}
    db = synth.descriptors(500, 64, 11, clusters=20)
    q, src = synth.planted_queries(db, 12, 12)
    vecs, qvecs = np.ascontiguousarray(db.T), np.ascontiguousarray(q.T)
    sc = np.dot(vecs.T, qvecs)
    rk = np.argsort(-sc, axis=0)                       # the reference's (unstable) call
    srch["scores"] = sc
    srch["ranks_ref_unstable"] = rk
    srch["ranks_stable"] = np.argsort(-sc, axis=0, kind="stable")
    sct = np.round(sc * 20).astype(np.float32) / 20     # tie-heavy, includes +-0
    srch["scores_ties"] = sct
    srch["ranks_ties_stable"] = np.argsort(-sct, axis=0, kind="stable")
    gnd = synth.gnd_okjunk(500, 12, 13, empty_every=5)
    with contextlib.redirect_stdout(io.StringIO()):
        m, aps, pr, prs = compute_map(rk, gnd, [1, 5, 10])
        srch["okjunk_map"] = np.array(m)
        srch["okjunk_aps"] = aps
        srch["okjunk_pr"] = pr
        srch["okjunk_prs"] = prs
        gnd2 = synth.gnd_emh(500, 12, 14)
        avg, per = compute_map_and_print("roxford5k", rk, gnd2)
    for k_, v_ in avg.items():
        srch["emh_" + k_] = np.array(v_)
    for k_, v_ in per.items():
        srch["emh_" + k_] = v_
    np.savez_compressed(os.path.join(OUT, "search.npz"), **srch)


if __name__ == "__main__":
    main()
