"""Pin oracle/oracle.py against outputs of the reference itself (tests/golden/,
written by oracle/make_golden.py from the live /root/reference import) and, for
CLAHE, against the installed cv2 wheel (the reference's arbiter)."""
import numpy as np
import pytest

from oracle import oracle, synth

RTOL = 2e-6   # oracle is fp64-then-round; reference is fp32 ATen (3.9e-7 measured)


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("si", range(len(synth.POOL_SHAPES)))
@pytest.mark.parametrize("kind", ["relu", "signed", "zeros"])
def test_pooling(golden, si, kind):
    g = golden("pooling")
    x = synth.fmap(synth.POOL_SHAPES[si], 100 + si, kind)
    tag = "s%d_%s" % (si, kind)
    assert np.array_equal(oracle.mac(x), g[tag + "_mac"])
    close(oracle.spoc(x), g[tag + "_spoc"], atol=1e-7)
    close(oracle.l2n(oracle.mac(x)), g[tag + "_l2n_mac"], atol=1e-7)
    for p in synth.POOL_PS:
        close(oracle.gem(x, p), g[tag + "_gem_p%g" % p], rtol=5e-6)
        close(oracle.l2n(oracle.gem(x, p)), g[tag + "_l2n_gem_p%g" % p], rtol=5e-6, atol=1e-8)


def test_gem_zeros_value():
    # gem(zeros) = 9.99999656e-07 (clamp -> eps)
    v = oracle.gem(np.zeros((1, 4, 3, 3), np.float32), 3.0)
    close(v, np.full((1, 4, 1, 1), 9.99999656e-07), rtol=1e-6)


def test_head(golden):
    g = golden("head")
    C = 128
    hw = [(32, 24), (23, 17), (16, 12)]
    for pooling, ps in (("gem", (3.0, 2.9137)), ("mac", (3.0,)), ("spoc", (3.0,))):
        for p in ps:
            for img in range(3):
                fm = [synth.fmap((1, C, h, w), 200 + 10 * img + s, "relu") for s, (h, w) in enumerate(hw)]
                outs = []
                for s in range(3):
                    o = oracle.net_tail(fm[s], pooling, p)
                    close(o, g["tail_%s_p%g_i%d_s%d" % (pooling, p, img, s)], rtol=5e-6, atol=1e-8)
                    outs.append(o[:, 0])
                msp = oracle.multiscale_msp(3, pooling, False, False, p)
                v = oracle.aggregate_tensor(outs, 3, C, msp)
                close(v, g["agg_%s_p%g_i%d" % (pooling, p, img)], rtol=5e-6, atol=1e-8)
                close(oracle.aggregate_tensor(outs[:1], 1, C, 1.0), g["agg1_%s_p%g_i%d" % (pooling, p, img)], rtol=5e-6, atol=1e-8)
                for dims in (None, 64, 32):
                    w = oracle.cirwhiten_postprocess(v, g["lw_m"], g["lw_P"], dims)
                    close(w, g["wh_%s_p%g_i%d_d%s" % (pooling, p, img, dims)], rtol=2e-5, atol=2e-7)
                    w2 = oracle.gem_head(fm, p, 1e-6, g["lw_m"], g["lw_P"], dims, pooling)
                    assert np.array_equal(w, w2)
    close(oracle.whitenapply(g["whitenapply_X"], g["lw_m"], g["lw_P"]), g["whitenapply_full"], rtol=1e-12, atol=1e-15)
    close(oracle.whitenapply(g["whitenapply_X"], g["lw_m"], g["lw_P"], 48), g["whitenapply_d48"], rtol=1e-12, atol=1e-15)


def test_clahe_golden(golden):
    g = golden("clahe")
    for key, hw, dist, clip, seed in synth.clahe_cases():
        img = synth.image_u8(hw, dist, seed)
        assert synth.sha(img) == str(g["in_sha_" + key]), "synthetic input drifted: " + key
        if hw in synth.CLAHE_LARGE and (dist != "gamma" or clip != 4):
            continue          # keep the CPU suite short; the large cases run on the GPU side
        out = oracle.clahe_u8(img, clip, 8, 8)
        assert synth.sha(out) == str(g["out_sha_" + key]), key
        if hw in synth.CLAHE_SMALL:
            assert np.array_equal(out, g["out_" + key]), key
    img = synth.image_u8((127, 93), "gamma", 77)
    assert np.array_equal(oracle.clahe_u8(img, 3, 4, 6), g["grid4x6_127x93"])
    chan = (synth.image_u8((200, 150), "gamma", 78).astype(np.float32) + np.float32(0.37)) / np.float32(255.3)
    assert np.array_equal(oracle.channel_clahe(chan, 4, 8), g["channelclahe_200x150"])


def test_image_clahe_lab_path(golden):
    cv2 = pytest.importorskip("cv2")
    g = golden("clahe")
    rgb = (np.random.RandomState(79).rand(90, 122, 3) ** 2.2).astype(np.float32)
    assert synth.sha(rgb) == str(g["imageclahe_in_sha"])
    lat = oracle.cv2_lab_lattice()
    import os
    shipped = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mdir_b200", "data", "rgb2lab_lut_s16.npy"))
    assert np.array_equal(lat, shipped.astype(np.int64)), "shipped Lab lattice differs from the installed cv2"
    assert np.array_equal(oracle.rgb2lab_cv(rgb, lat), g["rgb2lab_90x122"])                  # bit-exact restatement
    assert np.array_equal(oracle.rgb2lab_cv(rgb, lat), cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB))
    lab = g["rgb2lab_90x122"]
    np.testing.assert_allclose(oracle.lab2rgb_cv(lab), cv2.cvtColor(lab, cv2.COLOR_LAB2RGB), rtol=0, atol=1e-5)
    np.testing.assert_allclose(oracle.image_clahe(rgb, 4, 8, lat), g["imageclahe_90x122"], rtol=0, atol=1e-5)


def test_clahe_vs_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    for i, (hw, dist, clip, grid) in enumerate([((31, 57), "gamma", 4, (8, 8)), ((100, 100), "bimodal", 2, (8, 8)),
                                                ((65, 129), "uniform", 40, (8, 8)), ((50, 70), "gradient", 4, (5, 3)),
                                                ((3, 5), "uniform", 4, (8, 8)), ((64, 48), "flat", 4, (8, 8))]):
        img = synth.image_u8(hw, dist, 500 + i)
        ref = cv2.createCLAHE(clipLimit=clip, tileGridSize=grid).apply(img)
        assert np.array_equal(oracle.clahe_u8(img, clip, grid[0], grid[1]), ref), (hw, dist, clip, grid)


def test_search_and_map(golden):
    g = golden("search")
    db = synth.descriptors(500, 64, 11, clusters=20)
    q, _ = synth.planted_queries(db, 12, 12)
    vecs, qvecs = np.ascontiguousarray(db.T), np.ascontiguousarray(q.T)
    sc = oracle.scores(vecs, qvecs)
    close(sc, g["scores"], rtol=0, atol=1e-6)
    assert np.array_equal(oracle.ranks_from_scores(g["scores"]), g["ranks_stable"])
    assert np.array_equal(oracle.ranks_from_scores(g["scores_ties"]), g["ranks_ties_stable"])
    # the reference's own unstable argsort is a valid ordering of the same scores
    ru = g["ranks_ref_unstable"]
    assert np.array_equal(np.take_along_axis(g["scores"], ru, 0), np.take_along_axis(g["scores"], g["ranks_stable"], 0))
    idx, val = oracle.topk_from_scores(g["scores_ties"], 17)
    assert np.array_equal(idx, g["ranks_ties_stable"][:17])
    gnd = synth.gnd_okjunk(500, 12, 13, empty_every=5)
    m, aps, pr, prs = oracle.compute_map(ru, gnd, [1, 5, 10])
    close(m, g["okjunk_map"], rtol=1e-12)
    np.testing.assert_array_equal(np.isnan(aps), np.isnan(g["okjunk_aps"]))
    close(np.nan_to_num(aps), np.nan_to_num(g["okjunk_aps"]), rtol=1e-12)
    close(pr, g["okjunk_pr"], rtol=1e-12)
    close(np.nan_to_num(prs), np.nan_to_num(g["okjunk_prs"]), rtol=1e-12)
    avg, per, _ = oracle.compute_map_emh(ru, synth.gnd_emh(500, 12, 14))
    for k in avg:
        close(avg[k], g["emh_" + k], rtol=1e-12)
    for k in per:
        close(np.nan_to_num(per[k]), np.nan_to_num(g["emh_" + k]), rtol=1e-12)


def test_qe_dba_properties():
    # parity-unpinned restatements: check the defining properties only
    db = synth.descriptors(300, 32, 21, clusters=10)
    q, _ = synth.planted_queries(db, 5, 22)
    q2 = oracle.alpha_qe(db, q, alpha=3.0, n_qe=10)
    close(np.linalg.norm(q2, axis=1), np.ones(5), rtol=1e-6)
    q0 = oracle.alpha_qe(db, q, alpha=0.0, n_qe=4)       # alpha=0 == plain average QE
    idx, _ = oracle.topk_from_scores(db @ q.T, 4)
    ref = q[0] + db[idx[:, 0]].sum(0)
    close(q0[0], ref / np.linalg.norm(ref), rtol=1e-5, atol=1e-7)
    d2 = oracle.dba(db, alpha=3.0, k_dba=5, chunk=128)
    close(np.linalg.norm(d2, axis=1), np.ones(300), rtol=1e-6)
    close(d2, oracle.dba(db, alpha=3.0, k_dba=5, chunk=300), rtol=1e-6, atol=1e-7)


# ---- f4: hard-negative mining and whitening learning ----------------------------------------------

def test_mining_oracle_matches_reference_create_epoch_tuples(golden):
    g = golden("mining")
    for case in range(3):
        t = "c%d_" % case
        nnum = int(g[t + "nnum"])
        pos, ndist = oracle.mine_negatives(g[t + "qvecs"], g[t + "poolvecs"], g[t + "qclusters"], g[t + "poolclusters"], nnum)
        assert np.array_equal(g[t + "idxs2images"][pos], g[t + "nidxs"])
        np.testing.assert_allclose(ndist.reshape(-1), g[t + "ndist"], rtol=0, atol=2e-6)
        # the rule itself: never the query's cluster, never two from one cluster
        pc = g[t + "poolclusters"][pos]
        assert not np.any(pc == g[t + "qclusters"][:, None])
        assert all(len(set(row)) == nnum for row in pc)


def test_whitenlearn_oracle_matches_reference(golden):
    g = golden("whitenlearn")
    for case in range(2):
        t = "c%d_" % case
        N, D, seed, clusters = [int(x) for x in g[t + "X_recipe"]]
        X = synth.descriptors(N, D, seed, clusters=clusters).T.astype(np.float64)
        m, P = oracle.whitenlearn(X, g[t + "qidxs"], g[t + "pidxs"])
        np.testing.assert_allclose(m, g[t + "m"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(oracle.whitening_rows_aligned(P, g[t + "P"]), g[t + "P"], rtol=0, atol=1e-7 * np.abs(g[t + "P"]).max())
        app = oracle.whitenapply(X[:, :50], m, P)
        np.testing.assert_allclose(app.T @ app, g[t + "applied"].T @ g[t + "applied"], rtol=0, atol=1e-9)   # sign-free
        mp, Pp = oracle.pcawhitenlearn(X)
        np.testing.assert_allclose(mp, g[t + "pca_m"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(oracle.whitening_rows_aligned(Pp, g[t + "pca_P"]), g[t + "pca_P"], rtol=0,
                                   atol=1e-7 * np.abs(g[t + "pca_P"]).max())


def test_extract_vectors_oracle_matches_reference(golden):
    """extract_ss / extract_ms of the live reference (cirtorch extract_vectors on CPU) == oracle tail + aggregation on the
    same feature maps (the backbone is stock torch on both sides)."""
    import torch
    g = golden("extract")
    nn = torch.nn
    feats = nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.ReLU(), nn.Conv2d(16, 48, 3, stride=2, padding=1), nn.ReLU())
    feats.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")})
    p = float(g["p"])
    ms = [float(s) for s in g["ms"]]
    with torch.no_grad():
        for i in range(6):
            x = torch.from_numpy(g["input_%d" % i]).unsqueeze(0)
            close(oracle.net_tail(feats(x).numpy(), "gem", p)[:, 0], g["vecs_ss"][:, i], rtol=5e-6, atol=1e-7)
            outs = []
            for s in ms:
                xs = x if s == 1 else torch.nn.functional.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)
                outs.append(oracle.net_tail(feats(xs).numpy(), "gem", p)[:, 0])
            close(oracle.aggregate_tensor(outs, 3, 48, p), g["vecs_ms"][:, i], rtol=1e-5, atol=1e-7)
            close(oracle.aggregate_tensor(outs, 3, 48, 1.0), g["vecs_ms_msp1"][:, i], rtol=1e-5, atol=1e-7)
            # mdir's Compose([CirtorchWhiten(32), CirMultiscaleAggregation(True)]): every scale interpolated, msp rule, Lw
            fm = [feats(torch.nn.functional.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)).numpy() for s in ms]
            close(oracle.gem_head(fm, p, 1e-6, g["lw_m"], g["lw_P"], 32), g["compose_wh32"][i], rtol=2e-5, atol=2e-7)


def test_transform_classes_oracle_matches_reference_registry(golden):
    """ApplyClahe / CreateClahedImage / AddClaheFromRgb built by the reference's TRANSFORMS registry (golden) == the
    oracle's Lab -> CLAHE(L) -> RGB restatement."""
    pytest.importorskip("cv2")
    g = golden("transforms")
    pic = g["pic"]
    lat = oracle.cv2_lab_lattice()
    for key, clip, grid in (("apply_clahe_4_lab_8", 4, 8), ("apply_clahe_2_lab_4", 2, 4), ("create_clahed_1", 4, 8)):
        np.testing.assert_allclose(oracle.image_clahe(pic, clip, grid, lat), g[key], rtol=0, atol=2e-5)
    for key, clip, grid in (("add_clahe_fromrgb", 4, 8), ("add_clahe_fromrgb_2_4", 2, 4)):
        L = oracle.rgb2lab_cv(pic, lat)[:, :, 0] / np.float32(100.0)
        assert np.array_equal(g[key][:, :, :3], pic)
        assert np.array_equal(oracle.channel_clahe(L.astype(np.float32), clip, grid), g[key][:, :, 3])
