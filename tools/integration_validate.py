#!/usr/bin/env python
"""End-to-end integration check: the reference's own
``mdir.stages.validate.validate(scenario, ())`` run UNPATCHED and after ``mdir_b200.install()``, offline, on a
generated image folder + a random-weight resnet18 ``cirnet`` checkpoint written in mdir's checkpoint format.

    python tools/integration_validate.py fixture <root>          # images, TSVs, checkpoints, Lw  (deterministic)
    python tools/integration_validate.py run <root> <arm> <out.npz>

arms (one process each, so registries / CUDA state never leak between them):
    ref_cpu       the reference as shipped, CUDA hidden (CUDA_VISIBLE_DEVICES="")            -- the oracle
    ref_cuda      the reference as shipped on cuda:0 (same backbone arithmetic as ours)
    ours          mdir_b200.install(): POOLING / WRAPPERS_LABELS / TRANSFORMS / SCORES patched; batched extraction,
                  GPU CLAHE transform, tcgen05 scores, radix-sort ranks, mAP on the device
    ours_modules  mdir_b200.install(batched_extract=False, transforms=False): the reference's per-image loop drives OUR
                  nn.Modules and wrappers (GeM / L2N built by init_network from the yaml ``pooling`` key,
                  cirnet.py:10-22; CirtorchWhiten / CirMultiscaleAggregation through Compose)

scenarios (BASELINE.json configs[0] / configs[1] shapes in miniature, plus the paper's composition):
    c1     cirnet resnet18 GeM, single scale, no whitening; transforms "pil2np | totensor | normalize"
    c2     + CLAHE transform ("pil2np | apply_clahe | totensor | normalize"), wrappers 0_cirwhiten + 1_cirmultiscale
    c2seq  SequentialNetwork (1x1-conv normaliser -> cirnet, learning/network.py:204-236) under the same wrappers

Each run stores per scenario: the validate() result dict, per-query AP, and the (D, N) descriptor matrices that
``extract_vectors`` returned.  TEST INFRASTRUCTURE: imports the reference through oracle/ref_import.py."""
import json
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_SCENES, PER_SCENE, N_Q = 8, 4, 8
IMAGE_SIZE = 160


# ------------------------------------------------------------------------------------------ fixture
def make_fixture(root):
    import numpy as np
    import torch
    from PIL import Image
    from oracle import ref_import
    ref_import.import_reference()
    from mdir.components.model.network import initialize_model

    os.makedirs(os.path.join(root, "img"), exist_ok=True)
    rs = np.random.RandomState(7)

    def scene(seed):
        r = np.random.RandomState(seed)
        h, w = 200, 260
        img = np.zeros((h, w, 3), np.float32)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        for _ in range(14):                                    # blobs + gratings: texture a random conv net can tell apart
            cx, cy, s = r.uniform(0, w), r.uniform(0, h), r.uniform(10, 60)
            col = r.uniform(0, 1, 3).astype(np.float32)
            img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))[:, :, None] * col
            f, th = r.uniform(0.02, 0.3), r.uniform(0, np.pi)
            img += 0.15 * np.sin(f * (xx * np.cos(th) + yy * np.sin(th)))[:, :, None] * r.uniform(0, 1, 3).astype(np.float32)
        img -= img.min()
        return img / img.max()

    names, scene_of = [], []
    for s in range(N_SCENES):
        base = scene(100 + s)
        for v in range(PER_SCENE + 1):                         # PER_SCENE database views + 1 query view
            h0, w0 = rs.randint(0, 40), rs.randint(0, 50)
            hh, ww = rs.randint(140, 160), rs.randint(180, 210)
            crop = base[h0:h0 + hh, w0:w0 + ww]
            gain = rs.uniform(0.25, 1.0) if v % 2 else 1.0     # "night" views: darker + gamma
            crop = np.clip(crop * gain, 0, 1) ** (1.6 if v % 2 else 1.0)
            crop = np.clip(crop + rs.normal(0, 0.01, crop.shape), 0, 1)
            name = "s%02d_v%d.jpg" % (s, v)
            Image.fromarray((crop * 255).astype(np.uint8)).save(os.path.join(root, "img", name), quality=95)
            names.append(name)
            scene_of.append(s)
    db = [n for n in names if not n.endswith("_v%d.jpg" % PER_SCENE)]
    qs = [n for n in names if n.endswith("_v%d.jpg" % PER_SCENE)][:N_Q]
    with open(os.path.join(root, "db.tsv"), "w") as fh:
        fh.write("identifier\n" + "\n".join(db) + "\n")
    with open(os.path.join(root, "q.tsv"), "w") as fh:
        fh.write("query\tbbx\tok\tjunk\n")
        for i, qn in enumerate(qs):
            sc = qn[:3]
            ok = [d for d in db if d.startswith(sc)]
            junk = [db[(7 * i + 3) % len(db)]] if not db[(7 * i + 3) % len(db)].startswith(sc) else []
            bbx = "" if i % 3 else json.dumps([10, 8, 150, 120])                  # some queries carry a crop box
            fh.write("%s\t%s\t%s\t%s\n" % (qn, bbx, json.dumps(ok), json.dumps(junk)))

    torch.manual_seed(11)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    cir_params = {"architecture": "cirnet", "cir_architecture": "resnet18", "local_whitening": False, "pooling": "gem",
                  "regional": False, "whitening": False, "pretrained": False}
    cir = initialize_model(dict(cir_params))
    with torch.no_grad():
        cir.pool.p.fill_(2.9137)                                   # a trained, non-integer p
    for kind, transforms in (("c1", "pil2np | totensor | normalize"), ("c2", "pil2np | apply_clahe | totensor | normalize")):
        ckpt = {"type": "CirNetwork", "frozen": False,
                "network_params": {"model": dict(cir_params), "runtime": {"wrappers": "", "data": {"mean_std": [mean, std], "transforms": transforms}}},
                "model_state": cir.state_dict()}
        torch.save(ckpt, os.path.join(root, "net_%s.pth" % kind))
    # sequence: per-pixel normaliser (1x1 convs, tanh output in [-1, 1]) -> cirnet
    norm_params = {"architecture": "pixelconv_regr", "in_channels": 3, "out_channels": 3, "hidden": [8]}
    norm = initialize_model({**norm_params, "hidden": [8]})
    seq = {
        "net": {"type": "SequentialNetwork", "frozen": False, "sequence": ["norm", "cir"], "network_hierarchy": {"norm": [], "cir": []}},
        "norm": {"type": "SingleNetwork", "frozen": False,
                 "network_params": {"model": norm_params, "runtime": {"wrappers": "", "data": {"mean_std": [mean, std], "transforms": "pil2np | apply_clahe | totensor | normalize"}}},
                 "model_state": norm.state_dict()},
        "cir": {"type": "CirNetwork", "frozen": False,
                "network_params": {"model": dict(cir_params), "runtime": {"wrappers": "", "data": {"mean_std": [mean, std], "transforms": "pil2np | totensor | normalize"}}},
                "model_state": cir.state_dict()},
    }
    torch.save({**seq["net"], "_networks_included": {"norm": seq["norm"], "cir": seq["cir"]}}, os.path.join(root, "net_c2seq.pth"))
    D = 512
    g = np.random.RandomState(5)
    q_, _ = np.linalg.qr(g.randn(D, D))
    P = (q_ * g.uniform(0.5, 2.0, D)[:, None]).astype(np.float64)
    m = (g.randn(D, 1) * 0.02).astype(np.float64)
    with open(os.path.join(root, "lw.pkl"), "wb") as fh:
        pickle.dump({"m": m, "P": P}, fh)
    return root


def scenario(root, kind):
    wrappers = {"train": None, "eval": None if kind == "c1" else
                {"0_cirwhiten": {"whitening": os.path.join(root, "lw.pkl"), "dimensions": None}, "1_cirmultiscale": {"scales": True}}}
    return {
        "network": {"path": os.path.join(root, "net_%s.pth" % kind), "runtime": {"wrappers": wrappers}},
        "validation": {"type": "MultiCriterialValidation", "decisive_criterion": None,
                       "synth": {"type": "SingleValidation", "frequency": None, "network_overlay": None, "data": None,
                                 "criterion": {"type": "cirdatasetap", "image_size": IMAGE_SIZE,
                                               "dataset": {"name": "synth", "queries": os.path.join(root, "q.tsv"), "db": os.path.join(root, "db.tsv"),
                                                           "imgdir": os.path.join(root, "img") + "/"}}}},
        "data": {}}


# ------------------------------------------------------------------------------------------ one arm
def run_arm(root, arm, out_path, kinds=("c1", "c2", "c2seq")):
    if arm == "ref_cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import copy
    import numpy as np
    import torch
    torch.backends.cudnn.allow_tf32 = False                        # compare fp32 backbones, not TF32 convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    from oracle import ref_import
    ref_import.import_reference()
    import mdir.stages.validate as mvalidate
    import mdir.components.optim.score.cirscore as cirscore

    record = []
    info = {"arm": arm, "cuda": bool(torch.cuda.is_available())}
    if arm in ("ours", "ours_modules"):
        import mdir_b200
        from mdir_b200 import extract
        patched = mdir_b200.install() if arm == "ours" else mdir_b200.install(batched_extract=False, transforms=False)
        info["patched"] = sorted(patched.keys())
        if arm == "ours":
            inner = extract.extract_vectors

            def rec_ours(*a, **k):
                v = inner(*a, **k)
                record.append(v.clone())
                return v
            extract.extract_vectors = rec_ours
            fb = cirscore.extract_vectors

            def no_fallback(*a, **k):
                raise AssertionError("the batched path fell back to the reference extractor")
            cirscore.extract_vectors = no_fallback
            del fb
    if arm != "ours":
        inner_ref = cirscore.extract_vectors

        def rec_ref(*a, **k):
            v = inner_ref(*a, **k)
            record.append(v.clone())
            return v
        cirscore.extract_vectors = rec_ref
        if arm == "ours_modules":                                  # install() captured the original: re-install around the recorder
            import mdir_b200
            mdir_b200.install(batched_extract=False, transforms=False)

    nets = {}
    orig_load = mvalidate.load_network

    def spy_load(params, device):
        net = orig_load(params, device)
        nets["last"] = net
        return net
    mvalidate.load_network = spy_load

    out = {}
    for kind in kinds:
        del record[:]
        res = mvalidate.validate(copy.deepcopy(scenario(root, kind)), ())[0]["eval"]
        net = nets["last"]
        info["%s_pool_module" % kind] = type(net.model.pool).__module__ + "." + type(net.model.pool).__name__
        info["%s_device" % kind] = str(next(net.model.parameters()).device)
        wr = net.wrappers["eval"].wrappers if hasattr(net.wrappers["eval"], "wrappers") else []
        info["%s_wrappers" % kind] = [type(w).__module__ + "." + type(w).__name__ for w in wr]
        assert len(record) == 2, len(record)
        out[kind + "_vecs"] = record[0].cpu().numpy()
        out[kind + "_qvecs"] = record[1].cpu().numpy()
        aps = [float(res[k]) for k in sorted((k for k in res if k.startswith("synth/validation/score:")), key=str)]
        out[kind + "_result_keys"] = np.array(sorted(res.keys()))
        out[kind + "_result_vals"] = np.array([float(res[k]) if np.isscalar(res[k]) or isinstance(res[k], float) else np.nan for k in sorted(res.keys())])
        out[kind + "_map"] = np.array([res[k] for k in res if k.endswith("score_avg:map")][:1], dtype=np.float64)
        info["%s_n_scores" % kind] = len(aps)
    out["info"] = np.array(json.dumps(info))
    np.savez(out_path, **out)
    print(json.dumps(info))


if __name__ == "__main__":
    if sys.argv[1] == "fixture":
        print(make_fixture(sys.argv[2]))
    elif sys.argv[1] == "run":
        run_arm(sys.argv[2], sys.argv[3], sys.argv[4], tuple(sys.argv[5].split(",")) if len(sys.argv) > 5 else ("c1", "c2", "c2seq"))
    else:
        raise SystemExit(__doc__)
