"""mAP on the device: drop-in for ``compute_map`` / ``compute_map_and_print``
(mdir/external/cirtorch/utils/evaluate.py:39-152).  Same arguments and return values; the per-query
work (positions of positives after removing junk, trapezoidal AP, precision@k) runs in
csrc/evaluate.cu on the ranks array where it already lives."""
import numpy as np
import torch

from . import _lib

_NAN = float("nan")


def _flatten_gnd(gnd):
    """list of {'ok', ['junk']} -> (items int64, query-of-item int32, class-of-item int32 (0 ok / 1 junk), n_pos int32);
    one pass over the dicts, the per-item columns are built with np.repeat instead of per-query arrays."""
    oks = [np.asarray(g["ok"], dtype=np.int64).reshape(-1) for g in gnd]
    junks = [np.asarray(g["junk"], dtype=np.int64).reshape(-1) if "junk" in g else np.empty(0, np.int64) for g in gnd]
    n_ok = np.fromiter((a.shape[0] for a in oks), dtype=np.int64, count=len(gnd))
    n_junk = np.fromiter((a.shape[0] for a in junks), dtype=np.int64, count=len(gnd))
    qs = np.arange(len(gnd), dtype=np.int32)
    items = np.concatenate(oks + junks).astype(np.int64) if gnd else np.empty(0, np.int64)
    iq = np.concatenate([np.repeat(qs, n_ok), np.repeat(qs, n_junk)]).astype(np.int32)
    ic = np.concatenate([np.zeros(int(n_ok.sum()), np.int32), np.ones(int(n_junk.sum()), np.int32)])
    return items, iq, ic, n_ok.astype(np.int32)


def compute_map(ranks, gnd, kappas=(), device="cuda"):
    """ranks: (N_db, N_q) integer ranks (torch cuda tensor or numpy); gnd: list of {'ok', 'junk'}.
    Returns (map, aps, pr, prs) exactly like evaluate.py:39-111."""
    lib = _lib.lib()
    dev = torch.device(device)
    if isinstance(ranks, np.ndarray):
        ranks = torch.from_numpy(np.ascontiguousarray(ranks))
    r = ranks.to(dev, dtype=torch.int64).contiguous()
    n_db, n_q = r.shape
    assert n_q == len(gnd)
    kappas = [int(k) for k in kappas]
    items, iq, ic, npos = _flatten_gnd(gnd)
    with torch.cuda.device(dev):
        items_d, iq_d, ic_d = (torch.from_numpy(a).to(dev) for a in (items, iq, ic))
        npos_d = torch.from_numpy(npos).to(dev)
        kap_d = torch.tensor(kappas or [0], dtype=torch.int32, device=dev)
        aps_d = torch.empty((n_q,), dtype=torch.float64, device=dev)
        prs_d = torch.empty((n_q, max(len(kappas), 1)), dtype=torch.float64, device=dev)
        ws = torch.empty(lib.mdir_map_workspace_bytes(n_db, n_q), dtype=torch.uint8, device=dev)
        _lib.check(lib.mdir_compute_ap(_lib.ptr(r), n_db, n_q, _lib.ptr(items_d), _lib.ptr(iq_d), _lib.ptr(ic_d), items.shape[0],
                                       _lib.ptr(npos_d), _lib.ptr(kap_d), len(kappas), _lib.ptr(aps_d), _lib.ptr(prs_d), _lib.ptr(ws),
                                       _lib.stream()), "mdir_compute_ap")
    aps = aps_d.cpu().numpy()
    prs = prs_d.cpu().numpy()[:, :len(kappas)]
    valid = npos > 0
    nvalid = int(valid.sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        mp = float(aps[valid].sum() / nvalid) if nvalid else _NAN
        pr = prs[valid].sum(axis=0) / nvalid if nvalid else np.full(len(kappas), _NAN)
    return mp, aps, pr, prs


def compute_map_and_print(dataset, ranks, gnd, kappas=(1, 5, 10), device="cuda"):
    """evaluate.py:114-152: the old ok/junk protocol, or Easy / Medium / Hard for roxford5k / rparis6k."""
    if "ok" in gnd[0]:
        mp, aps, _, _ = compute_map(ranks, gnd, device=device)
        print('>> {}: mAP {:.2f}'.format(dataset, np.around(mp * 100, decimals=2)))
        return {"map": mp}, {"ap": aps}
    if dataset.startswith('roxford5k') or dataset.startswith('rparis6k'):
        def regroup(ok_keys, junk_keys):
            return [{"ok": np.concatenate([np.asarray(g[k], dtype=np.int64) for k in ok_keys]),
                     "junk": np.concatenate([np.asarray(g[k], dtype=np.int64) for k in junk_keys])} for g in gnd]
        res = {}
        for name, ok_keys, junk_keys in (("easy", ["easy"], ["junk", "hard"]), ("medium", ["easy", "hard"], ["junk"]),
                                         ("hard", ["hard"], ["junk", "easy"])):
            res[name] = compute_map(ranks, regroup(ok_keys, junk_keys), kappas, device=device)
        print('>> {}: mAP E: {}, M: {}, H: {}'.format(dataset, *[np.around(res[n][0] * 100, decimals=2) for n in ("easy", "medium", "hard")]))
        print('>> {}: mP@k{} E: {}, M: {}, H: {}'.format(dataset, list(kappas), *[np.around(res[n][2] * 100, decimals=2) for n in ("easy", "medium", "hard")]))
        return ({"map_easy": res["easy"][0], "map_medium": res["medium"][0], "map_hard": res["hard"][0]},
                {"ap_easy": res["easy"][1], "ap_medium": res["medium"][1], "ap_hard": res["hard"][1]})
    raise ValueError("unknown evaluation protocol for dataset %r" % dataset)
