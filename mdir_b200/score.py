"""Drop-in for mdir's evaluation criterion ``CirDatasetAp``
(mdir/components/optim/score/cirscore.py:16-80) and ``install()``, which patches the
reference's registries in place so that ``mdir.stages.validate.validate(scenario, ())`` runs
unchanged with the similarity + ranking (and pooling / whitening / CLAHE) on the B200 kernels.

Nothing here imports the reference at module import time: ``install()`` needs an importable
``mdir`` package (the user's checkout) and raises otherwise."""

from . import layers, wrappers, clahe, search, evaluate, extract


def rank_and_evaluate(dataset, vecs, qvecs, gnd, compute_map_and_print=None, device="cuda"):
    """cirscore.py:65-71 on the device: similarity (3xTF32, fp32-faithful) -> full ranks -> mAP.
    The (N_db, N_q) ranks never leave the GPU; pass the reference's compute_map_and_print to
    evaluate on the host instead (it receives the same int64 C-order array)."""
    import torch
    dev = torch.device(device if str(device) != "cpu" else "cuda")
    index = search.Index(vecs, dxn=True, device=dev, keep_fp32=True)
    q = search._as_dev_f32(qvecs, dev).t().contiguous()
    ranks = index.ranks(q, precision="fp32")
    if compute_map_and_print is not None:
        return compute_map_and_print(dataset, ranks.cpu().numpy(), gnd)
    return evaluate.compute_map_and_print(dataset, ranks, gnd, device=dev)


def make_cirdatasetap(base_cls, extract_vectors, compute_map_and_print, stopwatch_cls):
    """Build the replacement class against the reference's own base (dataset parsing, logging and
    mAP stay the reference's; only cirscore.py:69-70 changes)."""

    class CirDatasetAp(base_cls):
        def __call__(self, network, device, logger):
            stopwatch = stopwatch_cls()
            print('>> {}: database images...'.format(self.dataset))
            vecs = extract_vectors(network, self.images, self.image_size, self.transforms, device=device)
            print('>> {}: query images...'.format(self.dataset))
            if self.images == self.qimages and set(self.bbxs) == {None}:
                qvecs = vecs.clone()
            else:
                qvecs = extract_vectors(network, self.qimages, self.image_size, self.transforms, device=device, bbxs=self.bbxs)
            stopwatch.lap("extract_descriptors")
            print('>> {}: Evaluating...'.format(self.dataset))
            averages, scores = rank_and_evaluate(self.dataset, vecs, qvecs, self.gnd, device=device)
            stopwatch.lap("compute_score")
            first_score = scores[list(scores.keys())[0]]
            logger(None, len(first_score), "dataset", stopwatch.reset(), "scalar/time")
            logger(None, len(first_score), "score_avg", averages, "scalar/score")
            assert len({len(x) for x in scores.values()}) == 1
            for i, _ in enumerate(first_score):
                logger(i, len(first_score), "score", {x: scores[x][i] for x in scores}, "scalar/score")

    return CirDatasetAp


class _NoLoaderWorkers:
    """While active, torch.utils.data.DataLoader ignores num_workers (forces 0).  The reference's extract_vectors
    hard-codes six forked workers (imageretrievalnet.py:284-287); with the GPU-backed CLAHE transforms installed those
    workers would have to initialise CUDA after a fork, which CUDA refuses."""

    def __enter__(self):
        import torch.utils.data as tud
        self._tud, self._orig = tud, tud.DataLoader
        orig = self._orig

        class _Loader(orig):
            def __init__(self, *args, **kwargs):
                kwargs["num_workers"] = 0
                super().__init__(*args, **kwargs)

        tud.DataLoader = _Loader
        return self

    def __exit__(self, *exc):
        self._tud.DataLoader = self._orig
        return False


def _batched_or_reference(reference_extract, gpu_transforms):
    """extract_vectors that uses the batched device path and falls back to the reference's per-image loop for
    networks outside the hot path (branched / composite models, local / in-model whitening, regional pooling, other
    wrappers) -- ANY failure to recognise the network falls back, never a partial evaluation (extract._Plan)."""
    def extract_vectors(net, images, image_size, transform, bbxs=None, ms=[1], msp=1, print_freq=10, device=None):
        try:
            extract._Plan(net, ms, msp)                       # recognition only: cheap, raises NotImplementedError
        except NotImplementedError:
            if gpu_transforms:
                with _NoLoaderWorkers():
                    return reference_extract(net, images, image_size, transform, bbxs=bbxs, ms=ms, msp=msp, print_freq=print_freq, device=device)
            return reference_extract(net, images, image_size, transform, bbxs=bbxs, ms=ms, msp=msp, print_freq=print_freq, device=device)
        return extract.extract_vectors(net, images, image_size, transform, bbxs=bbxs, ms=ms, msp=msp, print_freq=print_freq,
                                       device=device)
    return extract_vectors


def _lab_or_reference(ours, ref_cls, colorspace_pos):
    """TRANSFORMS factory: colorspace 'lab' (the CLAHE scenario's) -> the device-backed class; 'luv' / 'lsh' / 'gray'
    -> the reference's own class, unchanged.  Arguments arrive as strings from the "name:arg:arg" mini-language
    (transform/__init__.py:35-44)."""
    def factory(*args, **kwargs):
        cs = kwargs.get("colorspace", args[colorspace_pos] if len(args) > colorspace_pos else "lab")
        if str(cs).lower() == "lab":
            return ours(*args, **kwargs)
        return ref_cls(*args, **kwargs)
    factory.__name__ = ours.__name__
    factory.device_class, factory.reference_class = ours, ref_cls
    return factory


def install(batched_extract=True, transforms=True):
    """Patch POOLING, WRAPPERS_LABELS, TRANSFORMS and SCORES of an importable ``mdir`` in place
    (SURVEY.md 8b).  Returns the dict of patched registry entries.
    transforms=False leaves the CLAHE entries of TRANSFORMS alone: use it for stages whose DataLoaders fork worker
    processes (training), where a GPU-backed transform cannot run."""
    try:
        import mdir  # noqa: F401
        import cirtorch.networks.imageretrievalnet as irn
        import mdir.components.data.wrapper as mwrap
        import mdir.components.data.transform as mtrans
        import mdir.components.optim.score as mscore
        import mdir.components.optim.score.cirscore as cirscore
        from mdir.tools.stats import StopWatch
    except ImportError as exc:
        raise RuntimeError("mdir_b200.install() needs the reference package `mdir` importable "
                           "(put the jenicek/mdir checkout on sys.path): %s" % exc)
    patched = {}
    for key, cls in layers.POOLING.items():
        irn.POOLING[key] = cls
        patched["POOLING[%s]" % key] = cls
    irn.L2N = layers.L2N
    patched["L2N"] = layers.L2N
    mwrap.WRAPPERS_LABELS["cirwhiten"] = wrappers.CirtorchWhiten
    mwrap.WRAPPERS_LABELS["cirmultiscale"] = wrappers.CirMultiscaleAggregation
    patched["WRAPPERS_LABELS[cirwhiten]"] = wrappers.CirtorchWhiten
    patched["WRAPPERS_LABELS[cirmultiscale]"] = wrappers.CirMultiscaleAggregation
    if transforms:
        # positional index of `colorspace` in each constructor (photometric_transforms.py:12,27)
        for key, cls, cs_pos in (("apply_clahe", clahe.ApplyClahe, 1), ("add_clahe_fromrgb", clahe.AddClaheFromRgb, 2),
                                 ("create_clahed", clahe.CreateClahedImage, 1)):
            if key in mtrans.TRANSFORMS:
                ref_cls = mtrans.TRANSFORMS[key]
                ref_cls = getattr(ref_cls, "reference_class", ref_cls)          # install() twice: keep the original
                mtrans.TRANSFORMS[key] = _lab_or_reference(cls, ref_cls, cs_pos)
                patched["TRANSFORMS[%s]" % key] = cls
    ev = _batched_or_reference(cirscore.extract_vectors, transforms) if batched_extract else cirscore.extract_vectors
    new_cls = make_cirdatasetap(cirscore.CirDatasetAp, ev, cirscore.compute_map_and_print, StopWatch)
    mscore.SCORES["cirdatasetap"] = new_cls
    patched["SCORES[cirdatasetap]"] = new_cls
    return patched
