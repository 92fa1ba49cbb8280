"""Batched descriptor extraction: drop-in for cirtorch's ``extract_vectors``
(mdir/external/cirtorch/networks/imageretrievalnet.py:277-324).

The reference runs ``vecs[:, i] = net(input).cpu()`` per image: ~20 tiny ATen launches for the head
and one blocking device->host copy per image.  Here the backbone (stock torch, untouched) runs per
image and scale, the feature maps stay on the device, and every ``group`` images the whole head --
pool, L2N, multi-scale aggregation, Lw whitening -- is ONE pass of the ``RetrievalHead`` kernels;
nothing is copied to the host until the end (or never, with ``return_device=True``: the (N, D) matrix
then feeds ``Index`` directly).

Three kinds of ``net`` are understood (recognised by class name, nothing of the reference is imported):
  * cirtorch's ``ImageRetrievalNet``: ``.features``, ``.pool``, ``.meta``  (+ ``ms`` / ``msp`` arguments,
    the semantics of ``extract_ss`` / ``extract_ms``);
  * mdir's ``CirNetwork`` / ``SingleNetwork``: ``.model`` (an ImageRetrievalNet) and ``.wrappers['eval']`` holding
    ``CirMultiscaleAggregation`` and/or ``CirtorchWhiten`` (mdir/learning/network.py:88-89);
  * mdir's ``SequentialNetwork`` of such networks (normaliser -> CirNet, network.py:204-236): the earlier
    networks run unchanged on every scaled image, the last one's head is batched.
Everything else -- subclasses that override ``forward`` (``ImageRetrievalNetBranched``), local whitening,
in-model whitening layers, regional pooling, unknown wrappers -- raises NotImplementedError, which
``install()``'s extractor turns into the reference's own per-image loop.
"""
import sys

import torch
import torch.nn.functional as F

from . import _lib
from .wrappers import RetrievalHead


class _Plan:
    """What the batched path will run for `net`.  The path is WHITELISTED by type: anything it does not positively
    recognise raises NotImplementedError, and install()'s extractor then falls back to the reference's own per-image
    loop -- a network with an overridden forward (ImageRetrievalNetBranched, cirnet.py:25-45) or an unknown
    composition must never be evaluated through `.model.features` alone."""

    def __init__(self, net, ms, msp):
        self.scales = list(ms)
        self.msp = float(msp)
        self.lw = None
        self.dimensions = None
        self.interp_unit_scale = False
        self.pre = []                                 # networks applied to every scaled image before the last model
        kind = type(net).__name__
        wr = None
        if hasattr(net, "networks") or hasattr(net, "sequence"):
            # mdir's SequentialNetwork (learning/network.py:204-236; e.g. U-Net normaliser -> CirNet,
            # examples/iccv19/eval_composition.yml): the eval wrappers of the LAST network wrap the whole chain, so
            # every scaled image goes through the earlier networks (their own wrappers included) and then through the
            # last model's backbone.
            if kind != "SequentialNetwork" or not hasattr(net, "networks") or not hasattr(net, "sequence"):
                raise NotImplementedError("composite network %s is outside the batched extraction path" % kind)
            seq = list(net.sequence)
            self.pre = [net.networks[k] for k in seq[:-1]]
            last = net.networks[seq[-1]]
            if type(last).__name__ not in ("CirNetwork", "SingleNetwork") or last.model is not net.model:
                raise NotImplementedError("last stage %s of the sequence is outside the batched extraction path" % type(last).__name__)
            for stage in self.pre:
                if type(stage).__name__ not in ("CirNetwork", "SingleNetwork"):
                    raise NotImplementedError("stage %s of the sequence is outside the batched extraction path" % type(stage).__name__)
            model, wr = net.model, net.wrappers
        elif hasattr(net, "model"):
            if kind not in ("CirNetwork", "SingleNetwork"):
                raise NotImplementedError("network wrapper %s is outside the batched extraction path" % kind)
            model, wr = net.model, getattr(net, "wrappers", None)
        else:
            model = net
        # the backbone must be cirtorch's ImageRetrievalNet with ITS forward (imageretrievalnet.py:93-115), not a subclass
        if type(model).__name__ != "ImageRetrievalNet" or not getattr(type(model).forward, "__qualname__", "").endswith("ImageRetrievalNet.forward"):
            raise NotImplementedError("model %s is outside the batched extraction path" % type(model).__name__)
        self.model = model
        if wr is not None:
            comp = wr.get("eval") if isinstance(wr, dict) else wr
            self.scales, self.msp = [1], None
            for w in getattr(comp, "wrappers", []):
                name = w.__class__.__name__
                if name == "CirMultiscaleAggregation":
                    self.scales = list(w.scales)
                    self.interp_unit_scale = len(self.scales) > 1          # wrapper.py:96-107 interpolates every scale
                elif name == "CirtorchWhiten":
                    self.lw = {"P": w.P.detach().cpu().numpy(), "m": w.m.detach().cpu().numpy()}
                    self.dimensions = w.dimensions
                else:
                    raise NotImplementedError("wrapper %s is outside the batched extraction path" % name)
        meta = model.meta
        if getattr(model, "lwhiten", None) is not None or getattr(model, "whiten", None) is not None or meta.get("regional"):
            raise NotImplementedError("local / in-model whitening and regional pooling are outside the hot path")
        self.pooling = meta["pooling"]
        if self.pooling not in ("gem", "mac", "spoc"):
            raise NotImplementedError("pooling %r is outside the hot path" % self.pooling)
        self.p = float(model.pool.p.item()) if self.pooling == "gem" else 3.0
        self.eps = float(getattr(model.pool, "eps", 1e-6))
        self.out_dim = self.dimensions or meta.get("out_channels", meta.get("outputdim"))


def _head(plan, device):
    head = RetrievalHead(plan.pooling, p=plan.p, eps=plan.eps, whitening=plan.lw, dimensions=plan.dimensions,
                         nscales=len(plan.scales), regional=False, model_whitening=False, device=device)
    if plan.msp is not None:                     # explicit msp of extract_ms (the wrapper path uses the msp rule)
        head.msp = plan.msp if len(plan.scales) > 1 else 1.0
    return head


def extract_from_tensors(net, tensors, ms=(1,), msp=1, device=None, group=32, return_device=False):
    """tensors: iterable of already transformed images, (1,3,H,W) or (3,H,W) float tensors (any sizes).
    Returns (D, N) float32 on the host like the reference, or the (N, D) device matrix."""
    plan = _Plan(net, ms, msp)
    dev = torch.device(device) if device is not None else next(plan.model.parameters()).device
    if dev.type != "cuda":
        raise _lib.MdirError("extract_vectors needs a CUDA device (no CPU path)")
    head = _head(plan, dev)
    chunks, pending = [], []

    def flush():
        if pending:
            chunks.append(head(pending))
            pending.clear()

    with torch.no_grad():
        for n_img, x in enumerate(tensors):
            if isinstance(x, dict):                  # missing image -> NaN row (genericdataset.py:54-59)
                flush()
                chunks.append(torch.full((1, plan.out_dim), float("nan"), device=dev))
                continue
            x = x.to(dev, non_blocking=True)
            if x.dim() == 3:
                x = x.unsqueeze(0)
            for s in plan.scales:
                xs = x if (s == 1 and not plan.interp_unit_scale) else F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)
                for stage in plan.pre:               # SequentialNetwork: earlier networks run as they are (stock torch)
                    xs = stage(xs)
                pending.append(plan.model.features(xs).float().contiguous())
            if (n_img + 1) % group == 0:
                flush()
        flush()
    out = torch.cat(chunks) if chunks else torch.empty((0, plan.out_dim), device=dev)
    if return_device:
        return out
    return out.t().contiguous().cpu()


def extract_vectors(net, images, image_size, transform, bbxs=None, ms=[1], msp=1, print_freq=10, device=None,
                    num_workers=0, group=32, return_device=False):
    """Same arguments as cirtorch's extract_vectors.  ``images`` is a list of paths (loaded with the
    reference's own ``ImagesFromList``, which must be importable) or a list of tensors.
    num_workers defaults to 0 because mdir_b200's CLAHE transforms use the GPU."""
    if len(images) and isinstance(images[0], torch.Tensor):
        return extract_from_tensors(net, images, ms, msp, device, group, return_device)
    try:
        from cirtorch.datasets.genericdataset import ImagesFromList
    except ImportError as exc:
        raise RuntimeError("loading images from paths needs the reference's cirtorch package importable: %s" % exc)
    model = getattr(net, "model", net)
    if hasattr(net, "eval"):
        net.eval()
    loader = torch.utils.data.DataLoader(ImagesFromList(root='', images=images, imsize=image_size, bbxs=bbxs, transform=transform),
                                         batch_size=1, shuffle=False, num_workers=num_workers, pin_memory=True)

    def stream():
        for i, inp in enumerate(loader):
            if (i + 1) % print_freq == 0 or (i + 1) == len(images):
                sys.stdout.write('\r>>>> {}/{} done...'.format(i + 1, len(images)))
            yield inp
        sys.stdout.write('\n')

    dev = device if device is not None else next(model.parameters()).device
    return extract_from_tensors(net, stream(), ms, msp, dev, group, return_device)
