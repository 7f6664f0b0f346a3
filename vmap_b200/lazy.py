"""Lazy handles that let the reference's call sequence

    emb = vmap(pe_model)(pe_param, pe_buffer, pcs)             # train.py:293
    alpha, color = vmap(fc_model)(fc_param, fc_buffer, emb)    # train.py:294
    loss, _ = loss.step_batch_loss(alpha, color, ...)          # train.py:303
    loss.backward(); optimiser.step()                          # train.py:324-325

run as ONE fused forward+loss+backward launch plus one fused AdamW launch: the first two
calls only record what to compute; ``step_batch_loss`` launches K0+K1.  Nothing here
computes on the CPU or in eager PyTorch -- if the CUDA library is missing the calls raise.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from .layout import FC_KEYS, PE_KEY

_DIRTY = weakref.WeakSet()        # ensembles holding gradients that optimiser.step() must consume


class LazyEmbedding:
    """Result of the positional embedding: just the points and who embeds them."""

    def __init__(self, pcs: torch.Tensor, pe=None, ens=None):
        self.pcs, self.pe, self.ens = pcs, pe, ens


class LazyHead:
    """alpha or colour of a LazyHeads; supports the few tensor idioms train.py applies."""

    def __init__(self, heads: "LazyHeads", kind: str):
        self.heads, self.kind = heads, kind

    def __getitem__(self, idx):          # bg_alpha[None, ...] (train.py:312)
        if idx is None or idx == (None, Ellipsis) or idx == (None,):
            return self
        raise IndexError("lazy network outputs only support [None, ...]; call .materialize() for values")

    def detach(self):
        return self

    def materialize(self) -> torch.Tensor:
        a, c = self.heads.materialize()
        return a if self.kind == "alpha" else c


class LazyHeads:
    def __init__(self, emb: LazyEmbedding, fc=None):
        self.emb, self.fc = emb, fc
        self.alpha, self.color = LazyHead(self, "alpha"), LazyHead(self, "color")

    def ensemble(self):
        if self.emb.ens is not None:
            return self.emb.ens
        return ensemble_for_modules(self.fc, self.emb.pe)

    def materialize(self):
        """Forward only (K4 / vmb_forward): alpha [..., 1], colour [..., 3]."""
        ens = self.ensemble()
        pcs = self.emb.pcs
        lead = pcs.shape[:-1]
        pts = pcs.reshape(ens.n_obj, -1, 3).contiguous().float()
        a, c = ens.eval_points(pts)
        return a.reshape(*lead, 1), c.reshape(*lead, 3)


def bind_modules(ens, row: int, fc, pe):
    """Copy one object's module parameters into row ``row`` of the packed block and make the
    module parameters VIEWS of that row: checkpoints (vmap.py:461-476) then serialise straight
    from the block and train.py:331-338's copy-back degenerates to a self-copy."""
    sd = dict(fc.named_parameters())
    with torch.no_grad():
        for k in FC_KEYS:
            dst = ens.view(k)[row]
            dst.copy_(sd[k].detach().to(dst.device, torch.float32))
            sd[k].data = dst
        dst = ens.view(PE_KEY)[row]
        dst.copy_(pe.B_layer.weight.detach().to(dst.device, torch.float32))
        pe.B_layer.weight.data = dst
        ens.scale[row] = float(pe.scale)
        if pe.scale.device != dst.device:
            pe.scale = pe.scale.to(dst.device)
            pe.frequency_bands = pe.frequency_bands.to(dst.device)
    fc._vmb_binding = (weakref.ref(ens), row)
    pe._vmb_binding = (weakref.ref(ens), row)


def ensemble_for_modules(fc, pe, device: Optional[torch.device] = None):
    """The ensemble a (fc, pe) module pair is bound to; an unbound pair (e.g. the separate
    background model, train.py:147-152,308-316) gets a private 1-object ensemble."""
    from .ensemble import VmapEnsemble
    b = getattr(fc, "_vmb_binding", None)
    if b is not None and b[0]() is not None:
        return b[0]()
    dev = device or next(fc.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("vmap_b200 modules must live on a CUDA device (there is no CPU path)")
    ens = VmapEnsemble(1, hidden=fc.hidden_size, n_unidir_funcs=pe.max_deg, scale=float(pe.scale), device=dev)
    bind_modules(ens, 0, fc, pe)
    ens.refresh_image()
    fc._vmb_private = ens          # keep it alive with the module
    return ens


class _FusedLoss(torch.autograd.Function):
    """Scalar loss whose gradient already sits in ``ens.grads`` (K1 computed it)."""

    @staticmethod
    def forward(ctx, anchor, ens, total):
        ctx.ens = ens
        return total.detach().view(())

    @staticmethod
    def backward(ctx, g):
        # d(total)/d(this loss) (1 for loss.backward()): handed to the AdamW launch, which multiplies the gradients
        # by it as it reads them -- no pass over ens.grads here
        ens = ctx.ens
        ens._grad_scale = g if ens._grad_scale is None else ens._grad_scale * g
        return None, None, None


def fused_loss(ens, total: torch.Tensor) -> torch.Tensor:
    """``total``: the one-element device tensor the step launch wrote the scalar loss into."""
    if getattr(ens, "_anchor", None) is None:
        ens._anchor = torch.zeros((), device=ens.device, requires_grad=True)
    _DIRTY.add(ens)
    return _FusedLoss.apply(ens._anchor, ens, total)
