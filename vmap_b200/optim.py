"""Optimiser with torch.optim.AdamW's protocol as train.py uses it (train.py:67,150-160,
324-326): ``add_param_group``, ``step()``, ``zero_grad(set_to_none=True)``.  ``step`` is
one fused AdamW launch (K2) per ensemble that received gradients since the last step --
the same "tensors without a grad are skipped" rule torch applies."""
from __future__ import annotations

from .lazy import _DIRTY


class FusedAdamW:
    def __init__(self, params=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = []
        if params is not None:
            self.add_param_group({"params": list(params)})

    def add_param_group(self, group):
        g = dict(self.defaults)
        g.update(group)
        self.param_groups.append(g)

    def step(self, closure=None):
        for ens in list(_DIRTY):
            ens.lr, ens.weight_decay = self.defaults["lr"], self.defaults["weight_decay"]
            ens.betas, ens.eps = self.defaults["betas"], self.defaults["eps"]
            ens.poll_status()             # the reference exit(-1)s on a loss explosion (render_rays.py:88-90)
            ens.adam_step()
        _DIRTY.clear()

    def zero_grad(self, set_to_none=True):
        for ens in list(_DIRTY):             # gradients produced but never stepped
            ens.grads.zero_()
            ens._grad_scale = None
        _DIRTY.clear()


AdamW = FusedAdamW
