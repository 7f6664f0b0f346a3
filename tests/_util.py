import os

import numpy as np
import torch

from oracle import vmap_oracle as vo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BATCH_KEYS = ("pcs", "z", "gt_depth", "gt_colour", "sem", "mask_depth")


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_step_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = {k: torch.from_numpy(g["p0_" + k]) for k in vo.ALL_KEYS}
    batch = {k: torch.from_numpy(g["in_" + k]) for k in BATCH_KEYS}
    return g, params, batch


def to_dev(batch, dev="cuda:0"):
    return {k: v.to(dev).contiguous() for k, v in batch.items()}


def make_ensemble(params, scale, hidden, impl="fp32", **kw):
    from vmap_b200.ensemble import VmapEnsemble
    n_obj = params[vo.PE_KEY].shape[0]
    ens = VmapEnsemble(n_obj, hidden=hidden, scale=scale, impl=impl, **kw)
    ens.load_stacked(params)
    return ens
