"""Synthetic Replica-shaped inputs and reference-distribution initial weights for benchmarks and demos.

Product-side utilities (bench.py, tools/): random ensembles with the reference's init distribution
(xavier-normal weights, torch.nn.Linear default biases, icosahedron PE directions; model.py:4-6,
trainer.py:32, embedding.py:51-76) and training batches whose sample depths follow the reference's
depth-guided strategy (vmap.py:366-459) in closed form.  No oracle code is imported here.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .embedding import ICOSAHEDRON_DIRS
from .layout import PE_KEY, tensor_shapes


def icosahedron_dirs(dtype=torch.float32) -> torch.Tensor:
    return torch.tensor(ICOSAHEDRON_DIRS, dtype=dtype)


def param_shapes(hidden: int, max_deg: int = 5):
    return tensor_shapes(hidden, max_deg)


def init_params(n_obj: int, hidden: int, max_deg: int = 5, seed: int = 0,
                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random ensemble with the reference's init distribution: xavier-normal
    weights (model.py:4-6, trainer.py:32), torch.nn.Linear default bias
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)), PE = icosahedron (embedding.py:76)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in param_shapes(hidden, max_deg).items():
        if k == PE_KEY:
            out[k] = icosahedron_dirs(dtype).expand(n_obj, -1, -1).clone()
        elif k.endswith("weight"):
            fan_out, fan_in = shp
            std = math.sqrt(2.0 / (fan_in + fan_out))
            out[k] = (torch.randn((n_obj,) + shp, generator=g) * std).to(dtype)
        else:
            wshape = param_shapes(hidden, max_deg)[k[:-4] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1])
            out[k] = ((torch.rand((n_obj,) + shp, generator=g) * 2 - 1) * bound).to(dtype)
    return out



def synthetic_batch(n_obj: int, n_rays: int, n_samples: int, seed: int = 0,
                    n_cam2surf: int = 1, dtype=torch.float32, empty_prob=(0.1, 0.3, 0.6, 0.1)):
    """Replica-shaped synthetic training batch (BASELINE.md section 3 / SURVEY.md 8d):
    depth U(0.5,4.5) with 10% invalid, labels p=(0.3,0.6,0.1), z drawn with the
    reference's depth-guided strategy (vmap.py:366-459) in closed form, pcs = o + d*z."""
    g = torch.Generator().manual_seed(seed)
    B, R, S = n_obj, n_rays, n_samples
    n1 = n_cam2surf
    n2 = S - n1
    depth = torch.rand(B, R, generator=g) * 4.0 + 0.5
    invalid = torch.rand(B, R, generator=g) < empty_prob[0]
    depth = torch.where(invalid, torch.zeros_like(depth), depth)
    u = torch.rand(B, R, generator=g)
    sem = torch.where(u < empty_prob[1], 0, torch.where(u < empty_prob[1] + empty_prob[2], 1, 2)).to(torch.uint8)
    rgb = torch.randint(0, 256, (B, R, 3), generator=g).to(torch.float32) / 255.0
    eps, other_eps = 0.1, 0.05
    maxb = depth.max(dim=1, keepdim=True).values
    ur = torch.rand(B, R, S, generator=g)
    lin1 = torch.arange(n1).view(1, 1, -1)
    lin2 = torch.arange(n2).view(1, 1, -1)
    linS = torch.arange(S).view(1, 1, -1)
    z = torch.empty(B, R, S)
    hi = (depth - eps)[..., None]
    z[..., :n1] = (lin1 + ur[..., :n1]) * hi / n1
    nrm = (torch.randn(B, R, n2, generator=g) * (eps / 3)).sort(-1).values.clamp(-eps, eps)
    z_this = depth[..., None] + nrm
    z_other = (depth - eps)[..., None] + (lin2 + ur[..., n1:]) * (eps + other_eps) / n2
    z[..., n1:] = torch.where((sem == 1)[..., None], z_this, z_other)
    z_inv = (linS + ur) * maxb[..., None] / S
    z = torch.where(invalid[..., None], z_inv, z)
    origin = (torch.rand(B, R, 3, generator=g) - 0.5)
    px = torch.rand(B, R, 2, generator=g)
    dirs = torch.stack([(px[..., 0] * 1200 - 599.5) / 600.0, (px[..., 1] * 680 - 339.5) / 600.0,
                        torch.ones(B, R)], -1)
    pcs = origin[..., None, :] + dirs[..., None, :] * z[..., None]
    return {
        "pcs": pcs.to(dtype).contiguous(), "z": z.to(dtype).contiguous(),
        "gt_depth": depth.to(dtype).contiguous(), "gt_colour": rgb.to(dtype).contiguous(),
        "sem": sem.contiguous(), "mask_depth": (~invalid).contiguous(),
    }
