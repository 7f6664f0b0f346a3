"""A consistent analytic scene for trained-quality parity (TEST INFRASTRUCTURE ONLY).

Each object is a sphere with a smooth colour field; rays that miss it hit a far wall and are
labelled 'other object'.  Sample depths follow the reference's depth-guided strategy
(vmap.py:366-459) so the training signal has the same structure as in vMAP.
"""
from __future__ import annotations

import torch


def sphere_batch(n_obj: int, n_rays: int, n_samples: int, seed: int, n_cam2surf: int = 1):
    g = torch.Generator().manual_seed(seed)
    B, R, S = n_obj, n_rays, n_samples
    n1, n2 = n_cam2surf, n_samples - n_cam2surf
    radius = 0.45 + 0.1 * torch.arange(B).float().view(B, 1) / max(B, 1)
    # camera origins on a shell of radius 2, looking at a jittered point near the centre
    o = torch.randn(B, R, 3, generator=g)
    o = 2.0 * o / o.norm(dim=-1, keepdim=True)
    tgt = (torch.rand(B, R, 3, generator=g) - 0.5) * 1.4
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    # ray-sphere intersection
    bq = (o * d).sum(-1)
    cq = (o * o).sum(-1) - radius ** 2
    disc = bq * bq - cq
    hit = disc > 0
    t_hit = -bq - torch.sqrt(disc.clamp_min(0))
    wall = 3.5 * torch.ones(B, R)
    depth = torch.where(hit, t_hit, wall)
    sem = torch.where(hit, 1, 0).to(torch.uint8)
    unknown = torch.rand(B, R, generator=g) < 0.05
    sem = torch.where(unknown, torch.full_like(sem, 2), sem)
    invalid = torch.rand(B, R, generator=g) < 0.05
    depth = torch.where(invalid, torch.zeros_like(depth), depth)
    x = o + d * depth[..., None]
    col_obj = 0.5 + 0.5 * torch.sin(3.0 * x + torch.tensor([0.0, 2.0, 4.0]))
    col = torch.where((sem == 1)[..., None], col_obj, torch.full_like(col_obj, 0.2))
    eps, oeps = 0.1, 0.05
    u = torch.rand(B, R, S, generator=g)
    k1, k2, kS = torch.arange(n1).view(1, 1, -1), torch.arange(n2).view(1, 1, -1), torch.arange(S).view(1, 1, -1)
    z = torch.empty(B, R, S)
    z[..., :n1] = (k1 + u[..., :n1]) * (depth - eps)[..., None] / max(n1, 1)
    nrm = (torch.randn(B, R, n2, generator=g) * (eps / 3)).sort(-1).values.clamp(-eps, eps)
    z_this = depth[..., None] + nrm
    z_other = (depth - eps)[..., None] + (k2 + u[..., n1:]) * (eps + oeps) / n2
    z[..., n1:] = torch.where((sem == 1)[..., None], z_this, z_other)
    z_inv = (kS + u) * depth.max(dim=1, keepdim=True).values[..., None] / S
    z = torch.where(invalid[..., None], z_inv, z)
    pcs = o[..., None, :] + d[..., None, :] * z[..., None]
    return {"pcs": pcs.contiguous(), "z": z.contiguous(), "gt_depth": depth.contiguous(),
            "gt_colour": col.contiguous(), "sem": sem.contiguous(), "mask_depth": (~invalid).contiguous()}


def quality(depth, colour, batch):
    """(colour PSNR over this-object rays, mean |depth error| over valid this-object rays)."""
    m = (batch["sem"] == 1)
    mse = ((colour - batch["gt_colour"]) ** 2)[m].mean()
    psnr = float(-10.0 * torch.log10(mse))
    md = m & batch["mask_depth"].bool()
    derr = float((depth - batch["gt_depth"]).abs()[md].mean())
    return psnr, derr
