"""Trainer with the reference's attributes (trainer.py:9-33): ``fc_occ_map``, ``pe``,
``obj_scale``, ``hidden_feature_size``, ``emb_size1``, ``emb_size2``, ``bound_extent``;
``eval_points`` runs the batched forward-only kernel (trainer.py:77-95)."""
from __future__ import annotations

import numpy as np
import torch

from . import embedding, model
from .lazy import ensemble_for_modules


class Trainer:
    def __init__(self, cfg):
        self.obj_id = cfg.obj_id
        self.device = cfg.training_device
        self.hidden_feature_size = cfg.hidden_feature_size
        self.obj_scale = cfg.obj_scale
        self.n_unidir_funcs = cfg.n_unidir_funcs
        self.emb_size1 = 21 * (3 + 1) + 3
        self.emb_size2 = 21 * (self.n_unidir_funcs + 1) + 3 - self.emb_size1
        self.load_network()
        self.bound_extent = 0.995 if self.obj_id == 0 else 0.9

    def load_network(self):
        self.fc_occ_map = model.OccupancyMap(self.emb_size1, self.emb_size2, hidden_size=self.hidden_feature_size)
        self.fc_occ_map.apply(model.init_weights).to(self.device)
        self.pe = embedding.UniDirsEmbed(max_deg=self.n_unidir_funcs, scale=self.obj_scale).to(self.device)

    def eval_points(self, points, chunk_size=100000):
        """(occupancy [N], colour [N,3]) or None when everything is empty (trainer.py:77-95).
        ``chunk_size`` bounds the points per launch (at least 131072: the kernel tiles internally)."""
        ens = ensemble_for_modules(self.fc_occ_map, self.pe)
        row = self.fc_occ_map._vmb_binding[1]
        pts = points.to(ens.device, torch.float32).reshape(-1, 3).contiguous()
        # one-row call: only THIS object's network runs, however many objects share the packed stack
        alpha, colour = ens.eval_points(pts, row=row, chunk=max(int(chunk_size), 1 << 17))
        occ = torch.sigmoid(alpha)
        if float(occ.max()) == 0:
            print("no occ")
            return None
        return occ, colour

    def meshing(self, bound, obj_center, grid_dim=256):
        """Marching-cubes meshing (trainer.py:35-75) is visualisation (skimage + trimesh on the CPU) and out of this
        path's scope; the part of it that runs the network -- ``eval_points`` on
        ``make_3D_grid`` points -- is provided above."""
        raise NotImplementedError("meshing is outside the accelerated path: use eval_points(make_3D_grid(...)) "
                                  "and run marching cubes with the reference's own trainer.meshing")


def make_3D_grid(occ_range=(-1., 1.), dim=256, device="cuda:0", transform=None, scale=None):
    """render_rays.make_3D_grid (render_rays.py:98-122): dim^3 query points."""
    t = torch.linspace(occ_range[0], occ_range[1], steps=dim, device=device)
    g = torch.stack(torch.meshgrid(t, t, t, indexing="ij"), dim=-1)
    if scale is not None:
        g = g * scale
    if transform is not None:
        g = g @ transform[:3, :3].T + transform[:3, 3]
    return g
