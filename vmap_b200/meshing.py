"""Marching-cubes meshing of one object (trainer.py:35-75, vis.py:6-19).  Visualisation
output, outside the accelerated path; only the grid evaluation runs on the GPU kernels."""
from __future__ import annotations

import numpy as np
import torch


def mesh_object(trainer, bound, obj_center, grid_dim=256):
    import skimage.measure
    import trimesh
    from .trainer import make_3D_grid
    rng = (-1.0, 1.0)
    scale_np = bound.extent / ((rng[1] - rng[0]) * trainer.bound_extent)
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3], tf[:3, :3] = bound.center, bound.R
    dev = trainer.device
    grid = make_3D_grid(rng, grid_dim, dev, torch.from_numpy(tf).to(dev),
                        torch.from_numpy(scale_np).float().to(dev)).view(-1, 3)
    grid = grid - obj_center.to(grid.device)
    ret = trainer.eval_points(grid)
    if ret is None:
        return None
    occ = ret[0].view(grid_dim, grid_dim, grid_dim).cpu().numpy()
    try:
        v, f, n, _ = skimage.measure.marching_cubes(occ, level=0.5, gradient_direction="ascent")
    except (ValueError, RuntimeError):
        print("marching cube failed")
        return None
    mesh = trimesh.Trimesh(vertices=v / (grid_dim - 1), vertex_normals=n, faces=f)
    mesh.apply_translation([-0.5, -0.5, -0.5]); mesh.apply_scale(2)
    mesh.apply_scale(scale_np); mesh.apply_transform(tf)
    ret = trainer.eval_points(torch.from_numpy(np.array(mesh.vertices)).float().to(dev))
    if ret is None:
        return None
    mesh.visual.vertex_colors = (ret[1] * 255).detach().cpu().numpy().astype(np.uint8)
    return mesh
