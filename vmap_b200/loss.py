"""loss.step_batch_loss with the reference's signature (loss.py:5-62), fused: this call is
where the forward (PE + MLP), volume render, loss AND backward of the whole object stack are
launched (K0 + K1).  Argument order as called at train.py:303-306: labels before depth mask."""
from __future__ import annotations

import torch

from .ensemble import rows_dense
from .lazy import LazyHead, fused_loss


def step_batch_loss(alpha, color, gt_depth, gt_color, sem_labels, mask_depth, z_vals,
                    color_scaling=5.0, opacity_scaling=10.0):
    if not isinstance(alpha, LazyHead) or not isinstance(color, LazyHead) or alpha.heads is not color.heads:
        raise TypeError("step_batch_loss expects the (alpha, color) pair returned by the fused model call; "
                        "there is no eager PyTorch loss path")
    heads = alpha.heads
    ens = heads.ensemble()
    pcs = heads.emb.pcs
    if pcs.dim() == 3:                      # single module called on [R,S,3] (train.py:310-312)
        pcs = pcs[None]
    B = ens.n_obj

    def lead(t, nd):
        return t if t.dim() == nd else t[None]

    batch = {
        "pcs": pcs.float(), "z": lead(z_vals, 3), "gt_depth": lead(gt_depth, 2), "gt_colour": lead(gt_color, 3),
        "sem": lead(sem_labels, 2), "mask_depth": lead(mask_depth, 2),
    }
    assert batch["pcs"].shape[0] == B, "object count differs from the stacked ensemble"
    batch = {k: (v if rows_dense(v) else v.contiguous()) for k, v in batch.items()}
    if batch["sem"].dtype != torch.uint8:
        batch["sem"] = batch["sem"].to(torch.uint8)
    ens.colour_scaling, ens.opacity_scaling = float(color_scaling), float(opacity_scaling)
    slot = ens.loss_calls % ens.loss_ring.numel()      # the launch writes the scalar loss itself (no reduction launch)
    ens.loss_calls += 1
    total = ens.loss_ring[slot:slot + 1]
    ens.forward_backward(batch, loss_out=total)
    ens._last_batch = batch                 # inputs must outlive the asynchronous launch
    return fused_loss(ens, total), None
