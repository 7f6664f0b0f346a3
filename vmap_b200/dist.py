"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

vMAP mode: objects are independent (the loss is a plain sum over objects, loss.py:59-60; parameters
and Adam state are private), so the object range is partitioned contiguously across ranks and the
step needs NO collective -- `shard_objects` is all there is.

iMAP mode: one whole-scene model replicated on every rank, the rays of a step sharded.  The loss
normalisers (mask counts, render_rays.py:86) and the any-empty early-out (render_rays.py:68-73) are
global quantities, so the counts are all-reduced first and handed to K1 (`counts` argument of
`vmb_step`); the 319 811-float gradient (H=256) is then all-reduced (sum) before the fused AdamW.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_objects(n_obj: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) object range of ``rank`` (sizes differ by at most 1)."""
    base, extra = divmod(n_obj, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class ReplicatedStep:
    """Ray-sharded data-parallel step for a replicated ensemble (iMAP whole-scene model).

    ``ens`` needs the VmapEnsemble interface: mask_counts(batch) -> int32 [B,4];
    forward_backward(batch, counts=...) accumulating into ``ens.grads`` and writing
    ``ens.loss_terms``; adam_step()."""

    def __init__(self, ens, group=None):
        self.ens, self.group = ens, group

    def step(self, local_batch) -> torch.Tensor:
        ens = self.ens
        counts = ens.mask_counts(local_batch)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
            ens.forward_backward(local_batch, counts=counts)
            dist.all_reduce(ens.grads, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(ens.loss_terms, op=dist.ReduceOp.SUM, group=self.group)
        else:
            ens.forward_backward(local_batch, counts=counts)
        ens.adam_step()
        return ens.loss_terms[:, 3].sum()
