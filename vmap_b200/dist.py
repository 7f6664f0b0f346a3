"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

vMAP mode: objects are independent (the loss is a plain sum over objects, loss.py:59-60; parameters
and Adam state are private), so the object range is partitioned contiguously across ranks and the
step needs NO collective -- `shard_objects` is all there is.

iMAP mode: one whole-scene model replicated on every rank, the rays of a step sharded.  The loss
normalisers (mask counts, render_rays.py:86) and the any-empty early-out (render_rays.py:68-73) are
global quantities, so the all-reduced counts are handed to K1 (`counts` argument of `vmb_step`); the
319 811-float gradient (H=256), the loss terms and the next step's counts are all-reduced (sum) as ONE packed
buffer before the fused AdamW -- one collective per step.
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
    """Ray-sharded data-parallel step for a replicated ensemble (iMAP whole-scene model, BASELINE configs[4]).

    ONE collective per step: gradients, the per-object loss terms and the NEXT step's mask counts travel in one packed
    fp32 buffer (``ens.grads`` and ``ens.loss_terms`` are re-pointed at views of it, so nothing is copied):

        [ grads  B x stride | loss terms  B x 4 | next-step mask counts  B x 4 (exact in fp32 below 2^24) ]

    The loss normalisers (render_rays.py:86) and the any-empty early-out (render_rays.py:68-73) are global quantities
    that the step kernel needs BEFORE it runs; they depend only on the inputs, so the counts of step i+1 ride on the
    all-reduce of step i (``next_batch``).  Without ``next_batch`` (first step, or a caller that does not know the next
    batch yet) the counts get their own small all-reduce.

    ``ens`` needs the VmapEnsemble interface: mask_counts(batch) -> int32 [B,4]; forward_backward(batch, counts=...)
    accumulating into ``ens.grads`` and writing ``ens.loss_terms``; adam_step()."""

    def __init__(self, ens, group=None):
        self.ens, self.group = ens, group
        B, stride = ens.grads.shape
        self.pack = torch.zeros(B * stride + 8 * B, dtype=torch.float32, device=ens.grads.device)
        self.pack[:B * stride].copy_(ens.grads.reshape(-1))
        ens.grads = self.pack[:B * stride].view(B, stride)
        ens.loss_terms = self.pack[B * stride:B * stride + 4 * B].view(B, 4)
        self.next_counts_f = self.pack[B * stride + 4 * B:].view(B, 4)
        self.counts = torch.zeros(B, 4, dtype=torch.int32, device=ens.grads.device)
        self._counts_for = None               # id() of the batch self.counts was reduced for
        self.collectives = 0

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def allreduce_bytes(self) -> int:
        return self.pack.numel() * 4

    def step(self, local_batch, next_batch=None) -> torch.Tensor:
        ens = self.ens
        multi = self._world() > 1
        if self._counts_for != id(local_batch):
            self.counts.copy_(ens.mask_counts(local_batch))
            if multi:
                dist.all_reduce(self.counts, op=dist.ReduceOp.SUM, group=self.group)
                self.collectives += 1
        ens.forward_backward(local_batch, counts=self.counts)
        if next_batch is not None:
            self.next_counts_f.copy_(ens.mask_counts(next_batch))
        if multi:
            dist.all_reduce(self.pack, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        if next_batch is not None:
            self.counts.copy_(self.next_counts_f.round())
            self._counts_for = id(next_batch)
        else:
            self._counts_for = None
        ens.adam_step()
        return ens.loss_terms[:, 3].sum()
