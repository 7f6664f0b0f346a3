"""CPU / gloo (world_size 2) tests of the multi-GPU host logic: object sharding, and the iMAP
ray-sharded step (count all-reduce -> local fwd/bwd with global normalisers -> grad all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vmap_oracle as vo
from vmap_b200.dist import ReplicatedStep, shard_objects


def test_shard_objects_partitions_contiguously():
    for n in (1, 7, 20, 160, 161):
        for w in (1, 2, 3, 8):
            spans = [shard_objects(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_objects(160, 8, 3) == (60, 80)


class OracleBackedEnsemble:
    """Stand-in with the VmapEnsemble interface, computing with the CPU oracle: lets the
    collective logic of ReplicatedStep run under gloo without a GPU."""

    def __init__(self, params, scale):
        self.orc = vo.OracleEnsemble(params, scale)
        b = params[vo.PE_KEY].shape[0]
        self.grads = torch.zeros(b, sum(v[0].numel() for v in params.values()))
        self.loss_terms = torch.zeros(b, 4)

    def mask_counts(self, batch):
        m_o, m_s = batch["sem"] != 0, batch["sem"] != 2
        m_d = batch["mask_depth"].bool() & m_o
        return torch.stack([m_d.sum(-1), m_o.sum(-1), m_s.sum(-1), torch.zeros_like(m_o.sum(-1))], -1).to(torch.int32)

    def forward_backward(self, batch, counts=None):
        p = self.orc.params
        alpha, col = vo.forward(p, self.orc.scale, batch["pcs"])
        depth, var, colr, opa = vo.render_outputs(alpha, col, batch["z"])
        m_o, m_s = (batch["sem"] != 0).float(), (batch["sem"] != 2).float()
        m_d = batch["mask_depth"].float() * m_o
        n = counts.float()
        on = (counts[:, :3] != 0).all(0).float()
        l_d = on[0] * ((depth - batch["gt_depth"]).abs() * m_d / (var.detach().sqrt() + 1e-4)).sum(-1) / (n[:, 0] + 1e-10)
        l_c = on[1] * ((colr - batch["gt_colour"]).abs().sum(-1) * m_o).sum(-1) / (n[:, 1] + 1e-10)
        l_o = on[2] * ((opa - m_o).abs() * m_s).sum(-1) / (n[:, 2] + 1e-10)
        tot = l_d + 5.0 * l_c + 10.0 * l_o
        gr = torch.autograd.grad(tot.sum(), [p[k] for k in vo.ALL_KEYS])
        self.grads.copy_(torch.cat([g.reshape(g.shape[0], -1) for g in gr], 1))
        self.loss_terms.copy_(torch.stack([l_d, l_c, l_o, tot], -1).detach())

    def adam_step(self):
        off = 0
        for k in vo.ALL_KEYS:
            n = self.orc.params[k][0].numel()
            self.orc.params[k].grad = self.grads[:, off:off + n].reshape(self.orc.params[k].shape).clone()
            off += n
        self.orc.opt.step()
        self.orc.opt.zero_grad(set_to_none=True)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    params = vo.init_params(1, 32, seed=3)
    full = vo.synthetic_batch(1, 64, 10, seed=8)
    lo, hi = rank * 32, (rank + 1) * 32
    local = {k: v[:, lo:hi].contiguous() for k, v in full.items()}
    ens = OracleBackedEnsemble(params, 2.0)
    stepper = ReplicatedStep(ens)
    # the next step's mask counts ride on this step's packed all-reduce: 1 counts collective up front, then ONE per step
    losses = [float(stepper.step(local, next_batch=local)) for _ in range(3)]
    if rank == 0:
        ret["losses"] = losses
        ret["collectives"] = stepper.collectives
        ret["w"] = ens.orc.params["mid1.0.0.weight"].detach().clone()
    dist.destroy_process_group()


def test_ray_sharded_replicated_step_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    params = vo.init_params(1, 32, seed=3)
    full = vo.synthetic_batch(1, 64, 10, seed=8)
    ref = vo.OracleEnsemble(params, 2.0)
    ref_losses = [float(ref.step(full)) for _ in range(3)]
    assert ret["losses"] == pytest.approx(ref_losses, rel=1e-4)
    assert ret["collectives"] == 1 + 3
    w = ret["w"]
    err = float((w - ref.params["mid1.0.0.weight"].detach()).norm() / ref.params["mid1.0.0.weight"].norm())
    assert err < 1e-5
