"""Oracle vs the UNMODIFIED reference modules, live (build container only: skipped where /root/reference is absent,
e.g. on the GPU box).  The committed goldens pin a few fixed cases; this sweeps more shapes / seeds through the same
harness train.py uses (functorch vmap over combine_state_for_ensemble, loss.step_batch_loss, torch.optim.AdamW)."""
import pytest
import torch

from oracle import _refload
from oracle import vmap_oracle as vo
from tests._util import rel_l2

pytestmark = pytest.mark.skipif(not _refload.available(), reason="reference tree not mounted")


def _reference_step(n_obj, hidden, n_rays, n_samples, scale, n_cam2surf, seed, n_steps):
    model, embedding, render_rays, loss = _refload.load("model", "embedding", "render_rays", "loss")
    from functorch import combine_state_for_ensemble, vmap
    torch.manual_seed(seed)
    e1, e2 = vo.emb_sizes(5)
    fcs = [model.OccupancyMap(e1, e2, hidden_size=hidden).apply(model.init_weights) for _ in range(n_obj)]   # trainer.py:27-33
    pes = [embedding.UniDirsEmbed(max_deg=5, scale=scale) for _ in range(n_obj)]
    batch = vo.synthetic_batch(n_obj, n_rays, n_samples, seed=seed + 100, n_cam2surf=n_cam2surf)
    opt = torch.optim.AdamW([torch.zeros(1, requires_grad=True)], lr=1e-3, weight_decay=0.013)             # train.py:67
    fc_model, fc_param, fc_buffer = combine_state_for_ensemble(fcs)                                        # utils.py:31
    pe_model, pe_param, pe_buffer = combine_state_for_ensemble(pes)
    for p in list(fc_param) + list(pe_param):
        p.requires_grad_()
    opt.add_param_group({"params": fc_param}); opt.add_param_group({"params": pe_param})
    names = [n for n, _ in fcs[0].named_parameters()]
    init = {n: p.detach().clone() for n, p in zip(names, fc_param)}
    init[vo.PE_KEY] = pe_param[0].detach().clone()

    def fwd_loss():
        emb = vmap(pe_model)(pe_param, pe_buffer, batch["pcs"])                                            # train.py:293
        alpha, col = vmap(fc_model)(fc_param, fc_buffer, emb)                                              # train.py:294
        return loss.step_batch_loss(alpha, col, batch["gt_depth"], batch["gt_colour"], batch["sem"],
                                    batch["mask_depth"], batch["z"])[0]                                    # train.py:303
    l0 = fwd_loss()
    l0.backward()
    grads = {n: p.grad.detach().clone() for n, p in zip(names, fc_param)}
    grads[vo.PE_KEY] = pe_param[0].grad.detach().clone()
    losses = [float(l0)]
    for s in range(n_steps):
        opt.step(); opt.zero_grad(set_to_none=True)
        l = fwd_loss(); losses.append(float(l))
        if s + 1 < n_steps:
            l.backward()
    final = {n: p.detach().clone() for n, p in zip(names, fc_param)}
    final[vo.PE_KEY] = pe_param[0].detach().clone()
    return batch, init, grads, losses, final


@pytest.mark.parametrize("n_obj,hidden,n_rays,n_samples,scale,n1,seed", [
    (1, 32, 17, 6, 2.0, 1, 21), (5, 32, 9, 10, 2.0, 1, 22), (2, 64, 13, 14, 5.0, 5, 23), (1, 128, 8, 10, 5.0, 5, 24),
    (3, 32, 1, 10, 2.0, 1, 25)])
def test_oracle_matches_live_reference_step(n_obj, hidden, n_rays, n_samples, scale, n1, seed):
    n_steps = 2
    batch, init, g_ref, l_ref, p_ref = _reference_step(n_obj, hidden, n_rays, n_samples, scale, n1, seed, n_steps)
    orc = vo.OracleEnsemble(init, scale)
    loss, grads = orc.grads(batch)
    assert abs(float(loss) - l_ref[0]) <= 2e-6 * abs(l_ref[0]) + 1e-7
    for k in vo.ALL_KEYS:
        assert rel_l2(grads[k], g_ref[k]) < 2e-5, k
    orc = vo.OracleEnsemble(init, scale)            # fresh optimiser state / no leftover .grad
    losses = [float(orc.step(batch)) for _ in range(n_steps)]
    for a, b in zip(losses, l_ref[:n_steps]):
        assert abs(a - b) <= 5e-6 * abs(b) + 1e-7
    for k in vo.ALL_KEYS:
        assert rel_l2(orc.params[k], p_ref[k]) < 2e-6, k
