"""Pin the CPU oracle (oracle/) against fixtures generated from the reference's own
modules (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so
from oracle import vmap_oracle as vo

STEP_CASES = ["step_vmap_h32", "step_bg_h128", "step_emptymask_h32"]


def load_step(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {k: torch.from_numpy(g["p0_" + k]) for k in vo.ALL_KEYS}
    batch = {k: torch.from_numpy(g["in_" + k]) for k in
             ("pcs", "z", "gt_depth", "gt_colour", "sem", "mask_depth")}
    return g, params, batch


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", STEP_CASES)
def test_forward_render_loss_match_reference(golden_dir, name):
    g, params, batch = load_step(golden_dir, name)
    ens = vo.OracleEnsemble(params, float(g["scale"]))
    with torch.no_grad():
        emb = vo.unidir_embed(batch["pcs"], ens.params[vo.PE_KEY], ens.scale)
        alpha, col = ens.forward(batch["pcs"])
        d, v, c, o = vo.render_outputs(alpha, col, batch["z"])
    assert rel_l2(emb[0, 0], g["emb_obj0_ray0"]) < 1e-6
    assert rel_l2(alpha, g["alpha"]) < 1e-5
    assert rel_l2(col, g["colour"]) < 1e-6
    assert rel_l2(d, g["r_depth"]) < 1e-6
    assert rel_l2(v, g["r_var"]) < 1e-5
    assert rel_l2(c, g["r_colour"]) < 1e-6
    assert rel_l2(o, g["r_opacity"]) < 1e-6
    loss = ens.loss(batch)
    assert abs(float(loss) - float(g["loss0"])) <= 2e-6 * abs(float(g["loss0"]))


@pytest.mark.parametrize("name", STEP_CASES)
def test_gradients_match_reference(golden_dir, name):
    g, params, batch = load_step(golden_dir, name)
    ens = vo.OracleEnsemble(params, float(g["scale"]))
    _, grads = ens.grads(batch)
    for k in vo.ALL_KEYS:
        assert rel_l2(grads[k], g["g0_" + k]) < 2e-5, k


@pytest.mark.parametrize("name", ["step_vmap_h32", "step_emptymask_h32"])
def test_adamw_trajectory_matches_reference(golden_dir, name):
    g, params, batch = load_step(golden_dir, name)
    n = int(g["n_steps"])
    ens = vo.OracleEnsemble(params, float(g["scale"]))
    losses = [float(ens.step(batch)) for _ in range(n)]
    losses.append(float(ens.loss(batch)))
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    for k in vo.ALL_KEYS:
        assert rel_l2(ens.params[k], g[f"p{n}_" + k]) < 2e-6, k


def test_emptymask_zeroes_depth_term_for_every_object(golden_dir):
    g, params, batch = load_step(golden_dir, "step_emptymask_h32")
    alpha, col = vo.forward(params, torch.full((2,), float(g["scale"])), batch["pcs"])
    l_d, l_c, l_o = vo.batch_loss_terms(alpha, col, batch["gt_depth"], batch["gt_colour"],
                                        batch["sem"], batch["mask_depth"], batch["z"])
    assert float(l_d.abs().sum()) == 0.0 and float(l_c.sum()) > 0 and float(l_o.sum()) > 0


def test_adamw_math_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000, dtype=torch.float64)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=0.013)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for t in range(1, 6):
        gr = torch.randn(1000, dtype=torch.float64)
        ref.grad = gr.clone()
        opt.step()
        p, m, v = vo.adamw_math(p, gr, m, v, t)
        assert rel_l2(p, ref.detach()) < 1e-14


def test_loss_explode_raises():
    B, R, S = 1, 4, 3
    alpha = torch.full((B, R, S, 1), 100.0)      # occ=1 at the first sample -> zero variance
    col = torch.zeros(B, R, S, 3)
    z = torch.ones(B, R, S)                       # weight 1/(0+1e-4)
    with pytest.raises(vo.LossExplode):
        vo.step_batch_loss(alpha, col, torch.full((B, R), 1e3), torch.zeros(B, R, 3),
                           torch.ones(B, R, dtype=torch.uint8), torch.ones(B, R, dtype=torch.bool), z)


@pytest.mark.parametrize("name", ["sampler_obj", "sampler_bg", "sampler_2kf"])
def test_sampler_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    t = {k: torch.from_numpy(g[k]) for k in ("rgbs_batch", "depth_batch", "t_wc_batch", "bbox", "rays_dir")}
    cfg = so.SamplerCfg(n_bins_cam2surface=int(g["n1"]))
    torch.manual_seed(int(g["seed"]))
    rnd = so.draw_randoms_reference_order(None, int(g["n_kf"]), [int(x) for x in g["latest"]],
                                          int(g["n_frames"]), int(g["n_samples"]), t["bbox"],
                                          t["rgbs_batch"], t["depth_batch"], cfg)
    rgb, depth, valid, lab, pcs, z = so.sample_from_randoms(
        rnd, t["rgbs_batch"], t["depth_batch"], t["t_wc_batch"], t["bbox"], t["rays_dir"], cfg)
    assert np.array_equal(rgb.numpy(), g["o_rgb"])
    assert np.array_equal(depth.numpy(), g["o_depth"])
    assert np.array_equal(valid.numpy(), g["o_valid"])
    assert np.array_equal(lab.numpy(), g["o_lab"])
    assert np.array_equal(z.numpy(), g["o_z"])            # same fp32 op order -> bit exact
    np.testing.assert_allclose(pcs.numpy(), g["o_pcs"], rtol=0, atol=1e-6)
