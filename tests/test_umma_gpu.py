"""GPU parity of the tcgen05 (fp16-operand) step kernel: BASELINE bar is <= 1e-3 relative L2
on rendered depth / colour vs the reference on identical inputs."""
import pytest
import torch

from oracle import vmap_oracle as vo
from tests._util import load_step_golden, make_ensemble, rel_l2, to_dev

pytestmark = pytest.mark.gpu

TOL_RENDER = 1e-3      # BASELINE.json north_star: rendered depth/colour within 1e-3 relative L2
TOL_GRAD = 4e-2        # fp16 dY operands (static loss scale 2^8), fp32 accumulation


def cos_sim(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def test_golden_render_umma():
    g, params, batch = load_step_golden("step_vmap_h32")
    ens = make_ensemble(params, float(g["scale"]), 32, impl="umma")
    d, v, c, o = ens.render(to_dev(batch))
    assert rel_l2(d, g["r_depth"]) < TOL_RENDER
    assert rel_l2(c, g["r_colour"]) < TOL_RENDER
    assert rel_l2(o, g["r_opacity"]) < TOL_RENDER


@pytest.mark.parametrize("cfg", [
    dict(B=4, R=301, S=10, n1=1),       # ragged last tile; CTAs straddle objects
    dict(B=5, R=30, S=10, n1=1),        # 3 tiles / object: every CTA spans two objects
    dict(B=2, R=100, S=14, n1=5),       # 5+9 samples
    dict(B=1, R=64, S=16, n1=5),        # exactly 8 rays / tile
    dict(B=3, R=40, S=1, n1=0),         # single sample per ray
    dict(B=2, R=33, S=20, n1=5),        # one ray per warp, 12 idle lanes
    dict(B=2, R=50, S=32, n1=8),        # BASELINE configs[4]'s 32 samples per ray on the fused kernel
    dict(B=20, R=1200, S=10, n1=1),     # BASELINE cfg 2 at full size, against the oracle
])
def test_oracle_parity_umma(cfg):
    B, R, S = cfg["B"], cfg["R"], cfg["S"]
    params = vo.init_params(B, 32, seed=7)
    if S > 1:
        batch = vo.synthetic_batch(B, R, S, seed=11, n_cam2surf=cfg["n1"])
    else:
        b2 = vo.synthetic_batch(B, R, 2, seed=5, n_cam2surf=1)
        batch = {"pcs": b2["pcs"][:, :, :1].contiguous(), "z": b2["z"][:, :, :1].contiguous(),
                 "gt_depth": b2["gt_depth"], "gt_colour": b2["gt_colour"], "sem": b2["sem"],
                 "mask_depth": b2["mask_depth"]}
    db = to_dev(batch)
    orc = vo.OracleEnsemble(params, 2.0)          # the oracle is the reference at every size
    loss_ref, g_ref = orc.grads(batch)
    d_ref, _, c_ref, o_ref = orc.render(batch)
    ens = make_ensemble(params, 2.0, 32, impl="umma")
    d, v, c, o = ens.render(db)
    errs = dict(depth=rel_l2(d, d_ref), colour=rel_l2(c, c_ref), opacity=rel_l2(o, o_ref))
    print("render rel-L2", errs)
    assert max(errs.values()) < TOL_RENDER, errs
    ens.forward_backward(db)
    loss = float(ens.loss_terms[:, 3].sum())
    assert abs(loss - float(loss_ref)) < 2e-3 * abs(float(loss_ref))
    got = ens.stacked(ens.grads)
    gerr = {k: (rel_l2(got[k], g_ref[k]), cos_sim(got[k], g_ref[k])) for k in vo.ALL_KEYS}
    print("grad (rel-L2, cos)", gerr)
    for k, (e, cs) in gerr.items():
        assert e < TOL_GRAD and cs > 0.999, (k, e, cs)


def test_umma_training_tracks_fp32_kernel():
    """200 AdamW steps on identical inputs: the fp16-operand path must follow the fp32 path."""
    B, R, S = 4, 240, 10
    params = vo.init_params(B, 32, seed=21)
    batches = [to_dev(vo.synthetic_batch(B, R, S, seed=100 + i)) for i in range(8)]
    a = make_ensemble(params, 2.0, 32, impl="fp32")
    u = make_ensemble(params, 2.0, 32, impl="umma")
    for it in range(200):
        a.step(batches[it % 8]); u.step(batches[it % 8])
    la = torch.stack([a.render(b)[0] for b in batches])
    lu = torch.stack([u.render(b)[0] for b in batches])
    a.forward_backward(batches[0], backward=False); u.forward_backward(batches[0], backward=False)
    l_a, l_u = float(a.loss_terms[:, 3].sum()), float(u.loss_terms[:, 3].sum())
    print("loss after 200 steps fp32/umma", l_a, l_u, "depth rel-L2", rel_l2(lu, la))
    # random (noise) targets make the trajectory chaotic; the bar here is only 'tracks closely'.
    # Trained-quality parity (PSNR within 0.2 dB) is tested on a consistent scene in test_train_gpu.py
    assert abs(l_a - l_u) < 0.05 * abs(l_a)
    assert rel_l2(lu, la) < 0.15
    u.check_status()


@pytest.mark.parametrize("N", [1000, 70001])
def test_eval_points_tensor_core_forward_matches_oracle(N):
    """Trainer.eval_points (trainer.py:77-90) through the forward half of the fused tcgen05 kernel."""
    B = 3
    params = vo.init_params(B, 32, seed=9)
    pts = (torch.rand(B, N, 3, generator=torch.Generator().manual_seed(1)) - 0.5) * 4
    alpha_ref, col_ref = vo.forward(params, torch.full((B,), 2.0), pts.view(B, N, 1, 3))
    ens = make_ensemble(params, 2.0, 32, impl="umma")
    alpha, col = ens.eval_points(pts.cuda())
    a32, c32 = ens.eval_points(pts.cuda(), impl="fp32")
    assert rel_l2(a32, alpha_ref.view(B, N)) < 1e-5
    print("eval_points tcgen05 vs oracle: alpha", rel_l2(alpha, alpha_ref.view(B, N)), "colour", rel_l2(col, col_ref.view(B, N, 3)))
    assert rel_l2(alpha, alpha_ref.view(B, N)) < 2e-3
    assert rel_l2(col, col_ref.view(B, N, 3)) < 1e-3


def test_eval_points_throughput_grid():
    """256^3-class query (meshing, trainer.py:35-75): tensor-core forward vs the fp32 CUDA-core kernel."""
    B, N = 4, 128 ** 3
    params = vo.init_params(B, 32, seed=3)
    pts = (torch.rand(B, N, 3, device="cuda") - 0.5) * 4
    ens = make_ensemble(params, 2.0, 32, impl="umma")
    res = {}
    for impl in ("umma", "fp32"):
        for _ in range(2):
            ens.eval_points(pts, impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            a, c = ens.eval_points(pts, impl=impl)
        e1.record(); torch.cuda.synchronize()
        res[impl] = e0.elapsed_time(e1) / 3
        print(f"eval_points {B} x {N} points, {impl}: {res[impl]:.2f} ms -> {B * N / res[impl] / 1e6:.2f} G points/s")
    assert res["umma"] < res["fp32"]


def test_step_is_reproducible(monkeypatch):
    """No floating-point atomics on the hidden-32 path: per-(CTA, object) gradient partials are reduced in segment
    order by the last CTA to finish the object.  With VMB_DETERMINISTIC=1 (one point group per CTA, so every wgrad
    accumulator sees a single in-order MMA stream) two runs from the same state give identical bits over 200 steps;
    in the default mode (two groups interleave their wgrad MMAs in arrival order) a single step agrees to fp32
    rounding of the summation order."""
    B, R, S = 7, 301, 10                      # CTAs straddle objects, ragged last tile
    params = vo.init_params(B, 32, seed=3)
    batches = [to_dev(vo.synthetic_batch(B, R, S, seed=50 + i)) for i in range(4)]
    monkeypatch.setenv("VMB_DETERMINISTIC", "1")
    outs = []
    for run in range(2):
        ens = make_ensemble(params, 2.0, 32, impl="umma")
        losses = []
        for it in range(200):
            losses.append(ens.step(batches[it % 4]))
        ens.check_status()
        outs.append((ens.params.clone(), ens.exp_avg_sq.clone(), torch.stack(losses)))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # the split protocol (backward into ens.grads, AdamW as a second call) reduces in the same order
    e1 = make_ensemble(params, 2.0, 32, impl="umma")
    e2 = make_ensemble(params, 2.0, 32, impl="umma")
    for it in range(5):
        e1.step(batches[it % 4])
        e2.forward_backward(batches[it % 4]); e2.adam_step()
    assert torch.equal(e1.params, e2.params) and torch.equal(e1.image, e2.image)
    monkeypatch.delenv("VMB_DETERMINISTIC")
    g = []
    for run in range(2):
        e = make_ensemble(params, 2.0, 32, impl="umma")
        e.forward_backward(batches[0])
        g.append(e.grads.clone())
    assert rel_l2(g[0], g[1]) < 1e-5


@pytest.mark.parametrize("impl", ["umma", "fp32"])
def test_step_returns_scalar_loss_without_a_reduction_launch(impl):
    """``vmb_step_args.loss_sum``: the step writes sum_b loss_terms[b][3] (loss.py:59-62 followed by train.py's
    ``batch_loss`` scalar) itself -- last CTA of the fused kernel, or a one-warp kernel behind the other paths."""
    B, R, S = 5, 130, 10
    params = vo.init_params(B, 32, seed=9)
    ens = make_ensemble(params, 2.0, 32, impl=impl)
    got, want = [], []
    for it in range(3):
        got.append(ens.step(to_dev(vo.synthetic_batch(B, R, S, seed=80 + it))))
        want.append(ens.loss_terms[:, 3].double().sum().item())
    ens.check_status()
    assert len({g.data_ptr() for g in got}) == 3                  # ring slots: earlier results are not overwritten
    for g, w in zip(got, want):
        assert abs(float(g) - w) <= 1e-5 * abs(w)
    out = torch.zeros(3, device=ens.device)
    ens.step(to_dev(vo.synthetic_batch(B, R, S, seed=90)), loss_out=out[1:2])
    assert out[0] == 0 and out[2] == 0 and abs(float(out[1]) - ens.loss_terms[:, 3].sum().item()) <= 1e-4 * float(out[1])
