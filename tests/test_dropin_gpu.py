"""GPU: the reference's train.py inner loop (train.py:180-338), written against the mirror
API exactly as train.py writes it, must train the same way as the CPU oracle."""
import os

import pytest
import torch

from oracle import vmap_oracle as vo
from tests._util import rel_l2

pytestmark = pytest.mark.gpu


def _small_cfg():
    from vmap_b200 import cfg as cfg_mod
    d = cfg_mod.replica_room0_dict()
    d["camera"].update(w=96, h=64, fx=60.0, fy=60.0, cx=47.5, cy=31.5)
    d["model"]["keyframe_buffer_size"] = 6
    d["model"]["keyframe_step"] = 1
    d["render"]["iters_per_frame"] = 4
    d["trainer"]["do_bg"] = 0
    return cfg_mod.Config(config_dict=d)


def _frame(cfg, g, n_obj):
    W, H = cfg.W, cfg.H
    rgb = torch.randint(0, 256, (W, H, 3), generator=g).to(torch.uint8).cuda()
    depth = (torch.rand(W, H, generator=g) * 3 + 0.8)
    depth[torch.rand(W, H, generator=g) < 0.1] = 0
    inst = (torch.rand(W, H, generator=g) * (n_obj + 1)).long().clamp(0, n_obj)      # 0..n_obj
    twc = torch.eye(4)
    twc[:3, 3] = torch.rand(3, generator=g) - 0.5
    return rgb, depth.cuda(), inst.cuda(), twc.cuda()


def test_train_loop_dropin_matches_oracle(tmp_path):
    import vmap_b200.loss as loss
    import vmap_b200.utils as utils
    from vmap_b200.optim import AdamW
    from vmap_b200.utils import vmap
    from vmap_b200.vmap import cameraInfo, sceneObject, sample_all

    torch.manual_seed(0)
    cfg = _small_cfg()
    cam_info = cameraInfo(cfg)
    g = torch.Generator().manual_seed(3)
    n_obj = 3
    optimiser = AdamW([torch.zeros(1)], lr=cfg.learning_rate, weight_decay=cfg.weight_decay)      # train.py:67
    obj_dict, fc_models, pe_models = {}, [], []
    for frame_id in range(3):                                                                      # train.py:95-164
        rgb, depth, inst, twc = _frame(cfg, g, n_obj)
        for obj_id in range(1, n_obj + 1):
            state = torch.zeros_like(inst, dtype=torch.uint8)
            state[inst == obj_id] = 1
            state[inst == 0] = 2
            bbox = torch.tensor([4.0, cfg.W - 5.0, 3.0, cfg.H - 4.0], device="cuda")
            if obj_id in obj_dict:
                obj_dict[obj_id].append_keyframe(rgb, depth, state, bbox, twc, frame_id)
            else:
                so = sceneObject(cfg, obj_id, rgb, depth, state, bbox, twc, frame_id)
                obj_dict[obj_id] = so
                optimiser.add_param_group({"params": so.trainer.fc_occ_map.parameters(), "lr": cfg.learning_rate})
                optimiser.add_param_group({"params": so.trainer.pe.parameters(), "lr": cfg.learning_rate})
                fc_models.append(so.trainer.fc_occ_map)
                pe_models.append(so.trainer.pe)
    init = {k: torch.stack([dict(m.named_parameters())[k].detach().cpu().clone() for m in fc_models])
            for k in vo.FC_KEYS}
    init[vo.PE_KEY] = torch.stack([p.B_layer.weight.detach().cpu().clone() for p in pe_models])
    fc_model, fc_param, fc_buffer = utils.update_vmap(fc_models, optimiser)                       # train.py:181
    pe_model, pe_param, pe_buffer = utils.update_vmap(pe_models, optimiser)                       # train.py:182

    # sampling: the reference's per-object loop (train.py:208-218) ...
    B_depth, B_rgb, B_dmask, B_omask, B_pcs, B_z = [], [], [], [], [], []
    for obj_id, obj_k in obj_dict.items():
        gt_rgb, gt_depth, valid, omask, pcs, z = obj_k.get_training_samples(
            cfg.n_iter_per_frame * cfg.win_size, cfg.n_samples_per_frame, cam_info.rays_dir_cache)
        assert gt_rgb.dtype == torch.uint8 and valid.dtype == torch.bool and omask.dtype == torch.uint8
        B_depth.append(gt_depth.reshape(-1)); B_rgb.append(gt_rgb.reshape(-1, 3)); B_dmask.append(valid)
        B_omask.append(omask); B_pcs.append(pcs.reshape(-1, pcs.shape[2], 3)); B_z.append(z.reshape(-1, z.shape[2]))
    B_pcs = torch.stack(B_pcs); B_depth = torch.stack(B_depth); B_rgb = torch.stack(B_rgb) / 255.  # train.py:255-260
    B_dmask = torch.stack(B_dmask); B_omask = torch.stack(B_omask); B_z = torch.stack(B_z)
    # ... and the one-launch replacement gives the same kind of batch
    allb = sample_all(list(obj_dict.values()), cfg.n_iter_per_frame * cfg.win_size, cfg.n_samples_per_frame,
                      cam_info.rays_dir_cache)
    assert allb["pcs"].shape == B_pcs.shape and allb["gt_colour"].shape == B_rgb.shape

    orc = vo.OracleEnsemble(init, cfg.obj_scale, lr=cfg.learning_rate, weight_decay=cfg.weight_decay)
    n = cfg.n_per_optim
    losses, ref_losses = [], []
    for it in range(cfg.n_iter_per_frame):                                                         # train.py:270-326
        idx = slice(it * n, (it + 1) * n)
        emb = vmap(pe_model)(pe_param, pe_buffer, B_pcs[:, idx, ...])
        alpha, color = vmap(fc_model)(fc_param, fc_buffer, emb)
        batch_loss, _ = loss.step_batch_loss(alpha, color, B_depth[:, idx].detach(), B_rgb[:, idx].detach(),
                                             B_omask[:, idx].detach(), B_dmask[:, idx].detach(), B_z[:, idx].detach())
        batch_loss.backward()
        optimiser.step()
        optimiser.zero_grad(set_to_none=True)
        losses.append(float(batch_loss))
        ref_losses.append(float(orc.step({"pcs": B_pcs[:, idx].cpu(), "z": B_z[:, idx].cpu(),
                                          "gt_depth": B_depth[:, idx].cpu(), "gt_colour": B_rgb[:, idx].cpu(),
                                          "sem": B_omask[:, idx].cpu(), "mask_depth": B_dmask[:, idx].cpu()})))
    print("loss", losses, "oracle", ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 5e-3 * abs(b)
    with torch.no_grad():                                                                          # train.py:331-338
        for model_id, (obj_id, obj_k) in enumerate(obj_dict.items()):
            for i, param in enumerate(obj_k.trainer.fc_occ_map.parameters()):
                param.copy_(fc_param[i][model_id])
            for i, param in enumerate(obj_k.trainer.pe.parameters()):
                param.copy_(pe_param[i][model_id])
    for model_id, obj_k in enumerate(obj_dict.values()):
        for k, p in obj_k.trainer.fc_occ_map.named_parameters():
            assert rel_l2(p, orc.params[k][model_id]) < 2e-2, k     # 4 AdamW steps on the fp16-operand path
    # checkpoint round trip with the reference's keys (vmap.py:461-491)
    first = next(iter(obj_dict.values()))
    first.save_checkpoints(str(tmp_path), 7)
    ck = torch.load(os.path.join(str(tmp_path), f"obj_{first.obj_id}_frame_7.pth"), weights_only=False)
    assert set(ck) == {"epoch", "FC_state_dict", "PE_state_dict", "obj_id", "bbox", "obj_scale"}
    assert tuple(ck["FC_state_dict"]) == vo.FC_KEYS and set(ck["PE_state_dict"]) == {"scale", "B_layer.weight"}
    before = first.trainer.fc_occ_map.in_layer[0].weight.detach().clone()
    with torch.no_grad():
        first.trainer.fc_occ_map.in_layer[0].weight.zero_()
    first.load_checkpoints(os.path.join(str(tmp_path), f"obj_{first.obj_id}_frame_7.pth"))
    assert torch.equal(before, first.trainer.fc_occ_map.in_layer[0].weight.detach())
    # forward-only inference (trainer.py:77-95)
    occ, colour = first.trainer.eval_points(torch.rand(500, 3, device="cuda") - 0.5)
    assert occ.shape == (500,) and colour.shape == (500, 3) and bool(((occ >= 0) & (occ <= 1)).all())


@pytest.mark.parametrize("impl", ["fp32", "auto"])
def test_separate_background_model_path(impl, monkeypatch):
    """do_bg: a non-vmapped H=128 model trained beside the stack (train.py:308-316).  "auto" resolves to the layer-wise
    tensor-core path (fp16 operands) for hidden 128; VMB_IMPL=fp32 pins the CUDA-core parity kernel."""
    monkeypatch.setenv("VMB_IMPL", impl)
    tol_loss, tol_par = (1e-4, 1e-4) if impl == "fp32" else (3e-3, 2e-2)
    import vmap_b200.loss as loss
    from vmap_b200.embedding import UniDirsEmbed
    from vmap_b200.model import OccupancyMap, init_weights
    from vmap_b200.optim import AdamW
    torch.manual_seed(1)
    fc = OccupancyMap(87, 42, hidden_size=128).apply(init_weights).cuda()
    pe = UniDirsEmbed(max_deg=5, scale=5.0).cuda()
    init = {k: v.detach().cpu().clone()[None] for k, v in fc.named_parameters()}
    init[vo.PE_KEY] = pe.B_layer.weight.detach().cpu().clone()[None]
    opt = AdamW([torch.zeros(1)], lr=1e-3, weight_decay=0.013)
    opt.add_param_group({"params": fc.parameters()})
    opt.add_param_group({"params": pe.parameters()})
    b = vo.synthetic_batch(1, 96, 14, seed=4, n_cam2surf=5)
    orc = vo.OracleEnsemble(init, 5.0)
    for _ in range(3):
        bg_embedding = pe(b["pcs"][0].cuda())
        bg_alpha, bg_color = fc(bg_embedding)
        bg_loss, _ = loss.step_batch_loss(bg_alpha[None, ...], bg_color[None, ...], b["gt_depth"].cuda(),
                                          b["gt_colour"].cuda(), b["sem"].cuda(), b["mask_depth"].cuda(), b["z"].cuda())
        bg_loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
        ref = float(orc.step(b))
        assert abs(float(bg_loss) - ref) < tol_loss * abs(ref)
    for k, p in fc.named_parameters():
        assert rel_l2(p, orc.params[k][0]) < tol_par, k
