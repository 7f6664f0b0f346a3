"""K4 frame ingest and the shared keyframe store on the GPU: integer results are exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io
from oracle import sampler_oracle as so
from tests._util import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _store(W, H, cap=8, max_id=512):
    from vmap_b200.keyframes import FrameStore
    return FrameStore(W, H, cap, DEV, max_id=max_id)


@pytest.mark.parametrize("name", ["ingest_small", "ingest_replica_size"])
def test_ingest_matches_reference_loader_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    inst, cls = g["inst"].astype(np.int32), g["cls"].astype(np.int32)
    W, H = inst.shape
    st = _store(W, H)
    rng = np.random.default_rng(0)
    rgb = torch.from_numpy(rng.integers(0, 255, (W, H, 3), dtype=np.uint8))
    depth = torch.from_numpy(rng.random((W, H), dtype=np.float32))
    slot, stats, bbox = st.ingest(rgb, depth, torch.from_numpy(inst), torch.eye(4), frame_id=0,
                                  cls=torch.from_numpy(cls), background_cls=g["background_cls"].tolist(),
                                  bbox_scale=float(g["bbox_scale"]))
    vis = st.visible_objects()
    assert sorted(vis) == g["ids"].tolist()
    for i, b in zip(g["ids"].tolist(), g["bboxes"]):
        assert vis[i].cpu().tolist() == [float(x) for x in b], i
    assert np.array_equal(st.inst[slot].cpu().numpy(), g["obj"].astype(np.int32))
    assert torch.equal(st.rgbx[slot, :, :, :3].cpu(), rgb) and torch.equal(st.depth[slot].cpu(), depth)
    cnt = np.bincount(inst.ravel(), minlength=st.max_id)
    assert np.array_equal(stats[:, 0].cpu().numpy(), cnt[:st.max_id])


def test_ingest_many_ids_unknown_and_out_of_range():
    """> 128 distinct ids in one CTA's pixels (block-table overflow path), id -1 kept, ids >= max_id ignored."""
    W, H, max_id = 257, 131, 300
    rng = np.random.default_rng(5)
    inst = rng.integers(-1, 340, (W, H)).astype(np.int32)             # salt-and-pepper: every warp sees many ids
    inst[40:120, 30:90] = 7                                            # one real object
    st = _store(W, H, max_id=max_id)
    slot, stats, bbox = st.ingest(torch.zeros(W, H, 3, dtype=torch.uint8), torch.zeros(W, H), torch.from_numpy(inst),
                                  torch.eye(4), bbox_scale=0.2)
    s = stats.cpu().numpy()
    for i in range(max_id):
        m = inst == i
        assert s[i, 0] == m.sum(), i
        if m.any():
            uu, vv = np.nonzero(m)
            assert s[i, 1:5].tolist() == [uu.min(), uu.max() + 1, vv.min(), vv.max() + 1], i
    out = st.inst[slot].cpu().numpy()
    keep = s[:, 7].astype(bool)
    exp = np.where(inst < 0, inst, np.where((inst < max_id) & keep[np.clip(inst, 0, max_id - 1)], inst, 0))
    assert np.array_equal(out, exp) and (out == -1).sum() == (inst == -1).sum()


def test_enlarge_arithmetic_matches_reference_table():
    g = np.load(os.path.join(GOLDEN, "ingest_enlarge_table.npz"))
    W, H = 1400, 64
    st = _store(W, H, cap=1, max_id=8)
    ext = [11, 12, 13, 14, 15, 19, 20, 21, 24, 25, 26, 29, 30, 31, 33, 37, 99, 100, 101, 333, 679, 680, 999, 1199, 1200, 1300]
    for si, sc in enumerate(g["scales"].tolist()):
        for e in ext:
            inst = np.zeros((W, H), dtype=np.int32)
            inst[40:40 + e, 10:40] = 3
            _, stats, bbox = _ingest_nostore(st, inst, sc)
            m = int(g["margins"][si][e - 1])
            b = bbox[3].cpu().tolist()
            if m == 0:
                assert int(stats[3, 7]) == 0
            else:
                assert b[0] == max(40 - m, 0) and b[1] == min(40 + e + m, W - 1), (sc, e, b, m)


def _ingest_nostore(st, inst, scale):
    import ctypes as C
    from vmap_b200 import _lib
    a = _lib.IngestArgs()
    t = torch.from_numpy(inst).to(DEV)
    a.width, a.height, a.inst, a.max_id = st.W, st.H, C.c_void_p(t.data_ptr()), st.max_id
    a.bbox_scale, a.min_extent = float(scale), 10
    a.stats, a.bbox = C.c_void_p(st.stats.data_ptr()), C.c_void_p(st.bbox.data_ptr())
    _lib.check(st._handle, st.lib.vmb_ingest_frame(st._handle, C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "vmb_ingest_frame")
    torch.cuda.synchronize()
    return -1, st.stats, st.bbox


def test_shared_store_sampler_is_bit_identical_to_per_object_copies():
    """Sampling through (store slot, bbox) tables + instance ids == sampling the reference's per-object
    rgbs_batch/depth_batch copies with the state channel of train.py:126-128, same seed -> same bits."""
    from vmap_b200.sampler import BatchedSampler, KeyframeSet, KeyframeTables
    W, H, KF, n_frames_store = 96, 72, 6, 9
    rng = np.random.default_rng(3)
    st = _store(W, H, cap=12)
    frames = []
    for f in range(n_frames_store):
        inst, _ = io.synthetic_instance_frame(W, H, 10, seed=100 + f)
        inst[rng.random((W, H)) < 0.05] = -1
        rgb = torch.from_numpy(rng.integers(0, 255, (W, H, 3), dtype=np.uint8))
        depth = torch.from_numpy((rng.random((W, H), dtype=np.float32) * 4).astype(np.float32))
        depth[torch.from_numpy(rng.random((W, H)) < 0.1)] = 0.0
        T = torch.eye(4); T[:3, 3] = torch.from_numpy(rng.random(3).astype(np.float32))
        slot = st.put(rgb, depth, torch.from_numpy(inst), T, frame_id=f)
        frames.append((slot, rgb, depth, inst, T))
    obj_ids = [0, 3, 11, 17]
    B = len(obj_ids)
    n_kf = [6, 4, 3, 5]
    kf_slot = np.zeros((B, KF), dtype=np.int32)
    kf_bbox = np.zeros((B, KF, 4), dtype=np.float32)
    sets = []
    for b, oid in enumerate(obj_ids):
        pick = rng.choice(n_frames_store, n_kf[b], replace=False)
        rgbs = torch.zeros(KF, W, H, 4, dtype=torch.uint8); deps = torch.zeros(KF, W, H); twc = torch.zeros(KF, 4, 4)
        for k, f in enumerate(pick):
            slot, rgb, depth, inst, T = frames[f]
            kf_slot[b, k] = slot
            u0, v0 = rng.integers(0, W // 2), rng.integers(0, H // 2)
            kf_bbox[b, k] = [u0, u0 + rng.integers(8, W // 2), v0, v0 + rng.integers(8, H // 2)]
            rgbs[k, :, :, :3] = rgb; rgbs[k, :, :, 3] = torch.from_numpy(io.state_mask(inst, oid))
            deps[k] = depth; twc[k] = T
        sets.append(KeyframeSet(rgbs.to(DEV), deps.to(DEV), twc.to(DEV), torch.from_numpy(kf_bbox[b]).to(DEV),
                                n_kf[b], [n_kf[b] - 2, n_kf[b] - 1]))
    tables = KeyframeTables(kf_slot, kf_bbox, obj_ids, n_kf, [[n - 2, n - 1] for n in n_kf])
    smp = BatchedSampler(DEV, n_bins_cam2surface=1, n_bins=9)
    rays = so.camera_ray_dirs(W, H, 60.0, 60.0, W / 2 - 0.5, H / 2 - 0.5).to(DEV).contiguous()
    a = smp.sample(sets, 10, 16, rays, seed=77, offset=3, want_u8=True)
    a = {k: v.clone() for k, v in a.items()}
    s = smp.sample_store(st, tables, 10, 16, rays, seed=77, offset=3, want_u8=True)
    for k in a:
        assert torch.equal(a[k], s[k]), k
    assert set(torch.unique(s["sem"]).tolist()) == {0, 1, 2}


def test_frame_store_refcounts_and_capacity():
    from vmap_b200 import _lib
    st = _store(16, 12, cap=3)
    z = lambda: (torch.zeros(16, 12, 3, dtype=torch.uint8), torch.zeros(16, 12), torch.zeros(16, 12, dtype=torch.int32), torch.eye(4))
    s0, s1, s2 = st.put(*z()), st.put(*z()), st.put(*z())
    assert st.n_used == 3
    with pytest.raises(_lib.VmbError):
        st.put(*z())
    st.acquire(s1); st.release(s1); assert st.n_used == 3
    st.release(s1); assert st.n_used == 2
    assert st.put(*z()) == s1


def test_frame_loop_with_shared_store_trains():
    """train.py:104-141,195-326 with the shared store: ingest -> objects from the kept instances ->
    sample_all (one launch, store mode) -> fused step.  Loss finite and decreasing, references consistent."""
    import types
    from vmap_b200 import vmap as vm
    from vmap_b200 import synth
    from vmap_b200.ensemble import VmapEnsemble
    W, H = 160, 120
    cfg = types.SimpleNamespace(do_bg=False, data_device=DEV, training_device=DEV, obj_scale=2.0, bg_scale=5.0,
                                hidden_feature_size=32, hidden_feature_size_bg=128, n_bins_cam2surface=1,
                                n_bins_cam2surface_bg=5, keyframe_step=2, keyframe_step_bg=5, min_depth=0.0,
                                max_depth=8.0, n_bins=9, n_unidir_funcs=5, surface_eps=0.1, stop_eps=0.05,
                                keyframe_buffer_size=5, W=W, H=H, fx=100.0, fy=100.0, cx=W / 2 - 0.5, cy=H / 2 - 0.5,
                                n_unidir=5)
    real_trainer = vm.trainer_mod.Trainer
    vm.trainer_mod.Trainer = lambda c: types.SimpleNamespace()
    try:
        st = _store(W, H, cap=16)
        cam = vm.cameraInfo(cfg)
        inst, cls = io.synthetic_instance_frame(W, H, 14, seed=21)
        objs = {}
        rng = np.random.default_rng(1)
        for fid in range(9):
            rgb = torch.from_numpy(rng.integers(0, 255, (W, H, 3), dtype=np.uint8))
            depth = torch.from_numpy((rng.random((W, H), dtype=np.float32) * 2 + 1).astype(np.float32))
            T = torch.eye(4); T[0, 3] = 0.01 * fid
            slot, _, _ = st.ingest(rgb, depth, torch.from_numpy(inst), T, frame_id=fid, cls=torch.from_numpy(cls),
                                   background_cls=[5, 12, 30, 31, 40, 60, 92, 93, 95, 97, 98, 79])
            for oid, bbox in st.visible_objects().items():
                if oid == 0:
                    continue
                bbox = bbox.cpu()
                if oid in objs:
                    objs[oid].append_keyframe(None, None, None, bbox, T, fid, frame_slot=slot)
                else:
                    objs[oid] = vm.sceneObject(cfg, oid, None, None, None, bbox, T, fid, store=st, frame_slot=slot)
            st.release(slot)
        olist = list(objs.values())
        assert len(olist) >= 3 and st.n_used <= 9       # objects prune independently: <= frames seen
        refs = sum(sum(o._held) for o in olist)
        assert sum(st.refcount) == refs
        ens = VmapEnsemble(len(olist), hidden=32, scale=2.0, device=DEV)
        ens.load_stacked(synth.init_params(len(olist), 32, seed=0))
        losses = []
        for it in range(30):
            batch = vm.sample_all(olist, 10, 12, cam.rays_dir_cache.contiguous(), seed=it)
            assert set(torch.unique(batch["sem"]).tolist()) <= {0, 1, 2} and bool((batch["sem"] == 1).any())
            losses.append(float(ens.step(batch)))
        ens.check_status()
        assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
        for o in olist:
            o.release_frames()
        assert st.n_used == 0
    finally:
        vm.trainer_mod.Trainer = real_trainer
