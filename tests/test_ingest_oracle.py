"""Frame-ingest oracle (oracle/ingest_oracle.py) against goldens produced by the reference's own data loader
(dataset.Replica.__getitem__ run on a synthetic Replica-format directory, oracle/make_golden.py), plus the
host bookkeeping of the shared keyframe store (no GPU needed)."""
import os
import random
import types

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io
from tests._util import GOLDEN


@pytest.mark.parametrize("name", ["ingest_small", "ingest_replica_size"])
def test_replica_frame_matches_reference_loader(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    inst, cls = g["inst"].astype(np.int32), g["cls"].astype(np.int32)
    bbox_dict, obj = io.replica_frame(inst, cls, set(g["background_cls"].tolist()), float(g["bbox_scale"]))
    assert sorted(bbox_dict) == g["ids"].tolist()
    for i, b in zip(g["ids"].tolist(), g["bboxes"]):
        assert bbox_dict[i].tolist() == b.tolist(), i
    assert np.array_equal(obj, g["obj"].astype(np.int32))


def test_enlarge_margin_is_float32_truncation():
    g = np.load(os.path.join(GOLDEN, "ingest_enlarge_table.npz"))
    for si, sc in enumerate(g["scales"].tolist()):
        for e, m in zip(g["extents"].tolist(), g["margins"][si].tolist()):
            r = io.enlarge_bbox([2000, 2000, 2000 + e, 2000 + e], sc, 10000, 10000)
            assert (0 if r is None else 2000 - r[0]) == m, (sc, e)


def test_state_mask():
    inst = np.array([[3, -1, 0], [3, 7, -1]], dtype=np.int32)
    assert io.state_mask(inst, 3).tolist() == [[1, 2, 0], [1, 0, 2]]


class _FakeStore:
    """Reference-counting half of keyframes.FrameStore (the GPU half is covered by the -m gpu tests)."""
    W, H = 8, 6

    def __init__(self, cap):
        self.refcount = [0] * cap
        self.next = 0

    def take(self):
        s = self.refcount.index(0)
        self.refcount[s] = 1
        return s

    def acquire(self, s):
        assert self.refcount[s] > 0
        self.refcount[s] += 1

    def release(self, s):
        assert self.refcount[s] > 0
        self.refcount[s] -= 1


def test_shared_store_objects_follow_the_reference_keyframe_policy(monkeypatch):
    """Two objects fed the same frames: one in the reference's per-object layout (CPU tensors), one in
    shared-store mode.  Slot tables, keyframe counts and latest-queue must evolve identically
    (vmap.py:208-268), and the store's reference counts must equal the table contents."""
    from vmap_b200 import vmap as vm
    monkeypatch.setattr(vm.trainer_mod, "Trainer", lambda cfg: types.SimpleNamespace())
    cfg = types.SimpleNamespace(do_bg=False, data_device="cpu", training_device="cpu", obj_scale=2.0, bg_scale=5.0,
                                hidden_feature_size=32, hidden_feature_size_bg=128, n_bins_cam2surface=1,
                                n_bins_cam2surface_bg=5, keyframe_step=3, keyframe_step_bg=5, min_depth=0.0,
                                max_depth=8.0, n_bins=9, n_unidir_funcs=5, surface_eps=0.1, stop_eps=0.05,
                                keyframe_buffer_size=6)
    W, H = _FakeStore.W, _FakeStore.H
    store = _FakeStore(64)
    frames = {}

    def new_frame(fid):
        g = torch.Generator().manual_seed(fid)
        rgb = torch.randint(0, 255, (W, H, 3), dtype=torch.uint8, generator=g)
        depth = torch.rand(W, H, generator=g)
        mask = torch.randint(0, 3, (W, H), dtype=torch.uint8, generator=g)
        bbox = torch.tensor([0., fid % W, 0., fid % H])
        return rgb, depth, mask, bbox, torch.eye(4)

    rgb, depth, mask, bbox, T = new_frame(0)
    slot = store.take(); frames[slot] = 0
    a = vm.sceneObject(cfg, 5, rgb, depth, mask, bbox, T, 0)
    b = vm.sceneObject(cfg, 5, None, None, None, bbox, T, 0, store=store, frame_slot=slot)
    store.release(slot)
    for fid in range(1, 40):
        rgb, depth, mask, bbox, T = new_frame(fid)
        slot = store.take(); frames[slot] = fid
        random.seed(fid); a.append_keyframe(rgb, depth, mask, bbox, T, fid)
        random.seed(fid); b.append_keyframe(None, None, None, bbox, T, fid, frame_slot=slot)
        store.release(slot)
        assert a.n_keyframes == b.n_keyframes and a.kf_pointer == b.kf_pointer
        assert a.lastest_kf_queue == b.lastest_kf_queue and a.kf_id_dict == b.kf_id_dict
        assert torch.equal(a.bbox[:a.n_keyframes].cpu(), b.bbox[:b.n_keyframes])
        # the frame an object's slot k points at is the frame the reference layout holds in slot k
        for fr, k in b.kf_id_dict.items():
            assert frames[b.kf_store_slot[k]] == fr
        held = sorted(b.kf_store_slot[k] for k in range(cfg.keyframe_buffer_size) if b._held[k])
        assert sorted(s for s, c in enumerate(store.refcount) for _ in range(c)) == held
    b.release_frames()
    assert sum(store.refcount) == 0
