"""CPU restatement of the per-frame ingest the shared keyframe store replaces.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product path never imports this.

Follows, with images stored [W][H] as the reference keeps them (dataset.py:87-91 transposes):
  * ``get_bbox2d_batch``      utils.py:75-84   (first / last+1 index of any() along each axis)
  * ``enlarge_bbox``          utils.py:36-57   (margin = int(0.5*scale*extent), clip to the image)
  * ``replica_frame``         dataset.py:96-131 (per-instance filter by class / size, bbox_dict in
                                                 [u_lo,u_hi,v_lo,v_hi] order, ``inst[obj_ == 0] = 0``)
  * ``state_mask``            train.py:126-128 (this object -> 1, id -1 -> 2, else 0)
Pinned by tests/golden/ingest_*.npz, produced by running the reference's own ``dataset.Replica``
loader on a synthetic Replica-format directory (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def get_bbox2d_batch(masks: np.ndarray):
    """utils.py:75-84 on a [n, d0, d1] boolean stack: (d0_min, d0_max+1, d1_min, d1_max+1) per mask."""
    n, d0, d1 = masks.shape
    rows = masks.any(axis=2)            # [n, d0]
    cols = masks.any(axis=1)            # [n, d1]
    rmins = rows.argmax(axis=1)
    rmaxs = d0 - rows[:, ::-1].argmax(axis=1)
    cmins = cols.argmax(axis=1)
    cmaxs = d1 - cols[:, ::-1].argmax(axis=1)
    return rmins, rmaxs, cmins, cmaxs


def enlarge_bbox(bbox, scale: float, w: int, h: int):
    """utils.py:36-57.  The reference multiplies an int64 torch scalar by a python float, which torch
    evaluates in float32 -- restated explicitly so the truncation lands on the same integer."""
    assert scale >= 0
    min_x, min_y, max_x, max_y = (int(v) for v in bbox)
    half = np.float32(0.5 * scale)
    margin_x = int(np.float32(max_x - min_x) * half)
    margin_y = int(np.float32(max_y - min_y) * half)
    if margin_y == 0 or margin_x == 0:
        return None
    min_x -= margin_x; max_x += margin_x; min_y -= margin_y; max_y += margin_y
    clip = lambda v, hi: int(min(max(v, 0), hi))
    return [clip(min_x, w - 1), clip(min_y, h - 1), clip(max_x, w - 1), clip(max_y, h - 1)]


def replica_frame(inst: np.ndarray, cls: np.ndarray, background_cls, bbox_scale: float = 0.2, min_extent: int = 10):
    """dataset.py:96-131 for one frame.  inst / cls: [W, H] int32.  Returns (bbox_dict, obj) where
    bbox_dict[id] = [u_lo, u_hi, v_lo, v_hi] (int64) and obj is the relabelled instance image."""
    bbox_dict = {}
    obj_ = np.zeros_like(cls)
    inst = inst.copy()
    inst_list, batch_masks = [], []
    for inst_id in np.unique(inst):
        m = inst == inst_id
        sem = np.unique(cls[m])
        assert sem.shape[0] != 0
        if int(sem[0]) in background_cls:           # one class per instance in the Replica renders
            continue
        batch_masks.append(m)
        inst_list.append(int(inst_id))
    if batch_masks:
        masks = np.stack(batch_masks)
        cmins, cmaxs, rmins, rmaxs = get_bbox2d_batch(masks)      # caller-side names of dataset.py:115
        for i in range(masks.shape[0]):
            w = rmaxs[i] - rmins[i]
            h = cmaxs[i] - cmins[i]
            if w <= min_extent or h <= min_extent:
                continue
            be = enlarge_bbox([rmins[i], cmins[i], rmaxs[i], cmaxs[i]], bbox_scale, w=cls.shape[1], h=cls.shape[0])
            obj_[masks[i]] = 1
            bbox_dict[inst_list[i]] = np.array([be[1], be[3], be[0], be[2]], dtype=np.int64)
    inst[obj_ == 0] = 0
    bbox_dict[0] = np.array([0, cls.shape[0], 0, cls.shape[1]], dtype=np.int64)
    return bbox_dict, inst


def state_mask(inst: np.ndarray, obj_id: int) -> np.ndarray:
    """train.py:126-128."""
    s = np.zeros(inst.shape, dtype=np.uint8)
    s[inst == obj_id] = 1
    s[inst == -1] = 2
    return s


def synthetic_instance_frame(W: int, H: int, n_inst: int, seed: int, n_class: int = 100):
    """Random blobs (rectangles / ellipses, some tiny, some overlapping) -> (inst [W,H] i32, cls [W,H] i32).
    Every instance carries one class, as in the Replica renders."""
    rng = np.random.default_rng(seed)
    inst = np.zeros((W, H), dtype=np.int32)
    cls_of = {0: 93}
    uu, vv = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    for k in range(1, n_inst + 1):
        iid = int(rng.integers(1, 4 * n_inst + 1))
        cu, cv = rng.integers(0, W), rng.integers(0, H)
        su, sv = int(rng.integers(2, max(3, W // 3))), int(rng.integers(2, max(3, H // 3)))
        if rng.random() < 0.5:
            m = (np.abs(uu - cu) <= su // 2) & (np.abs(vv - cv) <= sv // 2)
        else:
            m = ((uu - cu) / (su / 2 + 0.5)) ** 2 + ((vv - cv) / (sv / 2 + 0.5)) ** 2 <= 1.0
        inst[m] = iid
        cls_of.setdefault(iid, int(rng.integers(0, n_class)))
    cls = np.zeros((W, H), dtype=np.int32)
    for iid, c in cls_of.items():
        cls[inst == iid] = c
    return inst, cls
