"""CPU-only checks of the host-side mirror of the reference's call surface."""
import json
import os

import numpy as np
import pytest
import torch

from tests._util import GOLDEN
from vmap_b200 import cfg as cfg_mod
from vmap_b200.embedding import UniDirsEmbed
from vmap_b200.layout import ALL_KEYS, FC_KEYS
from vmap_b200.model import OccupancyMap


@pytest.mark.parametrize("tag", ["vMAP", "iMAP"])
def test_config_matches_reference_attribute_bag(tag):
    g = json.load(open(os.path.join(GOLDEN, "config_room0.json")))[tag]
    c = cfg_mod.Config(config_dict=g["raw"])
    mine = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in vars(c).items()}
    assert mine == g["attrs"]


def test_replica_dict_equals_shipped_json():
    g = json.load(open(os.path.join(GOLDEN, "config_room0.json")))
    for tag, imap in (("vMAP", False), ("iMAP", True)):
        d = cfg_mod.replica_room0_dict(imap=imap)
        raw = g[tag]["raw"]
        raw["dataset"]["path"] = ""
        assert d == raw


def test_module_state_dict_keys_match_reference():
    g = np.load(os.path.join(GOLDEN, "step_vmap_h32.npz"))
    m = OccupancyMap(87, 42, hidden_size=32)
    assert tuple(k for k, _ in m.named_parameters()) == FC_KEYS
    for k, p in m.named_parameters():
        assert tuple(p.shape) == g["p0_" + k].shape[1:], k
    pe = UniDirsEmbed(max_deg=5, scale=2.0)
    assert set(pe.state_dict()) == {"scale", "B_layer.weight"}
    np.testing.assert_allclose(pe.B_layer.weight.detach().numpy(), np.load(
        os.path.join(GOLDEN, "step_bg_h128.npz"))["p0_B_layer.weight"][0], atol=0)
    assert pe.embedding_size == 129 and ALL_KEYS[-1] == "B_layer.weight"


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA instead of computing on the host."""
    from vmap_b200 import loss as loss_mod
    from vmap_b200.utils import vmap
    pe = UniDirsEmbed(max_deg=5)
    fc = OccupancyMap(87, 42, hidden_size=32)
    alpha, col = fc(pe(torch.zeros(4, 10, 3)))
    with pytest.raises(Exception):
        loss_mod.step_batch_loss(alpha, col, torch.zeros(4), torch.zeros(4, 3), torch.zeros(4, dtype=torch.uint8),
                                 torch.ones(4, dtype=torch.bool), torch.zeros(4, 10))
    with pytest.raises(TypeError):
        vmap(fc)
    with pytest.raises(TypeError):
        fc(torch.zeros(4, 129))
    if not torch.cuda.is_available():
        from vmap_b200.ensemble import VmapEnsemble
        with pytest.raises(Exception):
            VmapEnsemble(2, device="cpu")


def test_sampler_tables_packing_round_trip():
    """SamplerTables packs the per-object tables of a sampler launch into one (pinned) buffer; the views handed to
    the C ABI must read back exactly what was filled in, in both keyframe layouts (host logic, no GPU)."""
    import ctypes as C
    import numpy as np
    from vmap_b200 import _lib
    from vmap_b200.sampler import KeyframeSet, KeyframeTables, SamplerTables
    B, KF = 3, 5
    # store mode
    rng = np.random.default_rng(0)
    kt = KeyframeTables(rng.integers(0, 9, (B, KF)).astype(np.int32), rng.random((B, KF, 4)).astype(np.float32),
                        [4, 7, 9], [5, 3, 2], [[3, 4], [1, 2], [0, 1]])
    t = SamplerTables("cpu", B, kf_stride=KF)
    t.fill_store(kt)
    t.upload()
    a = _lib.SampleArgs()
    t.bind(a)
    rd = lambda ptr, n, ct: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()
    assert rd(a.kf_slot, B * KF, C.c_int).tolist() == kt.kf_slot.reshape(-1).tolist()
    assert np.array_equal(rd(a.kf_bbox, B * KF * 4, C.c_float), kt.kf_bbox.reshape(-1).numpy())
    assert rd(a.obj_id, B, C.c_int).tolist() == [4, 7, 9] and rd(a.n_keyframes, B, C.c_int).tolist() == [5, 3, 2]
    assert rd(a.latest_kf, 2 * B, C.c_int).tolist() == [3, 4, 1, 2, 0, 1] and a.kf_stride == KF
    # per-object mode
    sets = [KeyframeSet(torch.zeros(KF, 4, 3, 4, dtype=torch.uint8), torch.zeros(KF, 4, 3), torch.zeros(KF, 4, 4),
                        torch.zeros(KF, 4), n, [n - 2, n - 1]) for n in (5, 4, 3)]
    t2 = SamplerTables("cpu", B)
    t2.fill_objects(sets)
    t2.upload()
    a2 = _lib.SampleArgs()
    t2.bind(a2)
    assert rd(a2.rgbs, B, C.c_longlong).tolist() == [s.rgbs_batch.data_ptr() for s in sets]
    assert rd(a2.bbox, B, C.c_longlong).tolist() == [s.bbox.data_ptr() for s in sets]
    assert rd(a2.n_keyframes, B, C.c_int).tolist() == [5, 4, 3]
    assert rd(a2.latest_kf, 2 * B, C.c_int).tolist() == [3, 4, 2, 3, 1, 2]


def test_synth_matches_oracle_generators():
    """`vmap_b200/synth.py` exists so that bench.py's product arm never imports `oracle/`; it must stay the same
    generator as the oracle's (same seeds -> same tensors), or the two bench arms would run different workloads."""
    import torch
    from oracle import vmap_oracle as vo
    from vmap_b200 import synth
    for kw in (dict(B=3, R=17, S=10, seed=5), dict(B=1, R=9, S=14, seed=6, n_cam2surf=5)):
        a = vo.synthetic_batch(kw["B"], kw["R"], kw["S"], seed=kw["seed"], **({"n_cam2surf": kw["n_cam2surf"]} if "n_cam2surf" in kw else {}))
        b = synth.synthetic_batch(kw["B"], kw["R"], kw["S"], seed=kw["seed"], **({"n_cam2surf": kw["n_cam2surf"]} if "n_cam2surf" in kw else {}))
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(torch.as_tensor(a[k]), torch.as_tensor(b[k])), k
    for H in (32, 128):
        pa, pb = vo.init_params(2, H, seed=11), synth.init_params(2, H, seed=11)
        assert pa.keys() == pb.keys()
        for k in pa:
            assert torch.equal(torch.as_tensor(pa[k]), torch.as_tensor(pb[k])), k


def test_rows_dense_equals_per_object_contiguity():
    """`ensemble.rows_dense(t)` must say exactly what `t[0].is_contiguous()` says (it replaces that check on the
    per-step path without building six views per call)."""
    import itertools
    import torch
    from vmap_b200.ensemble import rows_dense
    base = torch.zeros(4, 12, 5, 3)
    views = [base, base[:, 2:7], base[:, ::2], base[:, :, 1:3], base[..., 0], base[:, :, :, :2], base[1:2], base[:, 3:4],
             base.permute(0, 2, 1, 3), base[:, :, 2], torch.zeros(3, 7)[:, 1:5], torch.zeros(3, 7).t(), torch.zeros(2, 1, 6)[:, :, ::2],
             torch.zeros(5, 8, dtype=torch.uint8)[:, 2:6], base.expand(4, 12, 5, 3), torch.zeros(1, 9, 3).expand(4, 9, 3)]
    for v in views:
        assert rows_dense(v) == v[0].is_contiguous(), (tuple(v.shape), v.stride())
