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
