"""Names, shapes and packing of one object's trainable tensors.

Order and names are those of the reference's ``OccupancyMap.named_parameters()``
(model.py:17-52) followed by ``UniDirsEmbed.B_layer.weight`` (embedding.py:75-76),
i.e. exactly what ``utils.update_vmap`` stacks (utils.py:30-34) and what the
per-object checkpoints store (vmap.py:461-476).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

N_DIRS = 21
EMB_SIZE1 = N_DIRS * (3 + 1) + 3          # trainer.py:16

FC_KEYS: Tuple[str, ...] = (
    "in_layer.0.weight", "in_layer.0.bias",
    "mid1.0.0.weight", "mid1.0.0.bias",
    "cat_layer.0.weight", "cat_layer.0.bias",
    "mid2.0.0.weight", "mid2.0.0.bias",
    "out_alpha.weight", "out_alpha.bias",
    "color_linear.0.weight", "color_linear.0.bias",
    "out_color.weight", "out_color.bias",
)
PE_KEY = "B_layer.weight"
ALL_KEYS: Tuple[str, ...] = FC_KEYS + (PE_KEY,)


def emb_size2(n_unidir_funcs: int) -> int:
    return N_DIRS * (n_unidir_funcs + 1) + 3 - EMB_SIZE1   # trainer.py:17


def tensor_shapes(hidden: int, n_unidir_funcs: int = 5) -> Dict[str, Tuple[int, ...]]:
    h, e1, e2 = hidden, EMB_SIZE1, emb_size2(n_unidir_funcs)
    return {
        "in_layer.0.weight": (h, e1), "in_layer.0.bias": (h,),
        "mid1.0.0.weight": (h, h), "mid1.0.0.bias": (h,),
        "cat_layer.0.weight": (h, h + e1), "cat_layer.0.bias": (h,),
        "mid2.0.0.weight": (h, h), "mid2.0.0.bias": (h,),
        "out_alpha.weight": (1, h), "out_alpha.bias": (1,),
        "color_linear.0.weight": (h, h + e2), "color_linear.0.bias": (h,),
        "out_color.weight": (3, h), "out_color.bias": (3,),
        PE_KEY: (N_DIRS, 3),
    }


def host_offsets(hidden: int, n_unidir_funcs: int = 5) -> Tuple[int, int, List[int], List[int]]:
    """(count, stride, offsets, sizes) computed on the host; must agree with
    ``vmb_param_offsets`` (checked in tests/test_boundary.py)."""
    offs, sizes, o = [], [], 0
    for k in ALL_KEYS:
        n = 1
        for d in tensor_shapes(hidden, n_unidir_funcs)[k]:
            n *= d
        offs.append(o)
        sizes.append(n)
        o += n
    return o, (o + 31) // 32 * 32, offs, sizes
