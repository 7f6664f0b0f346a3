"""OccupancyMap with the reference's constructor and state_dict keys (model.py:16-85), so the
per-object checkpoints of vmap.py:461-491 interoperate.  The arithmetic runs in the fused
CUDA kernels; ``forward`` consumes the lazy embedding handle and returns lazy heads that
``loss.step_batch_loss`` (training) or ``.materialize()`` (inference) resolve."""
from __future__ import annotations

import torch

from .lazy import LazyEmbedding, LazyHeads


def init_weights(m, init_fn=torch.nn.init.xavier_normal_):
    if type(m) == torch.nn.Linear:       # model.py:4-6
        init_fn(m.weight)


def fc_block(in_f, out_f):
    return torch.nn.Sequential(torch.nn.Linear(in_f, out_f), torch.nn.ReLU(inplace=True))


class OccupancyMap(torch.nn.Module):
    def __init__(self, emb_size1, emb_size2, hidden_size=256, do_color=True, hidden_layers_block=1):
        super().__init__()
        if not do_color or hidden_layers_block != 1:
            raise ValueError("the fused kernels implement do_color=True, hidden_layers_block=1 "
                             "(the only configuration the reference constructs, trainer.py:27-31)")
        self.do_color = do_color
        self.embedding_size1, self.embedding_size2 = emb_size1, emb_size2
        self.hidden_size = hidden_size
        self.in_layer = fc_block(emb_size1, hidden_size)
        self.mid1 = torch.nn.Sequential(fc_block(hidden_size, hidden_size))
        self.cat_layer = fc_block(hidden_size + emb_size1, hidden_size)
        self.mid2 = torch.nn.Sequential(fc_block(hidden_size, hidden_size))
        self.out_alpha = torch.nn.Linear(hidden_size, 1)
        self.color_linear = fc_block(emb_size2 + hidden_size, hidden_size)
        self.out_color = torch.nn.Linear(hidden_size, 3)

    def forward(self, x, noise_std=None, do_alpha=True, do_color=True, do_cat=True):
        if noise_std is not None or not (do_alpha and do_color and do_cat):
            raise NotImplementedError("only the default forward (model.py:54-85 with all branches on) is fused")
        if not isinstance(x, LazyEmbedding):
            raise TypeError("OccupancyMap.forward expects the handle returned by UniDirsEmbed.forward; "
                            "there is no eager PyTorch path (the step runs in the fused CUDA kernels)")
        heads = LazyHeads(x, fc=self)
        return heads.alpha, heads.color
