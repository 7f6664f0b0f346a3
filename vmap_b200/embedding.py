"""UniDirsEmbed with the reference's constructor, parameters and state_dict keys
(embedding.py:43-91): ``B_layer.weight`` [21,3] (trainable), buffer ``scale``.

The embedding is never materialised on the GPU path: ``forward`` returns a lazy handle that
``OccupancyMap.forward`` / ``loss.step_batch_loss`` consume by launching the fused kernels.
"""
from __future__ import annotations

import torch

from .lazy import LazyEmbedding

_G, _S, _A, _B = 0.8506508, 0.5257311, 0.809017, 0.309017
ICOSAHEDRON_DIRS = (
    (_G, 0, _S), (_A, .5, _B), (_S, _G, 0), (1, 0, 0), (_A, .5, -_B), (_G, 0, -_S),
    (_B, _A, -.5), (0, _S, -_G), (.5, _B, -_A), (0, 1, 0), (-_S, _G, 0), (-_B, _A, -.5),
    (0, _S, _G), (-_B, _A, .5), (_B, _A, .5), (.5, _B, _A), (.5, -_B, _A), (0, 0, 1),
    (-.5, _B, _A), (-_A, .5, _B), (-_A, .5, -_B),
)


class UniDirsEmbed(torch.nn.Module):
    def __init__(self, min_deg=0, max_deg=2, scale=2.):
        super().__init__()
        if min_deg != 0:
            raise ValueError("the fused kernels assume min_deg == 0 (as every reference config does)")
        self.min_deg, self.max_deg = min_deg, max_deg
        self.n_freqs = max_deg - min_deg + 1
        self.tensor_scale = torch.tensor(scale, requires_grad=False)
        self.B_layer = torch.nn.Linear(3, 21, bias=False)
        self.B_layer.weight.data = torch.tensor(ICOSAHEDRON_DIRS, dtype=torch.float32)
        bands = 2.0 ** torch.linspace(self.min_deg, self.max_deg, self.n_freqs)
        self.register_buffer("frequency_bands", bands, persistent=False)
        self.register_buffer("scale", self.tensor_scale, persistent=True)

    @property
    def embedding_size(self):
        return 3 + 21 * self.n_freqs

    def forward(self, x):
        return LazyEmbedding(x, pe=self)
