"""Drop-in for ``models/vitvqgan.py:140-176`` (ViT-VQGAN ``Codebook``)."""
from __future__ import annotations

import torch

from ._module import _CodebookBase


class Codebook(_CodebookBase):
    """``Codebook(codebook_size=8192, codebook_dim=32, beta=0.25)`` -- reference vitvqgan.py:141-149.

    forward(z: (b, n, D)) -> (z_q (b, n, D) fp32, indices (b, n) int64, loss 0-dim) with
    loss = beta * mean((sg q - zn)^2) + mean((q - sg zn)^2)   (vitvqgan.py:166)
    """

    form = "vit"

    def __init__(self, codebook_size: int = 8192, codebook_dim: int = 32, beta: float = 0.25):
        super().__init__(codebook_size, codebook_dim, beta)
        self.embedding.weight.data.normal_()          # vitvqgan.py:149

    def forward(self, z: torch.Tensor):
        z_q, flat_idx, loss = self._quantise(z)
        return z_q, flat_idx.view(*z.shape[:-1]), loss

    def encode(self, z: torch.Tensor, index_dtype: torch.dtype = torch.int64) -> torch.Tensor:
        return super().encode(z, index_dtype).view(*z.shape[:-1])
