"""Drop-in for ``models/vqgan.py:138-182`` (CNN-VQGAN ``Codebook``)."""
from __future__ import annotations

import torch

from ._module import _CodebookBase


class Codebook(_CodebookBase):
    """``Codebook(codebook_size=1024, codebook_dim=256, beta=0.25)`` -- reference vqgan.py:139-146.

    forward(z: (b, D, h, w)) -> (z_q (b, D, h, w) fp32, indices (b*h*w,) int64 in (b, h, w) order, loss)
    with loss = mean((sg q - zn)^2) + beta * mean((q - sg zn)^2)   (vqgan.py:169).
    indices_to_embeddings does NOT renormalise (vqgan.py:178-182) -- kept bug-compatible.
    """

    form = "vqgan"

    def __init__(self, codebook_size: int = 1024, codebook_dim: int = 256, beta: float = 0.25):
        super().__init__(codebook_size, codebook_dim, beta)
        self.embedding.weight.data.uniform_(-1.0 / self.codebook_size, 1.0 / self.codebook_size)   # vqgan.py:146

    def forward(self, z: torch.Tensor):
        return self._quantise(z)
