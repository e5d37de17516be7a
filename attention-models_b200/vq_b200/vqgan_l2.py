"""The CNN-VQGAN ``Codebook`` WITHOUT l2 normalisation: plain squared-L2 nearest code on the raw vectors.

Not a drop-in for a class of the reference (both reference Codebooks normalise, /root/reference/models/vqgan.py:154-155,163);
it is the reference's CNN form with ``l2_norm`` replaced by the identity, which BASELINE.json's north_star names as "the L2
form".  Same surface as ``vq_b200.vqgan.Codebook``; the search is the exhaustive fp32 scan (the tensor-core filters' error
bounds assume unit rows)."""
from __future__ import annotations

import torch

from ._module import _CodebookBase


class Codebook(_CodebookBase):
    """forward(z: (b, D, h, w)) -> (z_q (b, D, h, w), indices (b*h*w,) int64, loss); d = (|z|^2 + |e|^2) - 2 z.e on raw vectors,
    loss = mean((sg q - z)^2) + beta * mean((q - sg z)^2), z_q = z + sg(q - z), q = E[idx]."""

    form = "l2"

    def __init__(self, codebook_size: int = 1024, codebook_dim: int = 256, beta: float = 0.25):
        super().__init__(codebook_size, codebook_dim, beta)
        self.embedding.weight.data.uniform_(-1.0 / self.codebook_size, 1.0 / self.codebook_size)

    def forward(self, z: torch.Tensor):
        return self._quantise(z)
