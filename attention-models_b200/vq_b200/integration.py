"""Swap the reference's quantiser for this one inside an already-built reference model."""
from __future__ import annotations

import torch.nn as nn


def patch_reference_model(model: nn.Module) -> nn.Module:
    """Replace ``model.codebook`` (a reference ``Codebook`` from models/vitvqgan.py or models/vqgan.py)
    by the B200 drop-in carrying the same weights, ``beta`` and sizes.  The wrappers' call sites
    (vitvqgan.py:193,200,208; vqgan.py:234,241,249) keep working unchanged.  Returns ``model``.
    """
    from . import vitvqgan, vqgan

    old = model.codebook
    ref_module = type(old).__module__
    cls = vqgan.Codebook if ref_module.endswith("vqgan") and not ref_module.endswith("vitvqgan") else vitvqgan.Codebook
    new = cls(old.codebook_size, old.codebook_dim, getattr(old, "beta", 0.25))
    new.embedding.load_state_dict(old.embedding.state_dict())
    new.to(old.embedding.weight.device)
    new.embedding.weight.requires_grad_(old.embedding.weight.requires_grad)
    new.train(old.training)
    model.codebook = new
    return model
