"""Swap the reference's quantiser for this one inside an already-built reference model."""
from __future__ import annotations

import types
from typing import Optional

import torch
import torch.nn as nn


def _encode_imgs_vit(self, imgs):
    """models/vitvqgan.py:204-210 with the quantiser's indices-only path: encoder -> pre_quant -> indices (b, n)."""
    enc_imgs = self.pre_quant(self.encoder(imgs))
    return self.codebook.encode(enc_imgs)


def _encode_imgs_vqgan(self, imgs):
    """models/vqgan.py:245-251 with the quantiser's indices-only path; ``rearrange(indices, '(b i) -> b i', b=b)``."""
    b = imgs.shape[0]
    enc_imgs = self.pre_quant(self.encoder(imgs))
    return self.codebook.encode(enc_imgs).view(b, -1)


def _fusable_now(enc_imgs) -> bool:
    # under autocast the reference's pre_quant Linear runs in bf16 and hands the quantiser bf16 rows (SURVEY.md section 8b
    # "modes"); the fused projection computes in fp32, which would be a different (better, but different) result -- so the
    # reference's two calls stay whenever autocast is on or the encoder's output is not fp32
    return enc_imgs.dtype == torch.float32 and not torch.is_autocast_enabled(enc_imgs.device.type)


def _forward_vit_fused(self, imgs):
    """models/vitvqgan.py:190-196 with ``pre_quant`` formed inside the quantiser's token preparation."""
    enc_imgs = self.encoder(imgs)
    if _fusable_now(enc_imgs):
        embeds, indices, loss = self.codebook.forward_projected(enc_imgs, self.pre_quant)
        if not torch.is_grad_enabled():
            # evaluation (vitvqgan.py:194 without a graph): post_quant of the winning codes is a row of the projected table;
            # z_q = zn + (q - zn) equals q to an ulp, so this agrees with post_quant(z_q) to fp32 rounding
            return self.decoder(self.codebook.decode_projected(indices, self.post_quant)), loss
    else:
        embeds, _, loss = self.codebook(self.pre_quant(enc_imgs))
    return self.decoder(self.post_quant(embeds)), loss


def _encode_imgs_vit_fused(self, imgs):
    """models/vitvqgan.py:204-210: encoder -> (pre_quant + quantiser, indices only) -> indices (b, n)."""
    enc_imgs = self.encoder(imgs)
    if _fusable_now(enc_imgs):
        return self.codebook.encode_projected(enc_imgs, self.pre_quant)
    return self.codebook.encode(self.pre_quant(enc_imgs))


def _decode_indices_fused(self, indices):
    """models/vitvqgan.py:198-202 / models/vqgan.py:239-243: lookup + ``post_quant`` as one gather from the projected
    codes when nothing on the way needs a gradient (generation); the reference's composition otherwise."""
    params = [self.codebook.embedding.weight, *self.post_quant.parameters()]
    if (torch.is_grad_enabled() and any(p.requires_grad for p in params)) or torch.is_autocast_enabled(indices.device.type):
        return self.decoder(self.post_quant(self.codebook.indices_to_embeddings(indices)))
    return self.decoder(self.codebook.decode_projected(indices, self.post_quant))


def patch_reference_model(model: nn.Module, form: Optional[str] = None, fast_encode: bool = True,
                          fuse_projections: bool = False) -> nn.Module:
    """Replace ``model.codebook`` (a reference ``Codebook`` from models/vitvqgan.py or models/vqgan.py)
    by the B200 drop-in carrying the same weights, ``beta`` and sizes.  The wrappers' call sites
    (vitvqgan.py:193,200,208; vqgan.py:234,241,249) keep working unchanged.  Returns ``model``.

    ``form``: "vit" or "vqgan"; by default taken from the module the old codebook's class lives in.
    ``fast_encode``: also rebind ``model.encode_imgs`` so that tokenisation (what MaskGIT / Muse / Parti call) takes the
    indices-only path of the quantiser (VQ_FLAG_INDICES_ONLY: no z_q, no loss, nothing saved) instead of running the
    whole forward and dropping two of its three results; same indices, same shape.
    ``fuse_projections``: also rebind ``forward`` / ``encode_imgs`` so that the ViT form's
    ``pre_quant`` Linear is computed inside the quantiser's first kernel (shapes ``supports_fused_pre_quant`` covers), and
    ``decode_indices`` (both forms) so that lookup + ``post_quant`` is one gather from the K projected codes.  Off by
    default: the fused GEMM sums in another order than cuBLAS, so z -- and with it an index on a row whose two best codes
    are closer than fp32 rounding -- can differ from the unfused path by rounding.
    """
    from . import vitvqgan, vqgan

    old = model.codebook
    if form is None:
        ref_module = type(old).__module__
        form = "vqgan" if ref_module.endswith("vqgan") and not ref_module.endswith("vitvqgan") else "vit"
    if form not in ("vit", "vqgan"):
        raise ValueError("form must be 'vit' or 'vqgan'")
    cls = vqgan.Codebook if form == "vqgan" else vitvqgan.Codebook
    new = cls(old.codebook_size, old.codebook_dim, getattr(old, "beta", 0.25))
    new.embedding.load_state_dict(old.embedding.state_dict())
    new.to(old.embedding.weight.device)
    new.embedding.weight.requires_grad_(old.embedding.weight.requires_grad)
    new.train(old.training)
    model.codebook = new
    if fast_encode and all(hasattr(model, a) for a in ("encoder", "pre_quant", "encode_imgs")):
        model.encode_imgs = types.MethodType(_encode_imgs_vqgan if form == "vqgan" else _encode_imgs_vit, model)
    if fuse_projections:
        if all(hasattr(model, a) for a in ("decoder", "post_quant", "decode_indices")):
            model.decode_indices = types.MethodType(_decode_indices_fused, model)
        if (form == "vit" and all(hasattr(model, a) for a in ("encoder", "pre_quant", "post_quant", "decoder"))
                and new.supports_fused_pre_quant(model.pre_quant)):
            model.forward = types.MethodType(_forward_vit_fused, model)
            model.encode_imgs = types.MethodType(_encode_imgs_vit_fused, model)
    return model
