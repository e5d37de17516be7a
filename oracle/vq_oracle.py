"""Restatement of the reference VQ codebook quantiser (test infrastructure only).

Follows, function by function, the two ``Codebook`` modules of
pranoyr/attention-models:

* ViT-VQGAN form  -- ``/root/reference/models/vitvqgan.py:140-176``
* CNN-VQGAN form  -- ``/root/reference/models/vqgan.py:138-182``

The arithmetic itself lives in PyTorch ATen (third-party, unpinned by the
reference; torch 2.11.0+cu128 in this image), reached through the same call
sites the reference uses (``F.normalize``, ``torch.sum``, ``torch.einsum``,
``torch.argmin``, embedding gather, ``torch.mean``).  The functions below are
device-agnostic: on CPU they are the CPU baseline / golden-vector checker, on a
CUDA device they reproduce what the reference itself would compute on that GPU
(used by the ``-m gpu`` parity tests; nothing here is shipped).

Parity pinning: see ``oracle/__init__.py`` and ``oracle/make_golden.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

VIT = "vit"      # reference models/vitvqgan.py:140-176
VQGAN = "vqgan"  # reference models/vqgan.py:138-182
L2 = "l2"        # NOT a form of the reference: the CNN form with the two l2_norm calls (models/vqgan.py:154-155,163) removed --
                 # plain squared-L2 nearest code on raw vectors, which BASELINE.json's north_star names.  PARITY UNPINNED for
                 # this form: there is no reference code to mint fixtures from; it is the reference's own expression
                 # (models/vqgan.py:157-176) with `l2_norm` replaced by the identity.
NORM_EPS = 1e-12  # F.normalize default eps (torch/nn/functional.py:5707-5708)


def unit_rows(x: torch.Tensor) -> torch.Tensor:
    """``l2_norm`` -- reference models/vitvqgan.py:16-17, models/vqgan.py:7-8."""
    return F.normalize(x, p=2, dim=-1)


def distance_matrix(zn_flat: torch.Tensor, en: torch.Tensor) -> torch.Tensor:
    """The T x K matrix the reference materialises.

    reference models/vitvqgan.py:157-159 / models/vqgan.py:157-159:
    ``(sum zn^2 [T,1] + sum en^2 [K]) - 2 * einsum('bd,nd->bn', zn, en)`` with
    exactly that association.
    """
    row_sq = torch.sum(zn_flat ** 2, dim=1, keepdim=True)
    code_sq = torch.sum(en ** 2, dim=1)
    cross = torch.einsum("bd,nd->bn", zn_flat, en)
    return row_sq + code_sq - 2 * cross


@dataclass
class QuantiserOut:
    z_q: torch.Tensor      # same layout as the input z
    indices: torch.Tensor  # int64; (b, n) for VIT, flat (b*h*w,) for VQGAN
    loss: torch.Tensor     # 0-dim


def _loss(form: str, beta: float, q: torch.Tensor, zn: torch.Tensor) -> torch.Tensor:
    commit = torch.mean((q.detach() - zn) ** 2)     # pulls zn towards the code
    codebook = torch.mean((q - zn.detach()) ** 2)   # pulls the code towards zn
    if form == VIT:       # reference models/vitvqgan.py:166
        return beta * commit + codebook
    if form in (VQGAN, L2):     # reference models/vqgan.py:169
        return commit + beta * codebook
    raise ValueError(form)


def quantise(form: str, z: torch.Tensor, weight: torch.Tensor, beta: float = 0.25) -> QuantiserOut:
    """``Codebook.forward`` of either form; returns (z_q, indices, loss) in the reference order.

    VIT   (reference models/vitvqgan.py:151-171): z is (..., D) token-major.
    VQGAN (reference models/vqgan.py:148-176):    z is (b, D, h, w); the reference
    permutes to (b, h, w, D) as a *view*, normalises the view, flattens (copy),
    and permutes z_q back; indices come back flat in (b, h, w) order.
    """
    dim = weight.shape[1]
    norm = (lambda t: t) if form == L2 else unit_rows        # L2: the same expressions on the raw vectors
    if form in (VQGAN, L2):
        z = z.permute(0, 2, 3, 1)                   # 'b d h w -> b h w d' (view)
    zn = norm(z)
    zn_flat = zn.reshape(-1, dim)                   # .view for VIT, copy for VQGAN
    en = norm(weight)
    d = distance_matrix(zn_flat, en)
    flat_idx = torch.argmin(d, dim=1)
    idx = flat_idx.view(*z.shape[:-1]) if form == VIT else flat_idx
    q = norm(F.embedding(flat_idx, weight)).view(*z.shape)
    loss = _loss(form, beta, q, zn)
    z_q = zn + (q - zn).detach()                    # straight-through estimator
    if form in (VQGAN, L2):
        z_q = z_q.permute(0, 3, 1, 2)               # 'b h w d -> b d h w'
    return QuantiserOut(z_q, idx, loss)


def indices_to_embeddings(form: str, indices: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """reference models/vitvqgan.py:173-176 (gather + l2norm) and
    models/vqgan.py:178-182 (gather only, 'b (h w) d -> b d h w', h = w = int(sqrt(n)))."""
    e = F.embedding(indices, weight)
    if form == VIT:
        return unit_rows(e)
    b, n, dim = e.shape
    side = int(n ** 0.5)
    return e.view(b, side, side, dim).permute(0, 3, 1, 2)


def code_histogram(indices: torch.Tensor, codebook_size: int) -> torch.Tensor:
    """Code-usage histogram.  Not in the reference (north_star addition);
    its oracle is a bincount of the reference indices."""
    return torch.bincount(indices.reshape(-1), minlength=codebook_size)


# --------------------------------------------------------------------------- #
# fwd + bwd through autograd, exactly as the reference trainer drives it
# (trainers/vitgqgan.py:171-184: loss + upstream grad through z_q).
# --------------------------------------------------------------------------- #

@dataclass
class StepOut:
    z_q: torch.Tensor
    indices: torch.Tensor
    loss: torch.Tensor
    grad_z: torch.Tensor
    grad_weight: torch.Tensor


def quantise_step(form: str, z: torch.Tensor, weight: torch.Tensor, beta: float,
                  upstream: torch.Tensor, loss_scale: float = 1.0) -> StepOut:
    """Objective ``(z_q * upstream).sum() + loss_scale * loss`` (cfg 3)."""
    z = z.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    out = quantise(form, z, w, beta)
    objective = (out.z_q * upstream).sum() + loss_scale * out.loss
    gz, gw = torch.autograd.grad(objective, (z, w))
    return StepOut(out.z_q.detach(), out.indices, out.loss.detach(), gz, gw)


def quantise_step_chunked(form: str, z: torch.Tensor, weight: torch.Tensor, beta: float,
                          upstream: torch.Tensor, chunk_tokens: int = 32768) -> StepOut:
    """Same result as :func:`quantise_step` without materialising the full T x K matrix.

    Rows are independent and the loss is a mean over N = T*D elements, so chunk
    c (with T_c tokens) contributes loss_c * T_c/T and gradients scaled likewise
    ("Large shapes").  Chunks split the leading (batch) dim.
    """
    b = z.shape[0]
    tokens_per_item = z[0].numel() // weight.shape[1]
    items = max(1, chunk_tokens // tokens_per_item)
    total = z.numel() // weight.shape[1]
    zq, idx, gz = [], [], []
    gw = torch.zeros_like(weight)
    loss = torch.zeros((), dtype=weight.dtype, device=weight.device)
    for s in range(0, b, items):
        zc, uc = z[s:s + items], upstream[s:s + items]
        frac = (zc.numel() // weight.shape[1]) / total
        o = quantise_step(form, zc, weight, beta, uc, loss_scale=frac)
        zq.append(o.z_q); idx.append(o.indices); gz.append(o.grad_z)
        gw += o.grad_weight
        loss += o.loss * frac
    return StepOut(torch.cat(zq), torch.cat(idx), loss, torch.cat(gz), gw)


def quantise_chunked(form: str, z: torch.Tensor, weight: torch.Tensor, beta: float = 0.25,
                     chunk_tokens: int = 32768) -> QuantiserOut:
    """Forward only, token-chunked (indices are chunk-invariant)."""
    b = z.shape[0]
    tokens_per_item = z[0].numel() // weight.shape[1]
    items = max(1, chunk_tokens // tokens_per_item)
    total = z.numel() // weight.shape[1]
    zq, idx = [], []
    loss = torch.zeros((), dtype=weight.dtype, device=weight.device)
    with torch.no_grad():
        for s in range(0, b, items):
            o = quantise(form, z[s:s + items], weight, beta)
            zq.append(o.z_q); idx.append(o.indices)
            loss += o.loss * ((z[s:s + items].numel() // weight.shape[1]) / total)
    return QuantiserOut(torch.cat(zq), torch.cat(idx), loss)


# --------------------------------------------------------------------------- #
# Closed-form backward, used to check the CUDA backward
# in float64 where autograd round-off would otherwise blur the comparison.
# --------------------------------------------------------------------------- #

def _normalise_backward(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    nrm = x.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    y = x / nrm
    return (g - y * (y * g).sum(-1, keepdim=True)) / nrm


def analytic_backward(form: str, z_tok: torch.Tensor, weight: torch.Tensor, flat_idx: torch.Tensor,
                      beta: float, upstream_tok: torch.Tensor, loss_grad: float = 1.0
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """z_tok / upstream_tok are token-major (T, D).  Returns (grad_z_tok, grad_weight)."""
    T, D = z_tok.shape
    n_elem = T * D
    zn = z_tok / z_tok.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    en = weight / weight.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    q = en[flat_idx]
    commit_w, codebook_w = (beta, 1.0) if form == VIT else (1.0, beta)
    g_zn = upstream_tok + loss_grad * commit_w * 2.0 * (zn - q) / n_elem
    grad_z = _normalise_backward(z_tok, g_zn)
    seg = torch.zeros_like(weight).index_add_(0, flat_idx, q - zn)
    grad_w = _normalise_backward(weight, loss_grad * codebook_w * (2.0 / n_elem) * seg)
    return grad_z, grad_w


# --------------------------------------------------------------------------- #
# Near-tie classification (north_star: rows whose top-2 fp32 distances differ by
# less than 1e-6 relative are counted and reported, not required to match).
# --------------------------------------------------------------------------- #

def top2_relative_gap(zn_flat: torch.Tensor, en: torch.Tensor, chunk: int = 8192) -> torch.Tensor:
    """Per row: (d2 - d1) / max(|d1|, tiny) of the two smallest entries of the fp32 ``d``."""
    gaps = []
    for s in range(0, zn_flat.shape[0], chunk):
        d = distance_matrix(zn_flat[s:s + chunk], en)
        two = torch.topk(d, 2, dim=1, largest=False).values
        gaps.append((two[:, 1] - two[:, 0]) / two[:, 0].abs().clamp_min(1e-30))
    return torch.cat(gaps)


def argmin_fp64(zn_flat: torch.Tensor, en: torch.Tensor, chunk: int = 8192) -> torch.Tensor:
    """Arbiter for near-ties: the same distance in float64."""
    out = []
    z64, e64 = zn_flat.double(), en.double()
    for s in range(0, z64.shape[0], chunk):
        out.append(torch.argmin(distance_matrix(z64[s:s + chunk], e64), dim=1))
    return torch.cat(out)


def classify_index_mismatches(idx_test: torch.Tensor, idx_ref: torch.Tensor, zn_flat: torch.Tensor,
                              en: torch.Tensor, rel_gap: float = 1e-6) -> dict:
    """Split mismatching rows into near-ties (allowed, reported) and hard mismatches (must be 0).

    A mismatching row is a near-tie when the fp32 distances of the two competing
    codes differ by less than ``rel_gap`` relative.
    """
    a, b = idx_test.reshape(-1), idx_ref.reshape(-1)
    rows = torch.nonzero(a != b).flatten()
    hard = 0
    for r in rows.tolist():
        zr = zn_flat[r:r + 1]
        pair = torch.stack([en[a[r]], en[b[r]]])
        d = distance_matrix(zr, pair).flatten()
        gap = (d[0] - d[1]).abs() / d.abs().min().clamp_min(1e-30)
        if not (gap < rel_gap):
            hard += 1
    return {"mismatch_rows": int(rows.numel()), "near_tie_rows": int(rows.numel()) - hard, "hard_rows": hard}


# --------------------------------------------------------------------------- #
# Seeded synthetic inputs.  CPU generator, so the same bits
# are produced in this container and on the GPU box.
# --------------------------------------------------------------------------- #

def make_codebook(form: str, K: int, D: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if form == VIT:   # reference models/vitvqgan.py:149  weight.normal_()
        return torch.randn(K, D, generator=g)
    # reference models/vqgan.py:146  weight.uniform_(-1/K, 1/K)
    return (torch.rand(K, D, generator=g) * 2 - 1) / K


def make_latents(shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def make_trained_like(weight: torch.Tensor, tokens: int, seed: int, sigma: float = 0.05) -> torch.Tensor:
    """z = en[randint] + sigma * N(0,1): skewed histograms, cancellation in the codebook gradient."""
    g = torch.Generator().manual_seed(seed)
    en = unit_rows(weight)
    pick = torch.randint(0, weight.shape[0], (tokens,), generator=g)
    return en[pick] + sigma * torch.randn(tokens, weight.shape[1], generator=g)


# ---------------------------------------------------------------------------------------------------
# first consumer of the tokens
# ---------------------------------------------------------------------------------------------------
def token_embed_inputs(vocab: int, dim: int, b: int, n: int, seed: int):
    """Seeded (table (vocab + 1, dim), pos_enc (1, n, dim), tokens (b, n), mask (b, n)) of the token-consumer fixture;
    the same draws, in the same order, as oracle/make_golden.py::token_embed_case."""
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(vocab + 1, dim, generator=g)
    pos = 0.02 * torch.randn(1, n, dim, generator=g)
    tokens = torch.randint(0, vocab, (b, n), generator=g)
    mask = torch.rand(b, n, generator=g) < 0.4
    return table, pos, tokens, mask


def masked_token_embeddings(tokens: torch.Tensor, mask, mask_token_id: int, table: torch.Tensor, pos_enc=None,
                            ignore_index: int = -1):
    """/root/reference/models/muse.py:149-150 (= models/maskgit.py:131-132): the two masked_fill lines;
    /root/reference/models/muse.py:90-91 (= models/maskgit.py:80-81): embedding lookup, ``+= pos_enc``."""
    if mask is None:
        mask = torch.zeros_like(tokens, dtype=torch.bool)
    input_ids = tokens.masked_fill(mask, mask_token_id)
    labels = tokens.masked_fill(~mask, ignore_index)
    embeds = torch.nn.functional.embedding(input_ids, table)
    if pos_enc is not None:
        embeds = embeds + pos_enc
    return embeds, input_ids, labels


def causal_token_embeddings(tokens: torch.Tensor, table: torch.Tensor, pe, start_token: torch.Tensor):
    """/root/reference/models/parti.py:98-106: the decoder input of the autoregressive model --
    ``token_emb(tokens[:, :-1])`` (``:100``), ``+ pe[:seq_len]`` (``:102`` via models/positional_encoding.py:40-41; its
    dropout is the identity in eval mode), start token in front (``:104-105``); the labels are the tokens (``:98``)."""
    inp, labels = tokens[:, :-1], tokens
    e = torch.nn.functional.embedding(inp, table)
    if pe is not None:
        e = e + pe[: e.size(1)]
    start = start_token.reshape(1, 1, -1).expand(tokens.shape[0], 1, -1)
    return torch.cat((start, e), dim=1), labels


# ---------------------------------------------------------------------------------------------------
# pre_quant / post_quant around the quantiser
# ---------------------------------------------------------------------------------------------------
def projection_inputs(C: int, D: int, seed: int, conv: bool = False):
    """Seeded weights of one projection: nn.Linear(C -> D) (``(D, C)``, ``(D,)``) or, with ``conv``, the 1x1
    nn.Conv2d(C -> D) (``(D, C, 1, 1)``, ``(D,)``); the same draws, in the same order, as
    oracle/make_golden.py::projected_case (uniform(-1/sqrt(C), 1/sqrt(C)) like both modules' default init)."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (C ** 0.5)
    w = (torch.rand(D, C, generator=g) * 2 - 1) * bound
    b = (torch.rand(D, generator=g) * 2 - 1) * bound
    return (w.reshape(D, C, 1, 1) if conv else w), b


def quantise_projected(x: torch.Tensor, w_pre: torch.Tensor, b_pre, weight: torch.Tensor, beta: float = 0.25):
    """/root/reference/models/vitvqgan.py:192-193: ``enc_imgs = self.pre_quant(enc_imgs)`` (nn.Linear, ``:185``), then
    ``self.codebook(enc_imgs)``.  Returns ``(z, QuantiserOut)``."""
    z = torch.nn.functional.linear(x, w_pre, b_pre)
    return z, quantise(VIT, z, weight, beta)


def decode_projected(form: str, indices: torch.Tensor, weight: torch.Tensor, w_post: torch.Tensor, b_post):
    """/root/reference/models/vitvqgan.py:199-200 (``post_quant`` = nn.Linear, ``:187``) and
    /root/reference/models/vqgan.py:241-242 (``post_quant`` = nn.Conv2d(dim, dim, 1), ``:228``):
    ``embeds = self.codebook.indices_to_embeddings(indices); embeds = self.post_quant(embeds)``."""
    e = indices_to_embeddings(form, indices, weight)
    if form == VIT:
        return torch.nn.functional.linear(e, w_post, b_post)
    return torch.nn.functional.conv2d(e, w_post.reshape(w_post.shape[0], -1, 1, 1), b_post)
