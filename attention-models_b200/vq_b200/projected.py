"""pre_quant / post_quant fused with the quantiser.

The reference wraps the quantiser in two projections (/root/reference/models/vitvqgan.py:185-187: ``nn.Linear``;
/root/reference/models/vqgan.py:226-228: 1x1 ``nn.Conv2d``) and calls them back to back with it
(vitvqgan.py:192-194, 199-200, 207-208; vqgan.py:233-235, 241-242, 248-249).

* ``quantise_projected`` / ``encode_indices_projected`` (ViT form, ``nn.Linear(C, 32)``): ``vq_forward_projected`` forms
  z = x W^T + b inside the token preparation -- the encoder rows are read once and z never travels through HBM.  The
  backward is the quantiser's own (``vq_backward``) followed by the two GEMMs of the Linear (cuBLAS, plain library GEMMs).
* ``ProjectedTable`` (both forms): ``post_quant(indices_to_embeddings(indices))`` as one gather from the (K, C) table of
  projected codes (``vq_project_codebook`` + ``vq_gather_projected``).

Shapes the fused forward does not cover (``prequant_supported`` false: D != 32, C not a multiple of 64, C > 768) are
for the caller to run unfused (``pre_quant`` then ``Codebook.forward``); nothing here falls back silently.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from . import functional as F_vq
from ._lib import FLAG_EXACT_SCAN, FLAG_INDICES_ONLY, FORM_VIT, LAYOUT_NCHW, LAYOUT_TOKEN_MAJOR, STATS_LEN
from .functional import PreparedCodebook, _ptr, _require_cuda, _scratch, _stream


def prequant_supported(in_features: int, codebook_dim: int) -> bool:
    """Whether ``vq_forward_projected`` covers a ``Linear(in_features, codebook_dim)`` (host logic only)."""
    return bool(_lib.load().vq_prequant_supported(int(in_features), int(codebook_dim)))


def _rows(x: torch.Tensor, C: int) -> torch.Tensor:
    if x.shape[-1] != C:
        raise ValueError(f"last dim of x is {x.shape[-1]}, pre_quant.in_features is {C}")
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


class _QuantiseProjected(torch.autograd.Function):
    """(x, W_pre, b_pre, codebook weight) -> (z_q, flat indices, loss, histogram, stats)."""

    @staticmethod
    def forward(ctx, x, w_pre, b_pre, weight, prepared, beta, flags, refresh, want_z):
        lib = _lib.load()
        dev = x.device
        K, D = prepared.K, prepared.D
        C = w_pre.shape[1]
        T = x.numel() // C
        lead = tuple(x.shape[:-1])
        need_grad = any(ctx.needs_input_grad[:4])
        z_q = torch.empty(*lead, D, dtype=torch.float32, device=dev)
        idx = torch.empty(T, dtype=torch.int64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        hist = torch.empty(K, dtype=torch.int32, device=dev)
        stats = torch.empty(STATS_LEN, dtype=torch.int64, device=dev)
        saved_zn = torch.empty(T, D, dtype=torch.float32, device=dev) if need_grad else None
        saved_denom = torch.empty(T, dtype=torch.float32, device=dev) if need_grad else None
        seg = torch.empty(K * D + K, dtype=torch.int64, device=dev) if ctx.needs_input_grad[3] else None
        z_out = torch.empty(*lead, D, dtype=torch.float32, device=dev) if want_z else None
        ws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, flags)
        ws = _scratch(ws_bytes, dev)
        if refresh and need_grad:
            prepared.blob = torch.empty_like(prepared.blob)      # see _Quantise.forward
        w_pre_c = w_pre.detach().contiguous()
        b_pre_c = None if b_pre is None else b_pre.detach().contiguous().float()
        with torch.cuda.device(dev):
            w_c = weight.detach().contiguous() if refresh else None
            _lib.check(lib.vq_forward_projected(_ptr(x), C, _ptr(w_pre_c), _ptr(b_pre_c), T, _ptr(w_c), _ptr(prepared.blob),
                                                K, D, FORM_VIT, float(beta), flags, max(T * D, 1), _ptr(z_q), _ptr(idx),
                                                _ptr(loss), _ptr(hist), _ptr(stats), _ptr(saved_zn), _ptr(saved_denom),
                                                _ptr(seg), _ptr(z_out), _ptr(ws), ws_bytes, _stream(dev)))
        if refresh:
            prepared.mark_current(weight)
        if need_grad:
            ctx.save_for_backward(saved_zn, saved_denom, idx, prepared.blob, hist,
                                  seg if seg is not None else saved_denom.new_empty(0), x, w_pre)
        ctx.has_seg = seg is not None
        ctx.has_bias = b_pre is not None
        ctx.meta = (FORM_VIT, float(beta), LAYOUT_TOKEN_MAJOR, T, 0, K, D, max(T * D, 1), (T, D))
        ctx.x_shape = tuple(x.shape)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*(t for t in (idx, hist, stats, z_out) if t is not None))
        return z_q, idx, loss.view(()), hist, stats, z_out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_zq, g_idx, g_loss, g_hist, g_stats, g_z):
        *saved, x, w_pre = ctx.saved_tensors
        want_x, want_wp, want_bp, want_w = ctx.needs_input_grad[:4]
        want_bp = want_bp and ctx.has_bias
        T, D = ctx.meta[3], ctx.meta[6]
        if g_zq is not None:
            g_zq = g_zq.reshape(T, D)
        # grad of the projected rows z (straight-through + commitment, through the normalisation), grad of the codebook
        grad_z, grad_w = F_vq._quantise_backward(saved, ctx.meta, ctx.has_seg, g_zq, g_loss,
                                                 want_x or want_wp or want_bp, want_w)
        grad_x = grad_wp = grad_bp = None
        # the Linear's own backward (models/vitvqgan.py:185): two plain GEMMs and a column sum
        if want_x:
            grad_x = torch.mm(grad_z, w_pre).view(ctx.x_shape)
        if want_wp:
            grad_wp = torch.mm(grad_z.t(), x.reshape(T, -1))
        if want_bp:
            grad_bp = grad_z.sum(0)
        return grad_x, grad_wp, grad_bp, grad_w, None, None, None, None, None


def _check_projection(x, w_pre, b_pre, weight):
    _require_cuda(x, "x")
    _require_cuda(w_pre, "pre_quant.weight")
    _require_cuda(weight, "the codebook weight")
    if w_pre.dim() != 2 or w_pre.dtype != torch.float32:
        raise TypeError("pre_quant.weight must be a 2-D float32 tensor (codebook_dim, in_features)")
    D, C = w_pre.shape
    if weight.dim() != 2 or weight.shape[1] != D:
        raise ValueError(f"pre_quant projects to {D} features, the codebook holds {tuple(weight.shape)}")
    if b_pre is not None and tuple(b_pre.shape) != (D,):
        raise ValueError(f"pre_quant.bias must have shape ({D},)")
    if not prequant_supported(C, D):
        raise ValueError(f"fused pre_quant covers Linear(C, 32) with C a multiple of 64 up to 768; got Linear({C}, {D}). "
                         "Run pre_quant and Codebook.forward unfused for this shape.")
    return C, D


def quantise_projected(x: torch.Tensor, w_pre: torch.Tensor, b_pre: Optional[torch.Tensor], weight: torch.Tensor,
                       beta: float = 0.25, prepared: Optional[PreparedCodebook] = None, exact_scan: bool = False,
                       always_refresh: bool = False, return_z: bool = False):
    """``Codebook(pre_quant(x))`` of ViTVQGAN.forward (vitvqgan.py:192-193) in one pass over ``x`` (..., C).

    Returns ``(z_q (..., D), flat_indices, loss, histogram, stats)`` (+ the projected rows ``z`` with ``return_z``).
    Differentiable with respect to ``x``, ``w_pre``, ``b_pre`` and ``weight``."""
    C, D = _check_projection(x, w_pre, b_pre, weight)
    if x.numel() == 0:          # no rows: what the reference returns (empty z_q / indices, NaN loss); nothing is launched
        out = F_vq._empty_result((*x.shape[:-1], D), weight.shape[0], x.device)
        return (*out, torch.empty((*x.shape[:-1], D), dtype=torch.float32, device=x.device)) if return_z else out
    refresh = False
    if prepared is None or not prepared.fits(weight) or prepared.raw:
        prepared = F_vq.prepare_codebook(weight)
    elif always_refresh or not prepared.matches(weight):
        refresh = True
    flags = FLAG_EXACT_SCAN if exact_scan else 0
    out = _QuantiseProjected.apply(_rows(x, C), w_pre, b_pre, weight, prepared, beta, flags, refresh, return_z)
    return out if return_z else out[:5]


@torch.no_grad()
def encode_indices_projected(x: torch.Tensor, w_pre: torch.Tensor, b_pre: Optional[torch.Tensor], weight: torch.Tensor,
                             prepared: Optional[PreparedCodebook] = None, exact_scan: bool = False,
                             index_dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """``encode_imgs`` behind the encoder (vitvqgan.py:207-209): flat indices of ``pre_quant(x)``, nothing else."""
    C, D = _check_projection(x, w_pre, b_pre, weight)
    if x.numel() == 0:
        return torch.empty(0, dtype=index_dtype, device=x.device)
    if prepared is None or not prepared.matches(weight):
        prepared = F_vq.prepare_codebook(weight)
    lib = _lib.load()
    x = _rows(x, C)
    dev = x.device
    K = prepared.K
    T = x.numel() // C
    bits = F_vq.token_bits(index_dtype)
    flags = FLAG_INDICES_ONLY | (FLAG_EXACT_SCAN if exact_scan else 0) | {64: 0, 32: _lib.FLAG_IDX32, 16: _lib.FLAG_IDX16}[bits]
    idx = torch.empty(T, dtype=index_dtype, device=dev)
    ws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, flags)
    ws = _scratch(ws_bytes, dev)
    w_pre_c = w_pre.detach().contiguous()
    b_pre_c = None if b_pre is None else b_pre.detach().contiguous().float()
    with torch.cuda.device(dev):
        _lib.check(lib.vq_forward_projected(_ptr(x), C, _ptr(w_pre_c), _ptr(b_pre_c), T, None, _ptr(prepared.blob), K, D,
                                            FORM_VIT, 0.25, flags, max(T * D, 1), None, _ptr(idx), None, None, None, None,
                                            None, None, None, _ptr(ws), ws_bytes, _stream(dev)))
    return idx


class ProjectedTable:
    """The K codes behind ``post_quant``: ``table[k] = post_quant(l2norm(E_k))`` (ViT form, ``nn.Linear``) or
    ``post_quant(E_k)`` (CNN form, 1x1 ``nn.Conv2d``), built once per (codebook, post_quant) state; ``gather`` is then
    ``decode_indices`` up to the decoder (vitvqgan.py:199-200, vqgan.py:241-242)."""

    def __init__(self, weight: torch.Tensor, w_post: torch.Tensor, b_post: Optional[torch.Tensor], form: str = "vit",
                 prepared: Optional[PreparedCodebook] = None):
        _require_cuda(weight, "the codebook weight")
        _require_cuda(w_post, "post_quant.weight")
        if form not in ("vit", "vqgan"):
            raise ValueError("form must be 'vit' or 'vqgan'")
        K, D = weight.shape
        w2 = w_post.detach().reshape(w_post.shape[0], -1).contiguous().float()     # (C, D) or (C, D, 1, 1)
        if w2.shape[1] != D:
            raise ValueError(f"post_quant takes {w2.shape[1]} features, the codebook has {D}")
        C = w2.shape[0]
        b2 = None if b_post is None else b_post.detach().contiguous().float()
        self.form, self.K, self.D, self.C = form, K, D, C
        self.table = torch.empty(K, C, dtype=torch.float32, device=weight.device)
        self._key = self.key_of(weight, w_post, b_post)
        lib = _lib.load()
        w = weight.detach().contiguous()
        with torch.cuda.device(weight.device):
            if form == "vit":
                if prepared is None or not prepared.matches(weight):
                    prepared = F_vq.prepare_codebook(weight)
                _lib.check(lib.vq_project_codebook(None, _ptr(prepared.blob), K, D, 1, _ptr(w2), _ptr(b2), C, _ptr(self.table),
                                                   _stream(weight.device)))
            else:
                _lib.check(lib.vq_project_codebook(_ptr(w), None, K, D, 0, _ptr(w2), _ptr(b2), C, _ptr(self.table),
                                                   _stream(weight.device)))

    @staticmethod
    def key_of(weight, w_post, b_post):
        return tuple((t.data_ptr(), t._version, t.device) for t in (weight, w_post, b_post) if t is not None)

    def matches(self, weight, w_post, b_post) -> bool:
        return self._key == self.key_of(weight, w_post, b_post)

    @torch.no_grad()
    def gather(self, indices: torch.Tensor, check_indices: bool = True) -> torch.Tensor:
        """(b, n) tokens (int64 / int32 / uint16) -> (b, n, C) rows (ViT form) or (b, C, h, w) maps (CNN form)."""
        _require_cuda(indices, "indices")
        idx = F_vq.as_tokens(indices)
        dev = idx.device
        T = idx.numel()
        if self.form == "vit":
            out = torch.empty(*idx.shape, self.C, dtype=torch.float32, device=dev)
            layout, hw = LAYOUT_TOKEN_MAJOR, 0
        else:
            if idx.dim() != 2:
                raise ValueError("vqgan decode expects (b, n) indices")
            b, n = idx.shape
            side = int(n ** 0.5)
            if side * side != n:
                raise ValueError(f"n={n} is not a perfect square")
            out = torch.empty(b, self.C, side, side, dtype=torch.float32, device=dev)
            layout, hw = LAYOUT_NCHW, n
        stats = torch.zeros(STATS_LEN, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().vq_gather_projected(_ptr(idx), F_vq.token_bits(idx.dtype), T, hw, _ptr(self.table), self.K,
                                                       self.C, layout, _ptr(out), _ptr(stats), _stream(dev)))
        if check_indices and int(stats[_lib.STAT_BAD_INDEX].item()) != 0:
            raise IndexError("index out of range in self")
        return out
