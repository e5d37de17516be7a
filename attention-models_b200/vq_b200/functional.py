"""Functional API over the C ABI: prepared codebook, differentiable quantise, encode, decode.

Mirrors what the reference's two ``Codebook.forward`` / ``indices_to_embeddings`` compute
(/root/reference/models/vitvqgan.py:151-176, /root/reference/models/vqgan.py:148-182); all arithmetic
runs in libvq_b200.so on the current CUDA stream.  PyTorch only allocates the buffers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import (FLAG_EXACT_SCAN, FLAG_INDICES_ONLY, FORM_VIT, FORM_VQGAN, FORM_VQGAN_L2, LAYOUT_NCHW, LAYOUT_TOKEN_MAJOR,
                   STATS_LEN)

# "l2": the CNN form without l2 normalisation (plain squared-L2 on raw vectors; not a form of the reference)
FORMS = {"vit": FORM_VIT, "vqgan": FORM_VQGAN, "l2": FORM_VQGAN_L2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"vq_b200: {what} must live on a CUDA device (got {t.device}); there is no CPU path")


def _scratch(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


@dataclass
class PreparedCodebook:
    """Unit codes (fp32 + fp16), squared norms and row norms of one weight tensor, in device memory."""
    blob: torch.Tensor
    K: int
    D: int
    weight_ptr: int
    weight_version: int
    raw: bool = False          # prepared for the un-normalised form ("l2"): the codes as they are

    def matches(self, weight: torch.Tensor, raw: bool = False) -> bool:
        return (self.weight_ptr == weight.data_ptr() and self.weight_version == weight._version
                and self.blob.device == weight.device and self.raw == raw)

    def fits(self, weight: torch.Tensor) -> bool:
        """Same shape and device: the blob can be re-filled in place (by vq_forward itself, see quantise)."""
        return self.blob.device == weight.device and tuple(weight.shape) == (self.K, self.D)

    def mark_current(self, weight: torch.Tensor) -> None:
        self.weight_ptr, self.weight_version = weight.data_ptr(), weight._version


def prepare_codebook(weight: torch.Tensor, raw: bool = False) -> PreparedCodebook:
    """``l2_norm(embedding.weight)`` and ``sum(embedd_norm**2, 1)`` (reference vitvqgan.py:154,158); ``raw``: the codes
    as they are and ``sum(weight**2, 1)`` for the un-normalised form."""
    _require_cuda(weight, "the codebook weight")
    if weight.dtype != torch.float32 or weight.dim() != 2:
        raise TypeError("codebook weight must be a 2-D float32 tensor")
    w = weight.detach().contiguous()
    K, D = w.shape
    nbytes = _lib.size_query("vq_codebook_bytes", K, D)
    blob = _scratch(nbytes, w.device)
    with torch.cuda.device(w.device):
        fn = _lib.load().vq_codebook_prepare_raw if raw else _lib.load().vq_codebook_prepare
        _lib.check(fn(_ptr(w), K, D, _ptr(blob), nbytes, _stream(w.device)))
    return PreparedCodebook(blob, K, D, weight.data_ptr(), weight._version, raw)


def _token_geometry(z: torch.Tensor, layout: int, D: int) -> Tuple[int, int]:
    if layout == LAYOUT_TOKEN_MAJOR:
        if z.shape[-1] != D:
            raise ValueError(f"last dim of z is {z.shape[-1]}, codebook_dim is {D}")
        return z.numel() // D, 0
    if z.dim() != 4 or z.shape[1] != D:
        raise ValueError(f"NCHW input must be (b, {D}, h, w); got {tuple(z.shape)}")
    hw = z.shape[2] * z.shape[3]
    return z.shape[0] * hw, hw


class _Quantise(torch.autograd.Function):
    """(z, weight) -> (z_q, flat indices, loss, histogram, stats) with the straight-through backward."""

    @staticmethod
    def forward(ctx, z, weight, prepared, form, beta, layout, flags, n_elem_total, sorted_segments, refresh):
        lib = _lib.load()
        dev = z.device
        K, D = prepared.K, prepared.D
        T, hw = _token_geometry(z, layout, D)
        n_total = int(n_elem_total) if n_elem_total else max(T * D, 1)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        z_q = torch.empty_like(z)
        idx = torch.empty(T, dtype=torch.int64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        hist = torch.empty(K, dtype=torch.int32, device=dev)
        stats = torch.empty(STATS_LEN, dtype=torch.int64, device=dev)
        saved_zn = torch.empty(T, D, dtype=torch.float32, device=dev) if need_grad else None
        saved_denom = torch.empty(T, dtype=torch.float32, device=dev) if need_grad else None
        # codebook-gradient segment sums: accumulated by the forward's finish pass (it holds q - zn), unless the
        # caller asks for the bucketed sum of the backward (sorted_segments; same integers, skew-insensitive)
        seg = (torch.empty(K * D + K, dtype=torch.int64, device=dev)
               if ctx.needs_input_grad[1] and not sorted_segments else None)
        ws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, flags)
        ws = _scratch(ws_bytes, dev)
        if refresh and need_grad:
            # an earlier forward whose backward has not run yet saved the old blob: refill a fresh one instead of
            # rewriting that one in place (autograd's version counters cannot see a write made by the library)
            prepared.blob = torch.empty_like(prepared.blob)
        with torch.cuda.device(dev):
            # refresh: the weights changed since `prepared` was filled -- the forward re-prepares the codebook itself
            # (in the launch that normalises the token rows where the shape allows)
            w_c = weight.detach().contiguous() if refresh else None      # kept alive until the launch is queued
            w_ptr = _ptr(w_c)
            _lib.check(lib.vq_forward(_ptr(z), layout, T, hw, w_ptr, _ptr(prepared.blob), K, D, form, float(beta), flags,
                                      n_total, _ptr(z_q), _ptr(idx), _ptr(loss), _ptr(hist), _ptr(stats),
                                      _ptr(saved_zn), _ptr(saved_denom), _ptr(seg), _ptr(ws), ws_bytes, _stream(dev)))
        if refresh:
            prepared.mark_current(weight)
        if need_grad:
            ctx.save_for_backward(saved_zn, saved_denom, idx, prepared.blob, hist,
                                  seg if seg is not None else saved_denom.new_empty(0))
        ctx.has_seg = seg is not None
        ctx.meta = (form, float(beta), layout, T, hw, K, D, n_total, tuple(z.shape))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(idx, hist, stats)
        return z_q, idx, loss.view(()), hist, stats

    @staticmethod
    @once_differentiable
    def backward(ctx, g_zq, g_idx, g_loss, g_hist, g_stats):
        grad_z, grad_w = _quantise_backward(ctx.saved_tensors, ctx.meta, ctx.has_seg, g_zq, g_loss,
                                            ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return grad_z, grad_w, None, None, None, None, None, None, None, None


def _quantise_backward(saved, meta, has_seg, g_zq, g_loss, want_z, want_w):
    """grad_z / grad_weight of one quantiser forward from what it saved (vq_backward*); shared by ``_Quantise`` and
    the projected forward (vq_b200/projected.py)."""
    lib = _lib.load()
    saved_zn, saved_denom, idx, blob, hist, seg_saved = saved
    form, beta, layout, T, hw, K, D, n_total, z_shape = meta
    dev = saved_zn.device
    if g_zq is not None:
        g_zq = g_zq.contiguous().float()
    g_loss = torch.zeros(1, dtype=torch.float32, device=dev) if g_loss is None else g_loss.reshape(1).float()
    grad_z = torch.empty(z_shape, dtype=torch.float32, device=dev) if want_z else None
    seg, from_forward = (seg_saved if has_seg else None), has_seg
    if want_w and seg is None:
        seg = torch.empty(K * D + K, dtype=torch.int64, device=dev)
    grad_w = torch.empty(K, D, dtype=torch.float32, device=dev) if want_w else None
    ws_bytes = _lib.size_query("vq_backward_workspace_bytes", T, K, D)
    ws = _scratch(ws_bytes, dev)
    with torch.cuda.device(dev):
        s = _stream(dev)
        if want_w and from_forward:        # one launch: grad_z + grad_weight from the forward's segment sums
            _lib.check(lib.vq_backward(_ptr(g_zq), layout, T, hw, _ptr(saved_zn), _ptr(saved_denom), _ptr(idx), _ptr(blob),
                                       K, D, form, beta, _ptr(g_loss), n_total, _ptr(seg), None, _ptr(grad_z), _ptr(grad_w),
                                       None, _ptr(ws), ws_bytes, s))
            return grad_z, grad_w
        _lib.check(lib.vq_backward_tokens(_ptr(g_zq), layout, T, hw, _ptr(saved_zn), _ptr(saved_denom), _ptr(idx),
                                          _ptr(hist), _ptr(blob), K, D, form, beta, _ptr(g_loss), n_total, _ptr(grad_z),
                                          None if from_forward or not want_w else _ptr(seg), _ptr(ws), ws_bytes, s))
        if want_w:
            _lib.check(lib.vq_backward_codebook(_ptr(seg), _ptr(blob), K, D, form, beta, _ptr(g_loss), n_total,
                                                _ptr(grad_w), None, None, s))
    return grad_z, grad_w


def _empty_result(z_q_shape, K: int, device):
    """What the reference returns for a batch without tokens (argmin over a (0, K) matrix, ``torch.mean`` of nothing): empty
    z_q and indices, a NaN loss; no kernel is launched (the ABI takes no NULL buffers)."""
    z_q = torch.empty(z_q_shape, dtype=torch.float32, device=device)
    idx = torch.empty(0, dtype=torch.int64, device=device)
    loss = torch.full((), float("nan"), dtype=torch.float32, device=device)
    hist = torch.zeros(K, dtype=torch.int32, device=device)
    stats = torch.zeros(STATS_LEN, dtype=torch.int64, device=device)
    return z_q, idx, loss, hist, stats


def _as_fp32_input(z: torch.Tensor) -> torch.Tensor:
    # under autocast the pre_quant projection hands over bf16; the reference's F.normalize returns fp32,
    # so all quantiser maths is fp32 (SURVEY.md section 8b "modes")
    if z.dtype != torch.float32:
        z = z.float()
    return z.contiguous()


def quantise(z: torch.Tensor, weight: torch.Tensor, form: str = "vit", beta: float = 0.25,
             prepared: Optional[PreparedCodebook] = None, exact_scan: bool = False,
             n_elem_total: Optional[int] = None, sorted_segments: bool = False, always_refresh: bool = False):
    """Full forward.  Returns ``(z_q, flat_indices, loss, histogram, stats)``.

    ``z``: (..., D) for ``form='vit'`` (token-major) or (b, D, h, w) for ``form='vqgan'``.
    Differentiable w.r.t. ``z`` (straight-through + commitment) and ``weight`` (codebook term).
    """
    _require_cuda(z, "z")
    _require_cuda(weight, "the codebook weight")
    if z.numel() == 0:
        return _empty_result(tuple(z.shape), weight.shape[0], z.device)
    refresh = False
    raw = form == "l2"
    if prepared is None or not prepared.fits(weight) or prepared.raw != raw:
        prepared = prepare_codebook(weight, raw)
    elif always_refresh or not prepared.matches(weight, raw):
        # stale (or not provably current: an edit through `weight.data` leaves the version counter alone) contents of the
        # right size: vq_forward refills the blob in the launch that normalises the rows (no extra launch)
        refresh = True
    layout = LAYOUT_TOKEN_MAJOR if form == "vit" else LAYOUT_NCHW
    flags = FLAG_EXACT_SCAN if exact_scan else 0
    return _Quantise.apply(_as_fp32_input(z), weight, prepared, FORMS[form], beta, layout, flags, n_elem_total,
                           sorted_segments, refresh)


_TOKEN_BITS = {torch.int64: 64, torch.int32: 32, torch.uint16: 16}


def token_bits(dtype: torch.dtype) -> int:
    """Wire formats of token indices the library reads and writes: int64 (the reference's), int32, uint16."""
    try:
        return _TOKEN_BITS[dtype]
    except KeyError:
        raise TypeError(f"token dtype {dtype}: torch.int64, torch.int32 or torch.uint16") from None


def as_tokens(t: torch.Tensor) -> torch.Tensor:
    """A contiguous tensor in one of the wire formats (other integer dtypes are widened to int64)."""
    return (t if t.dtype in _TOKEN_BITS else t.to(torch.int64)).contiguous()


def convert_tokens(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Tokens in another wire format (vq_tokens_convert; values must fit the target)."""
    t = as_tokens(t)
    if t.dtype == dtype:
        return t
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().vq_tokens_convert(_ptr(t), token_bits(t.dtype), _ptr(out), token_bits(dtype), t.numel(),
                                                 _stream(t.device)))
    return out


@torch.no_grad()
def encode_indices(z: torch.Tensor, weight: torch.Tensor, form: str = "vit",
                   prepared: Optional[PreparedCodebook] = None, exact_scan: bool = False,
                   want_hist: bool = False, index_dtype: torch.dtype = torch.int64):
    """``encode_imgs`` fast path: flat indices only (z_q, loss and the saved state are skipped).  ``index_dtype``:
    torch.int64 as the reference returns them, or the narrow wire formats torch.int32 / torch.uint16 (K <= 65536) that
    ``indices_to_embeddings`` and the token consumers read directly."""
    _require_cuda(z, "z")
    if z.numel() == 0:
        idx = torch.empty(0, dtype=index_dtype, device=z.device)
        return (idx, torch.zeros(weight.shape[0], dtype=torch.int32, device=z.device)) if want_hist else idx
    if prepared is None or not prepared.matches(weight, form == "l2"):
        prepared = prepare_codebook(weight, form == "l2")
    lib = _lib.load()
    z = _as_fp32_input(z)
    dev = z.device
    layout = LAYOUT_TOKEN_MAJOR if form == "vit" else LAYOUT_NCHW
    K, D = prepared.K, prepared.D
    T, hw = _token_geometry(z, layout, D)
    bits = token_bits(index_dtype)
    flags = FLAG_INDICES_ONLY | (FLAG_EXACT_SCAN if exact_scan else 0) | {64: 0, 32: _lib.FLAG_IDX32, 16: _lib.FLAG_IDX16}[bits]
    idx = torch.empty(T, dtype=index_dtype, device=dev)
    hist = torch.empty(K, dtype=torch.int32, device=dev) if want_hist else None
    ws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, flags)
    ws = _scratch(ws_bytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.vq_forward(_ptr(z), layout, T, hw, None, _ptr(prepared.blob), K, D, FORMS[form], 0.25, flags,
                                  max(T * D, 1), None, _ptr(idx), None, _ptr(hist), None, None, None, None, _ptr(ws),
                                  ws_bytes, _stream(dev)))
    return (idx, hist) if want_hist else idx


def _embedding_backward(ids: torch.Tensor, grad_out: torch.Tensor, vocab: int, dim: int, *, ids_per_seq: int = 0,
                        rows_per_seq: int = 0, row_shift: int = 0, hw: int = 0,
                        prepared: Optional[PreparedCodebook] = None) -> torch.Tensor:
    """grad_table (vocab, dim) of a lookup, through vq_embedding_backward (integer-accumulated: deterministic)."""
    dev = grad_out.device
    g = grad_out.contiguous().float()
    ids = ids.contiguous()
    n_rows = g.numel() // dim
    out = torch.empty(vocab, dim, dtype=torch.float32, device=dev)
    ws_bytes = _lib.size_query("vq_embedding_backward_bytes", vocab, dim)
    ws = _scratch(ws_bytes, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().vq_embedding_backward(_ptr(ids), ids.numel(), ids_per_seq, rows_per_seq, row_shift, _ptr(g), n_rows,
                                                     hw, vocab, dim, None if prepared is None else _ptr(prepared.blob), _ptr(out),
                                                     _ptr(ws), ws_bytes, _stream(dev)))
    return out


class _Decode(torch.autograd.Function):
    """indices -> embeddings, differentiable w.r.t. the codebook weights like the reference's nn.Embedding lookup
    (models/vitvqgan.py:173-176: the gradient also passes through the l2 normalisation; models/vqgan.py:178-182)."""

    @staticmethod
    def forward(ctx, weight, idx, form, prepared, stats):
        lib = _lib.load()
        dev = idx.device
        K, D = weight.shape
        T = idx.numel()
        w = weight.detach().contiguous()
        bits = token_bits(idx.dtype)
        if form == "vit":
            out = torch.empty(*idx.shape, D, dtype=torch.float32, device=dev)
            args = (_ptr(idx), bits, T, 0, None, _ptr(prepared.blob), K, D, 1, LAYOUT_TOKEN_MAJOR)
        else:
            b, n = idx.shape
            side = int(n ** 0.5)
            out = torch.empty(b, D, side, side, dtype=torch.float32, device=dev)
            args = (_ptr(idx), bits, T, n, _ptr(w), None, K, D, 0, LAYOUT_NCHW)
        with torch.cuda.device(dev):
            _lib.check(lib.vq_gather_tokens(*args, _ptr(out), _ptr(stats), _stream(dev)))
        ctx.form, ctx.prepared, ctx.shape = form, prepared, (K, D)
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        idx = convert_tokens(idx, torch.int64)
        K, D = ctx.shape
        if ctx.form == "vit":
            gw = _embedding_backward(idx, g, K, D, prepared=ctx.prepared)
        else:
            gw = _embedding_backward(idx, g, K, D, hw=idx.shape[1])
        return gw, None, None, None, None


def indices_to_embeddings(indices: torch.Tensor, weight: torch.Tensor, form: str = "vit",
                          prepared: Optional[PreparedCodebook] = None, check_indices: bool = True) -> torch.Tensor:
    """Decode gather (reference vitvqgan.py:173-176: l2norm(E[i]); vqgan.py:178-182: E[i] as (b, D, h, w)).

    ``indices``: (b, n) integer tensor (int64, or the narrow wire formats int32 / uint16).  Out-of-range indices raise IndexError like the reference's CPU
    path (one host sync for the check; pass ``check_indices=False`` to skip it).  Differentiable with respect to
    ``weight`` (as the reference's ``nn.Embedding`` lookup is) when gradients are enabled and it requires them.
    """
    _require_cuda(indices, "indices")
    _require_cuda(weight, "the codebook weight")
    dev = indices.device
    idx = as_tokens(indices)
    if form == "vit":
        if prepared is None or not prepared.matches(weight):
            prepared = prepare_codebook(weight)
    else:
        if idx.dim() != 2:
            raise ValueError("vqgan indices_to_embeddings expects (b, n) indices")
        n = idx.shape[1]
        side = int(n ** 0.5)
        if side * side != n:   # einops.rearrange in the reference fails the same way
            raise ValueError(f"n={n} is not a perfect square")
        prepared = None
    stats = torch.zeros(STATS_LEN, dtype=torch.int64, device=dev)
    if torch.is_grad_enabled() and weight.requires_grad:
        out = _Decode.apply(weight, idx, form, prepared, stats)
    else:
        with torch.no_grad():
            out = _Decode.apply(weight, idx, form, prepared, stats)
    if check_indices and int(stats[_lib.STAT_BAD_INDEX].item()) != 0:
        raise IndexError("index out of range in self")
    return out
