"""First consumer of the quantiser's tokens.

``masked_token_embeddings`` fuses what MaskGIT / Muse do right behind ``encode_imgs`` -- the mask fill
(/root/reference/models/muse.py:149-150, models/maskgit.py:131-132) and the token-embedding lookup plus
positional encoding (models/muse.py:90-91, models/maskgit.py:80-81) -- into one kernel of libvq_b200.so.
``causal_token_embeddings`` is the autoregressive counterpart (Parti: shift right, start token, sinusoidal rows).
Both are differentiable with respect to the embedding table through the library's deterministic embedding backward.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .functional import _ptr, _require_cuda, _stream, as_tokens, token_bits


class _TokenEmbed(torch.autograd.Function):
    """(table, pos, start | tokens, mask) -> embeds [, ids, labels]; backward: vq_embedding_backward for the table, the
    batch sum for a trainable positional table, the batch sum of position 0 for the start token."""

    @staticmethod
    def forward(ctx, table, pos, start, tok, m, mask_token_id, ignore_index, causal, stats):
        dev = tok.device
        b, n = tok.shape
        V, dim = table.shape
        w = table.detach().contiguous()
        p = None if pos is None else pos.detach().reshape(-1, dim).float().contiguous()
        embeds = torch.empty(b, n, dim, dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            if causal:
                st = start.detach().reshape(dim).float().contiguous()
                _lib.check(lib.vq_token_embed_causal_tokens(_ptr(tok), token_bits(tok.dtype), b * n, n, _ptr(w), V, dim, _ptr(p),
                                                            _ptr(st), _ptr(embeds), _ptr(stats), _stream(dev)))
                ids = labels = tok.new_empty(0)        # the labels of the causal form are the tokens themselves (the caller has them)
            else:
                ids = torch.empty(b, n, dtype=torch.int64, device=dev)
                labels = torch.empty(b, n, dtype=torch.int64, device=dev)
                _lib.check(lib.vq_token_embed_tokens(_ptr(tok), token_bits(tok.dtype), _ptr(m), b * n, n, int(mask_token_id),
                                                     int(ignore_index), _ptr(w), V, dim, _ptr(p), _ptr(embeds), _ptr(ids),
                                                     _ptr(labels), _ptr(stats), _stream(dev)))
        ctx.causal, ctx.shape = causal, (V, dim, b, n)
        ctx.pos_shape = None if pos is None else tuple(pos.shape)
        ctx.save_for_backward(tok if causal else ids)
        ctx.mark_non_differentiable(ids, labels)
        return embeds, ids, labels

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g, _g_ids, _g_labels):
        from .functional import _embedding_backward, convert_tokens
        (ids,) = ctx.saved_tensors
        ids = convert_tokens(ids, torch.int64)
        V, dim, b, n = ctx.shape
        g = g.contiguous().float()
        g_table = g_pos = g_start = None
        if ctx.needs_input_grad[0]:
            if ctx.causal:      # ids[b, i - 1] pairs with the gradient row (b, i): the last token of a sequence feeds nothing
                g_table = _embedding_backward(ids[:, :-1].contiguous(), g, V, dim, ids_per_seq=n - 1, rows_per_seq=n, row_shift=1)
            else:
                g_table = _embedding_backward(ids, g, V, dim)
        if ctx.needs_input_grad[1] and ctx.pos_shape is not None:
            gp = g.sum(0)                              # the broadcast `+= pos_enc` of the reference: autograd's own batch sum
            if ctx.causal:
                gp = torch.cat([gp[1:], gp.new_zeros(1, dim)])       # row i of the positional table met position i + 1
            g_pos = gp.reshape(ctx.pos_shape)
        if ctx.needs_input_grad[2] and ctx.causal:
            g_start = g[:, 0].sum(0)
        return g_table, g_pos, g_start, None, None, None, None, None, None


def _checked(tokens, table):
    _require_cuda(tokens, "tokens")
    _require_cuda(table, "the embedding table")
    if tokens.dim() != 2:
        raise ValueError("tokens must be (b, n)")
    if table.dtype != torch.float32 or table.dim() != 2 or table.shape[1] % 4 != 0:
        raise TypeError("table must be a (vocab, dim) float32 tensor with dim a multiple of 4")


def masked_token_embeddings(tokens: torch.Tensor, mask: Optional[torch.Tensor], mask_token_id: int, table: torch.Tensor,
                            pos_enc: Optional[torch.Tensor] = None, ignore_index: int = -1, check_indices: bool = True
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (embeds (b, n, dim) fp32, input_ids (b, n) int64, labels (b, n) int64).

    ``tokens``: (b, n) integer (int64, or the narrow wire formats int32 / uint16 of ``encode_indices``); ``mask``: (b, n) bool or None; ``table``: (vocab, dim) fp32 embedding weight;
    ``pos_enc``: (1, n, dim) or (n, dim) fp32 or None.  Differentiable with respect to ``table`` and ``pos_enc`` (a
    training step of MaskGIT / Muse).  Out-of-range ids raise IndexError like ``nn.Embedding`` on the CPU (one host
    sync; ``check_indices=False`` skips it)."""
    _checked(tokens, table)
    dev = tokens.device
    b, n = tokens.shape
    dim = table.shape[1]
    tok = as_tokens(tokens)
    m = None
    if mask is not None:
        if mask.shape != tokens.shape:
            raise ValueError("mask must have the shape of tokens")
        m = mask.to(torch.uint8).contiguous()
    if pos_enc is not None and pos_enc.numel() != n * dim:
        raise ValueError(f"pos_enc holds {pos_enc.numel() // dim} positions, tokens have {n}")
    stats = torch.zeros(_lib.STATS_LEN, dtype=torch.int64, device=dev)
    embeds, ids, labels = _TokenEmbed.apply(table, pos_enc, None, tok, m, mask_token_id, ignore_index, False, stats)
    if check_indices and int(stats[_lib.STAT_BAD_INDEX].item()) != 0:
        raise IndexError("index out of range in self")
    return embeds, ids, labels


def causal_token_embeddings(tokens: torch.Tensor, table: torch.Tensor, pos_enc: Optional[torch.Tensor],
                            start_token: torch.Tensor, check_indices: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Parti's decoder input (/root/reference/models/parti.py:98-106) in one kernel:
    ``cat(start_token, token_emb(tokens[:, :-1]) + pe[:n-1])`` -> (embeds (b, n, dim) fp32, labels = tokens (b, n)).

    ``pos_enc``: the first n - 1 (or more) rows of the sinusoidal table (models/positional_encoding.py:27-32), shape
    (>= n - 1, dim), or None; the dropout the reference applies behind it is left to the caller.  Differentiable with
    respect to ``table`` and ``start_token``."""
    _checked(tokens, table)
    dev = tokens.device
    b, n = tokens.shape
    dim = table.shape[1]
    tok = as_tokens(tokens)
    pos = None
    if pos_enc is not None:
        pos = pos_enc.reshape(-1, dim)
        if pos.shape[0] < n - 1:
            raise ValueError(f"pos_enc holds {pos.shape[0]} positions, {n - 1} are needed")
        pos = pos[:n].contiguous() if pos.shape[0] >= n else torch.cat([pos, pos.new_zeros(n - pos.shape[0], dim)])
    if start_token.numel() != dim:
        raise ValueError("start_token must hold dim values")
    stats = torch.zeros(_lib.STATS_LEN, dtype=torch.int64, device=dev)
    embeds, _ids, _labels = _TokenEmbed.apply(table, pos, start_token, tok, None, 0, 0, True, stats)
    if check_indices and int(stats[_lib.STAT_BAD_INDEX].item()) != 0:
        raise IndexError("index out of range in self")
    return embeds, tok
