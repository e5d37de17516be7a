"""First consumer of the quantiser's tokens.

``masked_token_embeddings`` fuses what MaskGIT / Muse do right behind ``encode_imgs`` -- the mask fill
(/root/reference/models/muse.py:149-150, models/maskgit.py:131-132) and the token-embedding lookup plus
positional encoding (models/muse.py:90-91, models/maskgit.py:80-81) -- into one kernel of libvq_b200.so.
Forward only: use it for sampling / frozen embeddings; a training step of the transformer keeps autograd's
``nn.Embedding``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .functional import _ptr, _require_cuda, _stream


@torch.no_grad()
def masked_token_embeddings(tokens: torch.Tensor, mask: Optional[torch.Tensor], mask_token_id: int, table: torch.Tensor,
                            pos_enc: Optional[torch.Tensor] = None, ignore_index: int = -1, check_indices: bool = True
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (embeds (b, n, dim) fp32, input_ids (b, n) int64, labels (b, n) int64).

    ``tokens``: (b, n) integer; ``mask``: (b, n) bool or None; ``table``: (vocab, dim) fp32 embedding weight;
    ``pos_enc``: (1, n, dim) or (n, dim) fp32 or None.  Out-of-range ids raise IndexError like ``nn.Embedding`` on
    the CPU (one host sync; ``check_indices=False`` skips it)."""
    _require_cuda(tokens, "tokens")
    _require_cuda(table, "the embedding table")
    if tokens.dim() != 2:
        raise ValueError("tokens must be (b, n)")
    if table.dtype != torch.float32 or table.dim() != 2 or table.shape[1] % 4 != 0:
        raise TypeError("table must be a (vocab, dim) float32 tensor with dim a multiple of 4")
    dev = tokens.device
    b, n = tokens.shape
    V, dim = table.shape
    tok = tokens.to(torch.int64).contiguous()
    m = None
    if mask is not None:
        if mask.shape != tokens.shape:
            raise ValueError("mask must have the shape of tokens")
        m = mask.to(torch.uint8).contiguous()
    pos = None
    if pos_enc is not None:
        pos = pos_enc.detach().reshape(-1, dim).float().contiguous()
        if pos.shape[0] != n:
            raise ValueError(f"pos_enc holds {pos.shape[0]} positions, tokens have {n}")
    w = table.detach().contiguous()
    embeds = torch.empty(b, n, dim, dtype=torch.float32, device=dev)
    ids = torch.empty(b, n, dtype=torch.int64, device=dev)
    labels = torch.empty(b, n, dtype=torch.int64, device=dev)
    stats = torch.zeros(_lib.STATS_LEN, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().vq_token_embed(_ptr(tok), _ptr(m), b * n, n, int(mask_token_id), int(ignore_index), _ptr(w), V, dim,
                                              _ptr(pos), _ptr(embeds), _ptr(ids), _ptr(labels), _ptr(stats), _stream(dev)))
    if check_indices and int(stats[_lib.STAT_BAD_INDEX].item()) != 0:
        raise IndexError("index out of range in self")
    return embeds, ids, labels
