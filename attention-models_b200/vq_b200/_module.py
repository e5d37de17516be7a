"""Shared base of the two drop-in ``Codebook`` modules."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as F_vq


class _CodebookBase(nn.Module):
    """Keeps the reference's public surface: ``codebook_size``, ``codebook_dim``,
    ``beta``, an ``embedding`` nn.Embedding (state_dict key ``embedding.weight``), ``forward(z) ->
    (z_q, indices, loss)`` and ``indices_to_embeddings(indices)``.

    Extras that the reference does not have (north_star additions): ``last_histogram`` (code usage of
    the last forward, int32 K), ``last_stats`` (device int64 counters: near-tie rows, ...),
    ``encode(z)`` (indices-only fast path used by ``encode_imgs`` callers).
    """

    form: str = "vit"

    def __init__(self, codebook_size: int, codebook_dim: int, beta: float):
        super().__init__()
        self.codebook_size = codebook_size
        self.codebook_dim = codebook_dim
        self.beta = beta
        self.embedding = nn.Embedding(self.codebook_size, self.codebook_dim)
        self.exact_scan = False          # True forces the exhaustive fp32 search (no tensor cores)
        self.sorted_segments = False     # True: codebook-gradient sums bucketed by code in the backward (skew-insensitive)
                                         # instead of integer reductions from the forward's finish pass; same bits
        self.last_histogram: Optional[torch.Tensor] = None
        self.last_stats: Optional[torch.Tensor] = None
        self._prepared: Optional[F_vq.PreparedCodebook] = None
        self._table = None               # projected codes behind post_quant (decode_projected)

    # The prepared codebook (unit codes, fp16 copy, ...) is a cache keyed on the weight's storage and version counter: a
    # frozen tokeniser (MaskGIT / Muse / Parti) prepares once.  A forward that can train the codebook never trusts the
    # cache (the refill rides in its first launch for free).  What the version counter cannot see is an edit through
    # `embedding.weight.data` (the reference's own init idiom, `weights_init`, codebook restarts) made while the module
    # is used for inference only: call invalidate_codebook() after such an edit.  load_state_dict() and .to() / .cuda()
    # invalidate by themselves.
    def invalidate_codebook(self) -> None:
        self._prepared = None
        self._table = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._prepared = self._table = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._prepared = self._table = None
        return super()._apply(fn, *args, **kwargs)

    def _prepared_codebook(self) -> F_vq.PreparedCodebook:
        w = self.embedding.weight
        raw = self.form == "l2"
        if self._prepared is None or not self._prepared.matches(w, raw):
            self._prepared = F_vq.prepare_codebook(w, raw)
        return self._prepared

    def _quantise(self, z: torch.Tensor):
        # a stale prepared codebook of the right size is refilled by the forward itself (training: the weights changed)
        w = self.embedding.weight
        if self._prepared is None or not self._prepared.fits(w):
            self._prepared = F_vq.prepare_codebook(w, self.form == "l2")
        trainable = torch.is_grad_enabled() and w.requires_grad
        z_q, flat_idx, loss, hist, stats = F_vq.quantise(z, self.embedding.weight, self.form, self.beta,
                                                         prepared=self._prepared,
                                                         exact_scan=self.exact_scan,
                                                         sorted_segments=self.sorted_segments,
                                                         always_refresh=trainable)
        self.last_histogram, self.last_stats = hist, stats
        return z_q, flat_idx, loss

    def encode(self, z: torch.Tensor, index_dtype: torch.dtype = torch.int64) -> torch.Tensor:
        """Flat indices only (what ``encode_imgs`` keeps; reference vitvqgan.py:204-210): int64 like the reference, or
        the narrow wire formats torch.int32 / torch.uint16 that ``indices_to_embeddings`` and the token consumers read."""
        return F_vq.encode_indices(z, self.embedding.weight, self.form, prepared=self._prepared_codebook(),
                                   exact_scan=self.exact_scan, index_dtype=index_dtype)

    def indices_to_embeddings(self, indices: torch.Tensor) -> torch.Tensor:
        prepared = self._prepared_codebook() if self.form == "vit" else None
        return F_vq.indices_to_embeddings(indices, self.embedding.weight, self.form, prepared=prepared)

    # ---- pre_quant / post_quant fused with the quantiser (vq_b200/projected.py) ----------
    def supports_fused_pre_quant(self, pre_quant: nn.Module) -> bool:
        """Whether ``forward_projected`` / ``encode_projected`` cover this ``pre_quant``: the ViT form behind an fp32
        ``nn.Linear(C, 32)`` with C a multiple of 64 up to 768 (``vq_prequant_supported``)."""
        from . import projected
        return (self.form == "vit" and isinstance(pre_quant, nn.Linear) and pre_quant.weight.dtype == torch.float32
                and pre_quant.out_features == self.codebook_dim
                and projected.prequant_supported(pre_quant.in_features, self.codebook_dim))

    def forward_projected(self, x: torch.Tensor, pre_quant: nn.Linear):
        """``self(pre_quant(x))`` (models/vitvqgan.py:192-193) with the projection formed inside the token preparation:
        ``(z_q (..., D), indices (...), loss)``; gradients reach ``x``, ``pre_quant`` and the codebook."""
        from . import projected
        w = self.embedding.weight
        if self._prepared is None or not self._prepared.fits(w):
            self._prepared = F_vq.prepare_codebook(w)
        trainable = torch.is_grad_enabled() and w.requires_grad
        z_q, flat_idx, loss, hist, stats = projected.quantise_projected(
            x, pre_quant.weight, pre_quant.bias, w, self.beta, prepared=self._prepared, exact_scan=self.exact_scan,
            always_refresh=trainable)
        self.last_histogram, self.last_stats = hist, stats
        return z_q, flat_idx.view(*x.shape[:-1]), loss

    def encode_projected(self, x: torch.Tensor, pre_quant: nn.Linear, index_dtype: torch.dtype = torch.int64) -> torch.Tensor:
        """Tokens of ``pre_quant(x)`` (models/vitvqgan.py:207-209), indices-only path."""
        from . import projected
        idx = projected.encode_indices_projected(x, pre_quant.weight, pre_quant.bias, self.embedding.weight,
                                                 prepared=self._prepared_codebook(), exact_scan=self.exact_scan,
                                                 index_dtype=index_dtype)
        return idx.view(*x.shape[:-1])

    def decode_projected(self, indices: torch.Tensor, post_quant: nn.Module) -> torch.Tensor:
        """``post_quant(self.indices_to_embeddings(indices))`` (models/vitvqgan.py:199-200, models/vqgan.py:241-242) as
        one gather from the table of projected codes; the table is rebuilt when the codebook or ``post_quant`` change
        (same caveat about edits through ``.data`` as above: ``invalidate_codebook()``).  Inference only (no autograd
        graph): a training step goes through ``forward``."""
        from . import projected
        if self.form not in ("vit", "vqgan"):
            raise ValueError("decode_projected: the reference's two forms only")
        w, wp, bp = self.embedding.weight, post_quant.weight, post_quant.bias
        if isinstance(post_quant, nn.Conv2d) and (post_quant.kernel_size != (1, 1) or post_quant.groups != 1):
            raise ValueError("post_quant must be a 1x1 convolution (models/vqgan.py:228)")
        if self._table is None or not self._table.matches(w, wp, bp):
            prepared = self._prepared_codebook() if self.form == "vit" else None
            self._table = projected.ProjectedTable(w, wp, bp, self.form, prepared)
        return self._table.gather(indices)

    def near_tie_rows(self) -> int:
        """Rows of the last forward whose two best fp32 distances were < 1e-6 relative apart (host sync)."""
        from ._lib import STAT_NEAR_TIE_ROWS
        return 0 if self.last_stats is None else int(self.last_stats[STAT_NEAR_TIE_ROWS].item())
