"""Token-sharded quantiser step for N GPUs of one box (one process per GPU, torch.distributed).

Rows (tokens) are independent, so each rank quantises its own shard against the replicated codebook and
forward / encode / decode need no collective.  The backward has exactly one exchange: a single
all-reduce(sum) of one packed **int64** buffer

    [ codebook-gradient segment sums (K*D, fixed point 2^-30) | non-finite counts per code (K)
    | code-usage histogram (K) | loss partial (fixed point 2^-24) | non-finite loss partials ]

Integer sums are exact and order-free, so the all-reduced result -- and therefore grad_weight, the
histogram and the loss -- is bit-identical to the single-GPU result on the concatenated batch, whatever
the world size or NCCL algorithm.  This equals the reference's DDP semantics (mean over ranks of
per-rank mean-loss gradients, trainers/vitgqgan.py:184 + trainers/utils/base_trainer.py:29-33) when
shards are equal, because every rank normalises by the global element count.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import STAT_LOSS_FIXED, STAT_NONFINITE, STATS_LEN


class PackedReduce:
    """Layout + all-reduce of the packed int64 buffer.  Device-agnostic (NCCL on GPU, gloo in CPU tests)."""

    def __init__(self, K: int, D: int):
        self.K, self.D = K, D
        self.seg_len = K * D + K
        self.length = self.seg_len + K + 2

    def allocate(self, device) -> torch.Tensor:
        return torch.empty(self.length, dtype=torch.int64, device=device)

    def seg(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[: self.seg_len]

    def hist(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[self.seg_len: self.seg_len + self.K]

    def loss_pair(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[self.seg_len + self.K:]

    def fill_side_channels(self, buf: torch.Tensor, hist_i32: torch.Tensor, stats: torch.Tensor) -> None:
        self.hist(buf).copy_(hist_i32)
        self.loss_pair(buf).copy_(stats[[STAT_LOSS_FIXED, STAT_NONFINITE]])

    def all_reduce(self, buf: torch.Tensor, group=None) -> torch.Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return buf

    def stats_from(self, buf: torch.Tensor) -> torch.Tensor:
        st = torch.zeros(STATS_LEN, dtype=torch.int64, device=buf.device)
        st[STAT_LOSS_FIXED] = self.loss_pair(buf)[0]
        st[STAT_NONFINITE] = self.loss_pair(buf)[1]
        return st


def shard_batch(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Rank r owns images [r*B/W, (r+1)*B/W); B must divide evenly."""
    b = x.shape[0]
    if b % world_size != 0:
        raise ValueError(f"batch {b} does not split evenly over {world_size} ranks")
    per = b // world_size
    return x[rank * per:(rank + 1) * per]


class ShardedQuantiser:
    """fwd + bwd of one shard through the C ABI, with the packed all-reduce when world_size > 1.

    ``step(z_local, upstream_local, weight)`` returns a dict with z_q, indices, loss (global), grad_z,
    grad_weight (global, identical on every rank), histogram (global, int64) and stats.
    """

    def __init__(self, form: str = "vit", beta: float = 0.25, world_size: Optional[int] = None,
                 exact_scan: bool = False, group=None):
        self.form, self.beta, self.group, self.exact_scan = form, float(beta), group, exact_scan
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size
        self._plans: Dict[tuple, Dict[str, torch.Tensor]] = {}

    @staticmethod
    def uses_tensor_cores(T: int, K: int, D: int) -> bool:
        return bool(_lib.load().vq_uses_tensor_cores(T, K, D))

    def _plan(self, z: torch.Tensor, K: int, D: int, T: int, flags: int) -> Dict[str, torch.Tensor]:
        """Buffers of one (shape, device) are allocated once and reused by every later step: the step
        itself then costs a handful of ctypes calls and no allocator traffic.  Outputs of step() are
        views of these buffers and are overwritten by the next step() with the same shape."""
        from .functional import _scratch
        dev = z.device
        key = (dev, tuple(z.shape), K, D, flags)
        plan = self._plans.get(key)
        if plan is None:
            pack = PackedReduce(K, D)
            fws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, flags)
            bws_bytes = _lib.size_query("vq_backward_workspace_bytes", T, K, D)
            cb_bytes = _lib.size_query("vq_codebook_bytes", K, D)
            plan = {
                "pack": pack, "fws_bytes": fws_bytes, "bws_bytes": bws_bytes, "cb_bytes": cb_bytes,
                "cb": _scratch(cb_bytes, dev), "fws": _scratch(fws_bytes, dev), "bws": _scratch(bws_bytes, dev),
                "z_q": torch.empty_like(z), "grad_z": torch.empty_like(z),
                "idx": torch.empty(T, dtype=torch.int64, device=dev),
                "loss": torch.empty(1, dtype=torch.float32, device=dev),
                "hist": torch.empty(K, dtype=torch.int32, device=dev),
                "stats": torch.empty(STATS_LEN, dtype=torch.int64, device=dev),
                "zn": torch.empty(T, D, dtype=torch.float32, device=dev),
                "denom": torch.empty(T, dtype=torch.float32, device=dev),
                "grad_w": torch.empty(K, D, dtype=torch.float32, device=dev),
                "buf": pack.allocate(dev),
            }
            self._plans[key] = plan
        return plan

    def step(self, z: torch.Tensor, upstream: Optional[torch.Tensor], weight: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .functional import FORMS, LAYOUT_NCHW, LAYOUT_TOKEN_MAJOR, _ptr, _require_cuda, _stream, _token_geometry
        _require_cuda(z, "z")
        lib = _lib.load()
        dev = z.device
        form_id = FORMS[self.form]
        layout = LAYOUT_TOKEN_MAJOR if self.form == "vit" else LAYOUT_NCHW
        z = z.contiguous()
        K, D = weight.shape
        T, hw = _token_geometry(z, layout, D)
        n_total = max(T * D * self.world_size, 1)
        flags = _lib.FLAG_EXACT_SCAN if self.exact_scan else 0
        p = self._plan(z, K, D, T, flags)
        pack, buf = p["pack"], p["buf"]
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        up = None if upstream is None else upstream.contiguous()
        cb, seg = _ptr(p["cb"]), _ptr(buf)          # the segment sums are the head of the packed buffer
        with torch.cuda.device(dev):
            s = _stream(dev)
            _lib.check(lib.vq_codebook_prepare(_ptr(w), K, D, cb, p["cb_bytes"], s))
            _lib.check(lib.vq_forward(_ptr(z), layout, T, hw, cb, K, D, form_id, self.beta, flags, n_total,
                                      _ptr(p["z_q"]), _ptr(p["idx"]), None, _ptr(p["hist"]), _ptr(p["stats"]),
                                      _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["fws"]), p["fws_bytes"], s))
            _lib.check(lib.vq_backward_tokens(_ptr(up), layout, T, hw, _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["idx"]),
                                              _ptr(p["hist"]), cb, K, D, form_id, self.beta, None, n_total, _ptr(p["grad_z"]), seg,
                                              _ptr(p["bws"]), p["bws_bytes"], s))
            if self.world_size > 1:
                pack.fill_side_channels(buf, p["hist"], p["stats"])
                pack.all_reduce(buf, self.group)
                red_stats = pack.stats_from(buf)
                hist_out = pack.hist(buf)
            else:
                red_stats, hist_out = p["stats"], p["hist"]
            _lib.check(lib.vq_loss_finalize(_ptr(red_stats), n_total, form_id, self.beta, _ptr(p["loss"]), s))
            _lib.check(lib.vq_backward_codebook(seg, cb, K, D, form_id, self.beta, None, n_total, _ptr(p["grad_w"]), s))
        return {"z_q": p["z_q"], "indices": p["idx"], "loss": p["loss"].view(()), "grad_z": p["grad_z"],
                "grad_weight": p["grad_w"], "histogram": hist_out, "stats": p["stats"]}
