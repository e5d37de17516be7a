"""Token-sharded quantiser step for N GPUs of one box (one process per GPU, torch.distributed).

Rows (tokens) are independent, so each rank quantises its own shard against the replicated codebook and
forward / encode / decode need no collective.  The backward has exactly one exchange: the sum over ranks of

    [ codebook-gradient segment sums (K*D, fixed point 2^-30) | non-finite counts per code (K)
    | code-usage histogram (K) | loss partial (fixed point 2^-24) | non-finite loss partials ]

all of them **integers**: sums are exact and order-free, so the result -- and therefore grad_weight, the
histogram and the loss -- is bit-identical to the single-GPU result on the concatenated batch, whatever
the world size or the order of the additions.  This equals the reference's DDP semantics (mean over ranks of
per-rank mean-loss gradients, trainers/vitgqgan.py:184 + trainers/utils/base_trainer.py:29-33) when
shards are equal, because every rank normalises by the global element count.

Two implementations of the exchange:

* ``exchange="peer"`` (default on CUDA): every rank's forward writes its partials into an exchange buffer that
  all peers map through CUDA IPC; ONE kernel per rank (`vq_backward_codebook_sharded`) publishes a flag, waits
  for the peers' flags, pulls their partials over NVLink, adds them and applies the codebook gradient.  The
  partials depend on neither the upstream gradient nor d(loss), so the kernel is launched on a side stream
  right after the forward and overlaps the token backward.
* ``exchange="collective"``: one `all_reduce(SUM)` of the packed int64 buffer (NCCL on GPUs; gloo in the CPU
  tests of the host logic), then the single-GPU codebook-gradient kernel.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

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
        assert STAT_NONFINITE == STAT_LOSS_FIXED + 2
        self.loss_pair(buf).copy_(stats[STAT_LOSS_FIXED:STAT_NONFINITE + 1:2])

    def all_reduce(self, buf: torch.Tensor, group=None) -> torch.Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return buf

    def stats_from(self, buf: torch.Tensor) -> torch.Tensor:
        st = torch.zeros(STATS_LEN, dtype=torch.int64, device=buf.device)
        st[STAT_LOSS_FIXED:STAT_NONFINITE + 1:2] = self.loss_pair(buf)
        return st


def shard_batch(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Rank r owns images [r*B/W, (r+1)*B/W); B must divide evenly."""
    b = x.shape[0]
    if b % world_size != 0:
        raise ValueError(f"batch {b} does not split evenly over {world_size} ranks")
    per = b // world_size
    return x[rank * per:(rank + 1) * per]


class PeerTimeoutError(RuntimeError):
    """The fused exchange gave up waiting for a peer (or was launched out of step): the codebook gradient and loss of
    that step are NaN on this rank and the exchange stays dead until every rank calls ``resync()``."""


def _unmap_exchange(ptrs, rank, device_index):
    """Best-effort clean-up when a PeerExchange is dropped without the collective close(): unmap the peers' buffers and
    free the own one, no barriers (a peer that still reads this rank's buffer has a bug of its own by then)."""
    try:
        lib = _lib.load()
        with torch.cuda.device(device_index):
            torch.cuda.synchronize()
            for r, p in enumerate(ptrs):
                if p and r != rank:
                    lib.vq_peer_close(p)
            if ptrs and ptrs[rank]:
                lib.vq_peer_free(ptrs[rank])
    except Exception:       # interpreter shutdown: the driver reclaims everything anyway
        pass


class PeerExchange:
    """Exchange buffers of all ranks of `group`, mapped into this process (CUDA IPC over NVLink).

    `torch.distributed` only carries the 64-byte IPC handles at set-up; the data path is the library's own
    kernel reading peer memory.

    ``timeout_s``: how long the exchange kernel waits for a peer's step before it gives up (None: the library default
    of 10 minutes).  Giving up is fatal and loud: NaN results, ``abort_flag`` raised (a pinned host int the kernel
    writes; ``raise_if_aborted()`` reads it without a device synchronisation), every later step fails until the
    collective ``resync()``."""

    def __init__(self, K: int, D: int, device: torch.device, group=None, timeout_s: Optional[float] = None):
        lib = _lib.load()
        self.K, self.D, self.device, self.group = K, D, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.PEER_MAX_RANKS:
            raise ValueError(f"peer exchange supports at most {_lib.PEER_MAX_RANKS} ranks")
        self.nbytes = _lib.size_query("vq_exchange_bytes", K, D)
        self.epoch = 0
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        with torch.cuda.device(device):
            _lib.check(lib.vq_peer_alloc(self.nbytes, ctypes.byref(own), handle))
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            self.ptrs: List[int] = []
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs.append(own.value)
                else:
                    p = ctypes.c_void_p()
                    try:
                        _lib.check(lib.vq_peer_open(ctypes.create_string_buffer(handles[r], _lib.IPC_HANDLE_BYTES),
                                                    ctypes.byref(p)))
                    except _lib.VQLibraryError as exc:
                        raise RuntimeError(
                            f"rank {self.rank} cannot map rank {r}'s exchange buffer ({exc}). The peer exchange needs CUDA "
                            "IPC and NVLink / PCIe peer access between all ranks of the group (one box); pass "
                            "exchange='collective' to use one NCCL all-reduce instead") from exc
                    self.ptrs.append(p.value)
        self.ptr_array = (ctypes.c_void_p * self.world)(*self.ptrs)
        self._slots = [self._slot_pointers(s) for s in (0, 1)]
        # pinned host memory is device-visible at the same address (unified addressing): the kernel raises this flag
        self.abort_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        timeout_ms = 0 if timeout_s is None else max(1, int(timeout_s * 1000))
        with torch.cuda.device(device):
            _lib.check(lib.vq_peer_configure(self.ptrs[self.rank], timeout_ms, self.abort_flag.data_ptr(),
                                             torch.cuda.current_stream(device).cuda_stream))
        import weakref
        self._ptr_holder = list(self.ptrs)          # what the finalizer unmaps if close() never ran
        self._finalizer = weakref.finalize(self, _unmap_exchange, self._ptr_holder, self.rank,
                                           device.index if device.index is not None else torch.cuda.current_device())
        dist.barrier(group=group)          # every rank mapped every buffer before the first kernel publishes into it

    def aborted(self) -> bool:
        return bool(self.abort_flag[0].item())      # a host read of pinned memory: no device synchronisation

    def raise_if_aborted(self) -> None:
        if self.aborted():
            raise PeerTimeoutError(
                f"rank {self.rank}: the peer exchange of step {self.epoch} gave up (a peer did not publish its partials in "
                "time, or the launches went out of step); grad_weight / loss of that step are NaN.  Call resync() on "
                "every rank (collective) to restart the exchange, or fall back to exchange='collective'.")

    def resync(self) -> None:
        """Collective: restart the exchange after a time-out (every rank, no step in flight)."""
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            _lib.check(lib.vq_peer_resync(self.ptrs[self.rank], torch.cuda.current_stream(self.device).cuda_stream))
            torch.cuda.synchronize(self.device)
        self.abort_flag.zero_()
        self.epoch = 0
        dist.barrier(group=self.group)

    def _slot_pointers(self, slot: int):
        seg, stats, hist = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.load().vq_exchange_slot(self.ptrs[self.rank], self.K, self.D, slot, ctypes.byref(seg),
                                                ctypes.byref(stats), ctypes.byref(hist)))
        return seg.value, stats.value, hist.value

    def next_step(self):
        """-> (slot, epoch, seg_ptr, stats_ptr, hist_ptr) of the step about to run."""
        self.epoch += 1
        slot = (self.epoch - 1) & 1
        return (slot, self.epoch) + self._slots[slot]

    def close(self) -> None:
        if not self.ptrs:
            return
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)     # nobody still reads a buffer that is about to be unmapped / freed
        with torch.cuda.device(self.device):
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    _lib.check(lib.vq_peer_close(p))
            dist.barrier(group=self.group)
            _lib.check(lib.vq_peer_free(self.ptrs[self.rank]))
        self.ptrs = []
        self._ptr_holder.clear()           # nothing left for the finalizer
        self._finalizer.detach()


class ShardedQuantiser:
    """fwd + bwd of one shard through the C ABI, with the one exchange of the backward when world_size > 1.

    ``step(z_local, upstream_local, weight)`` returns a dict with z_q, indices, loss (global), grad_z,
    grad_weight (global, identical on every rank), histogram (global) and stats.
    """

    def __init__(self, form: str = "vit", beta: float = 0.25, world_size: Optional[int] = None,
                 exact_scan: bool = False, group=None, exchange: str = "peer", graphs: bool = False,
                 peer_timeout_s: Optional[float] = None):
        self.form, self.beta, self.group, self.exact_scan = form, float(beta), group, exact_scan
        self.peer_timeout_s = peer_timeout_s      # None: the library default (10 minutes); see PeerExchange
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if exchange not in ("peer", "collective"):
            raise ValueError("exchange must be 'peer' or 'collective'")
        self.world_size, self.exchange = world_size, exchange
        # graphs: the launches of a step are captured once per (input buffers, weight buffer) in a CUDA graph and
        # replayed afterwards (one graph launch instead of five kernel launches; the step is capture-safe: no
        # allocation, no host synchronisation, device-side counters only).  `step(..., eager=True)` bypasses it.
        self.graphs = graphs
        self._graphs: Dict[tuple, tuple] = {}
        self._warmed = set()
        self.max_graphs = 64
        self.graph_kernel_launches = 0        # kernels run through graph replays (the library only counts eager launches)
        self._plans: Dict[tuple, Dict[str, object]] = {}
        import os
        self.trace = [] if os.environ.get("VQ_STEP_EVENTS") else None     # debugging: per-step CUDA events of the peer path

    @staticmethod
    def uses_tensor_cores(T: int, K: int, D: int) -> bool:
        return bool(_lib.load().vq_uses_tensor_cores(T, K, D))

    def _plan(self, z: torch.Tensor, K: int, D: int, T: int, flags: int) -> Dict[str, object]:
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
                "hist_total": torch.empty(K, dtype=torch.int64, device=dev),
                "stats": torch.empty(STATS_LEN, dtype=torch.int64, device=dev),
                "zn": torch.empty(T, D, dtype=torch.float32, device=dev),
                "denom": torch.empty(T, dtype=torch.float32, device=dev),
                "grad_w": torch.empty(K, D, dtype=torch.float32, device=dev),
                "buf": pack.allocate(dev),
            }
            if self.world_size > 1 and self.exchange == "peer":
                plan["peer"] = PeerExchange(K, D, dev, self.group, timeout_s=self.peer_timeout_s)
                plan["side"] = torch.cuda.Stream(dev)
                plan["ev_fwd"], plan["ev_x"] = torch.cuda.Event(), torch.cuda.Event()
            self._plans[key] = plan
        return plan

    def close(self) -> None:
        """Unmap / free the peer exchange buffers (collective: every rank must call it)."""
        self._graphs.clear()
        for plan in self._plans.values():
            if "peer" in plan:
                plan["peer"].close()
        self._plans.clear()

    def resync(self) -> None:
        """Collective: restart the peer exchange of every plan after a PeerTimeoutError."""
        self._graphs.clear()
        self._warmed.clear()
        for plan in self._plans.values():
            if "peer" in plan:
                plan["peer"].resync()

    def step(self, z: torch.Tensor, upstream: Optional[torch.Tensor], weight: torch.Tensor,
             eager: bool = False) -> Dict[str, torch.Tensor]:
        # a time-out of an earlier step's exchange is reported here, at the latest one step later (the kernel raises a
        # pinned host flag; reading it does not synchronise with the device)
        for plan in self._plans.values():
            peer = plan.get("peer")
            if peer is not None:
                peer.raise_if_aborted()
        # a graph is tied to buffer addresses: only caller-owned contiguous buffers qualify (a .contiguous() copy would
        # be a fresh address on every call), and the cache is bounded
        if (not self.graphs or eager or not self._graph_capable() or not z.is_contiguous() or not weight.is_contiguous()
                or (upstream is not None and not upstream.is_contiguous())):
            return self._step(z, upstream, weight)
        key = (z.data_ptr(), tuple(z.shape), None if upstream is None else upstream.data_ptr(), weight.data_ptr(),
               tuple(weight.shape), self._graph_phase(z, weight))
        hit = self._graphs.get(key)
        if hit is None:
            if key not in self._warmed or len(self._graphs) >= self.max_graphs:
                self._warmed.add(key)                     # first use: a real eager step (allocations, kernel attributes)
                if len(self._warmed) > 16 * self.max_graphs:
                    self._warmed.clear()
                return self._step(z, upstream, weight)
            torch.cuda.synchronize(z.device)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.load().vq_kernel_launches()
            with torch.cuda.graph(g):
                out = self._step(z, upstream, weight)
            n_kernels = _lib.load().vq_kernel_launches() - n0
            self._graph_rewind(z, weight)                 # the capture advanced the host-side step count, nothing ran
            hit = (g, out, (z, upstream, weight), n_kernels)   # the graph reads these buffers: keep them alive
            self._graphs[key] = hit
        hit[0].replay()
        self.graph_kernel_launches += hit[3]
        self._graph_advance(z, weight)
        return hit[1]

    # Graphs are offered where a step has no per-step host state in its kernel arguments: always on one GPU; with the
    # peer exchange for token-major rows (one launch for the backward, step number on the device, two graphs per input
    # buffer: one per exchange slot).  The NCCL exchange and the two-stream NCHW backward stay eager.
    def _graph_capable(self) -> bool:
        return self.world_size == 1 or (self.exchange == "peer" and self.form == "vit")

    def _peer_of(self, z, weight) -> Optional[PeerExchange]:
        for key, plan in self._plans.items():
            if key[1] == tuple(z.shape) and key[2:4] == tuple(weight.shape) and "peer" in plan:
                return plan["peer"]
        return None

    def _graph_phase(self, z, weight) -> int:
        peer = self._peer_of(z, weight) if self.world_size > 1 else None
        return 0 if peer is None else peer.epoch & 1

    def _graph_rewind(self, z, weight) -> None:
        peer = self._peer_of(z, weight) if self.world_size > 1 else None
        if peer is not None:
            peer.epoch -= 1

    def _graph_advance(self, z, weight) -> None:
        peer = self._peer_of(z, weight) if self.world_size > 1 else None
        if peer is not None:
            peer.epoch += 1

    def _step(self, z: torch.Tensor, upstream: Optional[torch.Tensor], weight: torch.Tensor) -> Dict[str, torch.Tensor]:
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
        cb = _ptr(p["cb"])
        peer: Optional[PeerExchange] = p.get("peer")
        with torch.cuda.device(dev):
            s = _stream(dev)
            if peer is not None:
                slot, epoch, seg, stats_ptr, hist_ptr = peer.next_step()
            else:
                seg, stats_ptr, hist_ptr = _ptr(buf), _ptr(p["stats"]), _ptr(p["hist"])   # seg sums head the packed buffer
            # the weights of a training step changed: the forward prepares the codebook itself (same launch as the rows)
            _lib.check(lib.vq_forward(_ptr(z), layout, T, hw, _ptr(w), cb, K, D, form_id, self.beta, flags, n_total,
                                      _ptr(p["z_q"]), _ptr(p["idx"]), None, hist_ptr, stats_ptr,
                                      _ptr(p["zn"]), _ptr(p["denom"]), seg, _ptr(p["fws"]), p["fws_bytes"], s))
            if peer is not None:
                # The exchange needs nothing of the backward, so the two run side by side: in one launch for token-major
                # rows, else the exchange on the caller's stream and the token backward on a side stream.
                tr = None
                if self.trace is not None:
                    tr = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                    self.trace.append(tr)
                    tr[1].record()
                if layout == LAYOUT_TOKEN_MAJOR:
                    # one launch: the exchange + codebook gradient on the first blocks, grad_z on the others
                    # (a captured launch carries no host-side step number: the kernel reads the device's)
                    _lib.check(lib.vq_backward_sharded(peer.ptr_array, peer.world, peer.rank, slot,
                                                       0 if torch.cuda.is_current_stream_capturing() else epoch, _ptr(up), T,
                                                       _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["idx"]), cb, K, D, form_id,
                                                       self.beta, None, n_total, _ptr(p["grad_z"]), _ptr(p["grad_w"]),
                                                       _ptr(p["hist_total"]), _ptr(p["loss"]), _ptr(p["stats"]), s))
                else:
                    side = p["side"]
                    p["ev_fwd"].record()
                    _lib.check(lib.vq_backward_codebook_sharded(peer.ptr_array, peer.world, peer.rank, slot, epoch, cb, K, D,
                                                                form_id, self.beta, None, n_total, _ptr(p["grad_w"]),
                                                                _ptr(p["hist_total"]), _ptr(p["loss"]), _ptr(p["stats"]), s))
                    side.wait_event(p["ev_fwd"])
                    _lib.check(lib.vq_backward_tokens(_ptr(up), layout, T, hw, _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["idx"]),
                                                      None, cb, K, D, form_id, self.beta, None, n_total, _ptr(p["grad_z"]), None,
                                                      _ptr(p["bws"]), p["bws_bytes"], side.cuda_stream))
                    p["ev_x"].record(side)
                    torch.cuda.current_stream(dev).wait_event(p["ev_x"])
                if tr is not None:
                    tr[2].record(); tr[3].record(); tr[4].record()
                hist_out = p["hist_total"]
            elif self.world_size == 1:
                # one launch: grad_z, grad_weight and the loss
                _lib.check(lib.vq_backward(_ptr(up), layout, T, hw, _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["idx"]), cb, K, D,
                                           form_id, self.beta, None, n_total, seg, stats_ptr, _ptr(p["grad_z"]), _ptr(p["grad_w"]),
                                           _ptr(p["loss"]), _ptr(p["bws"]), p["bws_bytes"], s))
                hist_out = p["hist"]
            else:
                _lib.check(lib.vq_backward_tokens(_ptr(up), layout, T, hw, _ptr(p["zn"]), _ptr(p["denom"]), _ptr(p["idx"]),
                                                  None, cb, K, D, form_id, self.beta, None, n_total, _ptr(p["grad_z"]), None,
                                                  _ptr(p["bws"]), p["bws_bytes"], s))
                pack.fill_side_channels(buf, p["hist"], p["stats"])
                pack.all_reduce(buf, self.group)
                red_stats = pack.stats_from(buf)
                hist_out = pack.hist(buf)
                _lib.check(lib.vq_backward_codebook(seg, cb, K, D, form_id, self.beta, None, n_total, _ptr(p["grad_w"]),
                                                    _ptr(red_stats), _ptr(p["loss"]), s))
        return {"z_q": p["z_q"], "indices": p["idx"], "loss": p["loss"].view(()), "grad_z": p["grad_z"],
                "grad_weight": p["grad_w"], "histogram": hist_out, "stats": p["stats"]}
