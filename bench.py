#!/usr/bin/env python
"""bench.py -- VQ codebook quantiser throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config cfg1|cfg2|cfg2fwd|cfg3|cfg4|sweep:T,K,D|cfg3pre] [--scaling weak|strong]

Default workload (config.workload): BASELINE.json configs[2] -- the ViT-VQGAN training-step quantiser, fwd + bwd
(straight-through + codebook gradient), K = 8192 codes x D = 32, batch 256 x 1024 tokens (262 144 tokens), fp32,
synthetic N(0,1) latents and upstream gradients, N(0,1) codebook.  `--scaling weak` (default): every rank quantises
the config's whole batch (262 144 tokens PER GPU) against the replicated codebook; `--scaling strong`: the config's batch
is split over the ranks (cfg3 as BASELINE names it: 256 images over 8 GPUs = 32 768 tokens per GPU).  The only exchange
of the step is the backward's sum of the integer codebook-gradient partials, the usage histogram and the loss partial
(fused peer-memory kernel over NVLink, or one NCCL all-reduce with --exchange collective).

One "step" = one pass of the hot path over one batch:
  mode step       codebook prepare, forward (z_q, indices, loss), backward (grad_z, grad_weight)
  mode encode     encode_imgs: indices only, frozen (prepared) codebook
  mode roundtrip  encode_imgs + decode_indices
Inputs rotate over several resident sets so that every step reads data that is not in L2.

Prints ONE JSON line (rank 0).  See the task contract for the keys; notes:
  value      tokens/s over all ranks with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the host-buffer path (pinned host -> device -> pinned host inside the timed region)
  roofline   the nearest-code search kernel: algorithmic 2*K*D flop/token over its live CUDA-event time; `peak` is the
             measured BURST bf16 peak when the timed region is a sub-second burst at full clocks, the SUSTAINED one
             otherwise; `sustained_leg` repeats the measurement over a >= 2 s run with the clock trace beside it
  hbm_side   the HBM-bound kernels against the measured copy bandwidth, by ALGORITHMIC bytes and by the
             bytes the kernel interfaces actually move
  parity     checks made outside the timed region; at N > 1: cross-rank bit identity of grad_weight / loss / histogram
             and equality with a single-GPU step on the concatenated batch
  cpu_baseline  the oracle port of the reference (torch CPU ops, all host threads) on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "attention-models_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import bench_inputs as bi

BETA = 0.25
UNIT = "tokens/s"
MODE_TAG = {"step": "fwd_bwd", "encode": "encode", "roundtrip": "encode_decode"}
# dram__bytes_read.sum + dram__bytes_write.sum of one k_dist_tc16 launch at cfg3 (ncu --set full, profiles/): the fp16
# token rows (16.8 MB) + the 0.5 MB fp16 codebook; the verdict records are still in L2 when the kernel ends
DRAM_TRAFFIC_FILTER_CFG3 = 17.33e6


def metric_name(cfg):
    return f"vq_tokens_per_sec_{MODE_TAG[cfg['mode']]}_K{cfg['K']}_D{cfg['D']}"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING a timed region.

    In-process NVML polling (a few ms period) because the timed region can be shorter than one
    `nvidia-smi -lms 200` tick; falls back to nvidia-smi if pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self._stop, self._thread, self._h = gpu_index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv, self._h, self.max_mhz = None, None, None

    def _poll(self):
        nv = self._nv
        while not self._stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                except Exception:
                    watts = None
                self.samples.append((float(mhz), int(mask), watts))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        """Starts polling and returns once the first samples are in: the first NVML queries of a process take tens of
        milliseconds (more with 8 ranks initialising at once) and would otherwise stall the host inside the timed region;
        those warm-up samples are dropped."""
        if self._h is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            t0 = time.perf_counter()
            while len(self.samples) < 3 and time.perf_counter() - t0 < 2.0:
                time.sleep(0.001)
            self.samples.clear()
        return self

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=1)
        if not self.samples:
            return self._smi_once()
        sm = [s[0] for s in self.samples]
        watts = [s[2] for s in self.samples if s[2] is not None]
        reasons = sorted({name for _, mask, _ in self.samples for name, bit in self.REASONS if mask & bit})
        return {"sm_mhz": statistics.median(sm), "sm_mhz_min": min(sm), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "power_w_median": statistics.median(watts) if watts else None, "samples": len(sm),
                "source": "nvml polled during the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu)], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi after the timed region (nvml unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's Codebook, CPU, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference_tokens_per_s(cfg, sample_tokens: int, repeats: int, warmup: int, min_seconds: float = 0.0):
    """The config's mode on a bounded sample (whole batch items up to `sample_tokens` tokens) through oracle/vq_oracle.py.
    `min_seconds` > 0: keep repeating (at most 200 passes) until the timed passes add up to that much CPU work."""
    import torch
    from oracle import vq_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    form, K, D = cfg["form"], cfg["K"], cfg["D"]
    per_item = bi.tokens_of(cfg["shape"][1:], D) if form == "vit" else cfg["shape"][2] * cfg["shape"][3]
    items = max(1, min(cfg["shape"][0], sample_tokens // per_item))
    shape = (items,) + tuple(cfg["shape"][1:])
    w = bi.make_codebook(form, K, D, 0)
    z = bi.make_latents(shape, 3)
    up = bi.make_latents(shape, 4)
    tokens = items * per_item
    chunk = 16384

    def run():
        if cfg["mode"] == "step":
            vo.quantise_step_chunked(form, z, w, BETA, up, chunk_tokens=chunk)
        else:
            out = vo.quantise_chunked(form, z, w, BETA, chunk_tokens=chunk)       # encode_imgs runs the whole forward
            if cfg["mode"] == "roundtrip":
                vo.indices_to_embeddings(form, out.indices.reshape(items, -1), w)

    times = []
    i = 0
    while i < warmup + repeats or (min_seconds > 0 and sum(times) < min_seconds and len(times) < 200):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        i += 1
    sec = sum(times) / len(times)
    what = {"step": "fwd + autograd backward", "encode": "forward (encode_imgs keeps the indices)",
            "roundtrip": "forward + indices_to_embeddings"}[cfg["mode"]]
    sample = (f"{tokens} tokens ({items} batch items) of the workload, token-chunked {chunk}, torch CPU ops restating the "
              f"reference Codebook ({what}), {sec:.2f} s per pass, {len(times)} timed passes = {sum(times):.1f} s of CPU work")
    return tokens / sec, cores, tokens, sec, sample


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, cores, tokens, sec, sample = cpu_reference_tokens_per_s(cfg, 16384, repeats=max(1, args.steps),
                                                                 warmup=max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": metric_name(cfg), "value": tps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "K": cfg["K"], "D": cfg["D"], "sample_tokens_per_step": tokens},
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import torch
    import torch.distributed as dist

    import vq_b200
    from vq_b200 import _lib
    from vq_b200 import dist as vq_dist
    from vq_b200 import functional as F_vq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = _peaks()

    form, K, D, mode = cfg["form"], cfg["K"], cfg["D"], cfg["mode"]
    shape = bi.local_shape(cfg, world, args.scaling)
    T = bi.tokens_of(shape, D)
    tok_major = form == "vit"
    set_bytes = T * D * 4 * (2 if mode == "step" else 1)
    n_sets = max(args.sets, min(64, math.ceil(140e6 / set_bytes)))          # rotating inputs exceed the 126 MB L2
    n_sets += n_sets & 1      # even: with the peer exchange a CUDA graph is tied to (input set, exchange slot parity)
    weight = bi.make_codebook(form, K, D, 0).to(dev).requires_grad_(mode == "step")
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    zs = [torch.randn(*shape, device=dev, generator=g) for _ in range(n_sets)]
    ups = [torch.randn(*shape, device=dev, generator=g) for _ in range(n_sets)] if mode == "step" else None
    stepper = vq_dist.ShardedQuantiser(form, BETA, world_size=world, exact_scan=args.exact_scan, exchange=args.exchange,
                                       graphs=not args.no_graphs)
    prepared = F_vq.prepare_codebook(weight.detach()) if mode != "step" else None
    counter = [0]                                     # steps so far: input sets rotate without a break between phases

    def one_step(eager=False):
        i = counter[0]
        counter[0] += 1
        if mode == "step":
            return stepper.step(zs[i % n_sets], ups[i % n_sets], weight, eager=eager)
        idx = F_vq.encode_indices(zs[i % n_sets], weight, form, prepared=prepared, exact_scan=args.exact_scan)
        if mode == "roundtrip":
            dec = F_vq.indices_to_embeddings(idx.view(shape[0], -1), weight, form, prepared=prepared, check_indices=False)
            return {"indices": idx, "decoded": dec}
        return {"indices": idx}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = mode == "step" and stepper.graphs and stepper._graph_capable()
    # with graphs, every input set is seen twice before timing: one eager step, one capture
    for i in range(max(args.warmup, 2 * n_sets)):      # (also lets the caching allocator settle on every rotating set)
        one_step()
    barrier()
    sampler = ClockSampler(local_rank).start()
    # the step runs as a replayed CUDA graph; every profile_every-th step of the timed region runs eagerly instead, with
    # CUDA-event pairs around the filter and exact/finish kernels (events cannot be read out of a graph replay)
    search_mask = (1 << _lib.PROFILE_SEARCH) | (1 << _lib.PROFILE_EXACT_FINISH)
    _lib.check(lib.vq_profile_begin(1 if graphed else args.profile_every, search_mask))
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    replayed0 = stepper.graph_kernel_launches
    wall0 = time.perf_counter()
    start.record()
    for i in range(args.steps):
        out = one_step(eager=graphed and (i % args.profile_every == 0))
    stop.record()
    barrier()
    timed_wall_s = time.perf_counter() - wall0
    replayed_launches = stepper.graph_kernel_launches - replayed0      # kernels run by the graph replays of the timed region
    ms_total = start.elapsed_time(stop)
    search_ms, search_n, launches = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(lib.vq_profile_end(ctypes.byref(search_ms), ctypes.byref(search_n), ctypes.byref(launches)))

    def read_slots(names):
        res = {}
        for name in names:
            ms_, n_ = ctypes.c_double(0), ctypes.c_int64(0)
            _lib.check(lib.vq_profile_slot(getattr(_lib, "PROFILE_" + name.upper()), ctypes.byref(ms_), ctypes.byref(n_)))
            res[name] = ms_.value / max(1, n_.value)
        return res

    slots = read_slots(["exact_finish"])
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * T / (ms_step * 1e-3)
    search_avg_ms = search_ms.value / max(1, search_n.value)
    exact_avg_ms = slots["exact_finish"]

    # ---- the other kernel families: a short pass after the timed region, every kernel bracketed (an event pair adds
    # ~5 us of idle time in front of its kernel, so these are upper bounds), and the same steps seen by CUPTI
    # (torch.profiler: in-stream kernel durations without that idle time)
    _lib.check(lib.vq_profile_begin(1, 0))
    for i in range(5):
        one_step(eager=True)
    torch.cuda.synchronize()
    _lib.check(lib.vq_profile_end(None, None, None))
    slots.update(read_slots(["prep_codebook", "prep_tokens", "tail", "backward_tokens", "codebook_grad"]))
    cupti_us = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(5):
                one_step(eager=True)
            torch.cuda.synchronize()
        acc = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and "vq::" in e.name:
                key = e.name.split("(")[0].replace("void ", "")
                a_ = acc.setdefault(key, [0, 0.0])
                a_[0] += 1
                a_[1] += e.time_range.end - e.time_range.start
        cupti_us = {k: v[1] / v[0] for k, v in acc.items()}
    except Exception as exc:      # CUPTI unavailable: the event-based numbers stand
        cupti_us = {"unavailable": repr(exc)[:100]}
    barrier()

    # ---- sustained leg: the same step back to back for >= 2 s, with the clock / power trace beside it ----------------
    sustained = None
    if not args.skip_sustained:
        n_sus = max(args.steps, int(math.ceil(2.2 / max(ms_step * 1e-3, 1e-6))))
        n_sus = min(n_sus, 200000)
        s_sampler = ClockSampler(local_rank).start()
        every = max(50, n_sus // 40)
        _lib.check(lib.vq_profile_begin(1 if graphed else every, search_mask))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for i in range(n_sus):
            one_step(eager=graphed and (i % every == 0))
        s1.record()
        barrier()
        sm_, sn_ = ctypes.c_double(0), ctypes.c_int64(0)
        _lib.check(lib.vq_profile_end(ctypes.byref(sm_), ctypes.byref(sn_), None))
        read_slots(["exact_finish"])
        s_clocks = s_sampler.stop()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        s_ms_step = float(ts.item()) / n_sus
        s_search_ms = sm_.value / max(1, sn_.value)
        s_tf = 2.0 * K * D * T / (s_search_ms * 1e-3) / 1e12 if s_search_ms > 0 else 0.0
        sustained = {"steps": n_sus, "seconds": float(ts.item()) * 1e-3, "ms_per_step": s_ms_step,
                     "value": world * T / (s_ms_step * 1e-3), "search_avg_launch_ms": s_search_ms,
                     "search_tflops": s_tf, "search_frac_of_sustained_peak": s_tf / peaks["tf_sustained"],
                     "clocks": s_clocks,
                     "note": "graph replays back to back; the search kernel is bracketed by events on a few eager steps only"}

    # ---- module path: the drop-in nn.Module + autograd (what a user of the reference calls) --------------------------
    module_path = None
    if mode == "step" and not args.skip_module:
        from vq_b200.vitvqgan import Codebook as VitCodebook
        from vq_b200.vqgan import Codebook as VqganCodebook
        m = (VitCodebook if tok_major else VqganCodebook)(K, D, BETA).to(dev)
        with torch.no_grad():
            m.embedding.weight.copy_(weight)
        one = torch.ones((), device=dev)
        zr = [z.clone().requires_grad_(True) for z in zs[:min(n_sets, 4)]]

        def module_step(i):
            z = zr[i % len(zr)]
            z.grad = None
            m.embedding.weight.grad = None
            z_q, _idx, loss = m(z)
            torch.autograd.backward([z_q, loss], [ups[i % n_sets], one])     # upstream gradient fed directly: no extra kernels

        for i in range(max(3, 2 * len(zr))):
            module_step(i)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_mod = max(5, min(args.steps, 50))
        m0.record()
        for i in range(n_mod):
            module_step(i)
        m1.record()
        torch.cuda.synchronize()
        mod_ms = m0.elapsed_time(m1) / n_mod
        module_path = {"value": T / (mod_ms * 1e-3), "unit": UNIT + " (this rank, no exchange)", "ms_per_step": mod_ms,
                       "what": f"vq_b200.{'vitvqgan' if tok_major else 'vqgan'}.Codebook.forward + torch.autograd.backward"}
        del m, zr
        barrier()

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----------------------------------------
    e2e = None
    if not args.skip_e2e:
        stream = torch.cuda.current_stream(dev).cuda_stream
        e_steps = max(3, min(args.steps, 10))
        numel = T * D
        if mode == "step" and tok_major:
            arena_bytes = _lib.size_query("vq_host_step_arena_bytes", T, K, D)
            arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
            hz = [torch.randn(T, D).pin_memory() for _ in range(2)]
            hg = [torch.randn(T, D).pin_memory() for _ in range(2)]
            hw_ = weight.detach().cpu().pin_memory()
            o_zq, o_gz = torch.empty(T, D).pin_memory(), torch.empty(T, D).pin_memory()
            o_idx = torch.empty(T, dtype=torch.int64).pin_memory()
            o_loss, o_gw = torch.empty(1).pin_memory(), torch.empty(K, D).pin_memory()
            o_stats = torch.empty(_lib.STATS_LEN, dtype=torch.int64).pin_memory()

            def host_step(i):
                _lib.check(lib.vq_host_step(hz[i % 2].data_ptr(), hg[i % 2].data_ptr(), T, hw_.data_ptr(), K, D, 0,
                                            BETA, o_zq.data_ptr(), o_idx.data_ptr(), o_loss.data_ptr(), o_gz.data_ptr(),
                                            o_gw.data_ptr(), o_stats.data_ptr(), arena.data_ptr(), arena_bytes, stream))

            h2d = numel * 4 * 2 + K * D * 4
            d2h = numel * 4 * 2 + T * 8 + K * D * 4 + 4 + _lib.STATS_LEN * 8
            note = ("vq_host_step (C ABI): pinned host in/out, all outputs (z_q, idx, loss, grad_z, grad_weight) copied back "
                    "every step")
        else:
            hz = [torch.randn(*shape).pin_memory() for _ in range(2)]
            hg = [torch.randn(*shape).pin_memory() for _ in range(2)] if mode == "step" else None
            dz, dg = torch.empty(*shape, device=dev), (torch.empty(*shape, device=dev) if mode == "step" else None)
            o_idx = torch.empty(T, dtype=torch.int64).pin_memory()
            o_a = torch.empty(*shape).pin_memory() if mode != "encode" else None        # z_q (step) / decoded (roundtrip)
            o_b = torch.empty(*shape).pin_memory() if mode == "step" else None          # grad_z
            o_gw = torch.empty(K, D).pin_memory() if mode == "step" else None
            o_loss = torch.empty(1).pin_memory() if mode == "step" else None

            def host_step(i):
                dz.copy_(hz[i % 2], non_blocking=True)
                if mode == "step":
                    dg.copy_(hg[i % 2], non_blocking=True)
                    r = stepper.step(dz, dg, weight, eager=True)
                    o_a.copy_(r["z_q"], non_blocking=True); o_b.copy_(r["grad_z"], non_blocking=True)
                    o_idx.copy_(r["indices"], non_blocking=True); o_gw.copy_(r["grad_weight"], non_blocking=True)
                    o_loss.copy_(r["loss"].reshape(1), non_blocking=True)
                else:
                    idx = F_vq.encode_indices(dz, weight, form, prepared=prepared)
                    o_idx.copy_(idx, non_blocking=True)
                    if mode == "roundtrip":
                        o_a.copy_(F_vq.indices_to_embeddings(idx.view(shape[0], -1), weight, form, prepared=prepared,
                                                             check_indices=False), non_blocking=True)

            h2d = numel * 4 * (2 if mode == "step" else 1)
            d2h = T * 8 + (numel * 4 * 2 + K * D * 4 + 4 if mode == "step" else (numel * 4 if mode == "roundtrip" else 0))
            note = "functional API with pinned host tensors: H2D copy of the inputs, the op, D2H copy of every output"
        for i in range(2):
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record()
        for i in range(e_steps):
            host_step(i)
        ee.record()
        torch.cuda.synchronize()
        e_ms = max(es.elapsed_time(ee), (time.perf_counter() - t0) * 1e3) / e_steps
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        # the floor of this path: the same bytes as plain pinned copies, both directions at once, no kernels
        d_in, d_out = torch.empty(h2d // 4, device=dev), torch.empty(max(d2h // 4, 1), device=dev)
        h_in, h_out = torch.empty(h2d // 4).pin_memory(), torch.empty(max(d2h // 4, 1)).pin_memory()
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        pcie_ms = (time.perf_counter() - t0) / 5 * 1e3
        del d_in, d_out, h_in, h_out
        e2e = {"value": world * T / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e_ms, "pcie_floor_ms": pcie_ms, "frac_of_pcie_floor": pcie_ms / e_ms,
               "note": note + "; pcie_floor_ms = the same bytes as plain pinned copies in both directions at once, no kernels"}

    # ---- parity, outside the timed region ------------------------------------------------------------------------------
    parity = {}
    stats = out["stats"].tolist() if isinstance(out, dict) and "stats" in out else None
    if stats:
        parity["near_tie_rows_last_step"] = stats[_lib.STAT_NEAR_TIE_ROWS]
        parity["fallback_rows_last_step"] = stats[_lib.STAT_FALLBACK_ROWS]
    # tensor-core search == exhaustive fp32 search on a sample of the rows of the last input set
    sample_items = max(1, min(shape[0], 8192 // max(1, T // shape[0])))
    zs_s = zs[0][:sample_items].contiguous()
    w_d = weight.detach()
    prep_s = prepared if prepared is not None else F_vq.prepare_codebook(w_d)
    parity["tensor_core_search_equals_exhaustive_on_sample"] = bool(torch.equal(
        F_vq.encode_indices(zs_s, w_d, form, prepared=prep_s), F_vq.encode_indices(zs_s, w_d, form, prepared=prep_s, exact_scan=True)))
    parity["sample_tokens"] = bi.tokens_of(tuple(zs_s.shape), D)
    peer_timeouts = 0
    if mode == "step":
        hist = out["histogram"].to(torch.int64)
        parity["histogram_sum_equals_global_tokens"] = int(hist.sum().item()) == world * T
        if world > 1:
            if args.exchange == "peer":
                peer_timeouts = int(out["stats"][_lib.STAT_PEER_TIMEOUT].item())
            # (a) every rank holds the same bits of the reduced quantities
            packed = torch.cat([out["grad_weight"].reshape(-1).view(torch.int32).to(torch.int64), hist,
                                out["loss"].reshape(1).view(torch.int32).to(torch.int64)])
            gathered = [torch.empty_like(packed) for _ in range(world)]
            dist.all_gather(gathered, packed)
            parity["cross_rank_bit_identical"] = all(bool(torch.equal(gathered[0], g_)) for g_ in gathered[1:])
            # (b) a small sharded step equals the single-GPU step on the concatenation of all ranks' shards, bit for bit
            #     (DDP mean of per-rank mean-loss gradients == the global batch, trainers/vitgqgan.py:184)
            small = max(1, min(2, shape[0]))
            z_s, u_s = zs[0][:small].contiguous(), ups[0][:small].contiguous()
            sh = {k: v.clone() for k, v in stepper.step(z_s, u_s, weight, eager=True).items()}
            zg = [torch.empty_like(z_s) for _ in range(world)]
            ug = [torch.empty_like(u_s) for _ in range(world)]
            dist.all_gather(zg, z_s)
            dist.all_gather(ug, u_s)
            single = vq_dist.ShardedQuantiser(form, BETA, world_size=1)
            ref = single.step(torch.cat(zg), torch.cat(ug), weight)
            rows = slice(rank * small, (rank + 1) * small)
            same = (torch.equal(sh["grad_weight"], ref["grad_weight"]) and float(sh["loss"]) == float(ref["loss"])
                    and torch.equal(sh["histogram"].to(torch.int64), ref["histogram"].to(torch.int64))
                    and torch.equal(sh["z_q"], ref["z_q"][rows]) and torch.equal(sh["grad_z"], ref["grad_z"][rows])
                    and torch.equal(sh["indices"], ref["indices"].view(world * small, -1)[rows].reshape(-1)))
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity["equals_single_gpu_on_concatenated_batch"] = bool(flag.item())
            parity["concat_check_tokens"] = world * bi.tokens_of(tuple(z_s.shape), D)
            single.close()
    if getattr(stepper, "trace", None):
        torch.cuda.synchronize()
        tr = stepper.trace[args.warmup + 2: args.warmup + args.steps]
        n_ = len(tr) - 1
        f = lambda a, b: sum(t_[a].elapsed_time(t_[b]) for t_ in tr[:-1]) / n_ * 1e3
        nxt = sum(tr[i][4].elapsed_time(tr[i + 1][1]) for i in range(n_)) / n_ * 1e3
        print(f"[rank {rank}] us: fwd_end->exchange_end {f(1, 2):.1f}  fwd_end->bwd_tokens_end {f(1, 3):.1f}  "
              f"fwd_end->join {f(1, 4):.1f}  join->next fwd_end {nxt:.1f}", file=sys.stderr, flush=True)
    stepper.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert peer_timeouts == 0, "a peer never published its step (exchange kernel timed out)"
    for key in ("tensor_core_search_equals_exhaustive_on_sample", "histogram_sum_equals_global_tokens",
                "cross_rank_bit_identical", "equals_single_gpu_on_concatenated_batch"):
        assert parity.get(key, True), f"parity check failed: {key}"

    # ---- roofline of the dominant kernel (nearest-code search) -------------------------------------------------------
    flops_per_launch = 2.0 * K * D * T
    achieved_tf = flops_per_launch / (search_avg_ms * 1e-3) / 1e12 if search_avg_ms > 0 else 0.0
    tc_path = not (args.exact_scan or not stepper.uses_tensor_cores(T, K, D))
    tc16 = tc_path and D == 32 and K % 512 == 0
    at_full_clock = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    burst = timed_wall_s < 1.0 and at_full_clock
    peak_tf = peaks["tf_burst"] if burst else peaks["tf_sustained"]
    search_total_ms = search_avg_ms + (exact_avg_ms if tc_path else 0.0)
    roofline = {"kernel": ("k_dist_tc16: tcgen05 distance + running-maximum filter (z.C^T for every token x code)" if tc16 else
                           "k_dist_tc: tcgen05 distance filter, fp32 accumulators (D = 256: clusters of two 128-row CTAs sharing the "
                           "codebook stream by TMA multicast)" if tc_path else
                           "k_scan_exact: exhaustive fp32 distance + argmin"),
                "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf,
                "traffic": DRAM_TRAFFIC_FILTER_CFG3 if (tc16 and T == 262144 and K == 8192) else None,
                "peak_kind": "burst" if burst else "sustained",
                "peak_source": peaks["source"] + (f": burst bf16 peak (timed region {timed_wall_s * 1e3:.0f} ms at "
                                                   f"{clocks.get('sm_mhz')} MHz)" if burst else
                                                   ": sustained bf16 peak (timed region >= 1 s or clocks below max)"),
                "frac_vs_burst_peak": achieved_tf / peaks["tf_burst"], "frac_vs_sustained_peak": achieved_tf / peaks["tf_sustained"],
                "avg_launch_ms": search_avg_ms, "share_of_step": search_avg_ms / ms_step,
                "algorithmic": f"2*K*D = {2 * K * D} flop/token x {T} tokens per launch",
                "path": ("tcgen05 fp16 filter (fp16 accumulators) -> exact fp32 rescoring + finish kernel" if tc16 else
                         "tcgen05 filter (fp32 accumulators) -> exact fp32 rescoring -> finish kernel" if tc_path else
                         "exact fp32 SIMT scan"),
                "distance_plus_argmin": {
                    "what": "filter + everything behind it: exact rescoring, the rows the filter left undecided, and the finish "
                            "pass (idx / z_q / loss / histogram / segment sums; one launch at D = 32, 2-3 launches otherwise)",
                    "ms": search_total_ms,
                    "tflops": flops_per_launch / (search_total_ms * 1e-3) / 1e12 if search_total_ms > 0 else 0.0,
                    "frac_vs_burst_peak": flops_per_launch / (search_total_ms * 1e-3) / 1e12 / peaks["tf_burst"] if search_total_ms > 0 else 0.0,
                    "frac_vs_sustained_peak": flops_per_launch / (search_total_ms * 1e-3) / 1e12 / peaks["tf_sustained"] if search_total_ms > 0 else 0.0,
                } if tc_path else None,
                "sustained_leg": sustained}

    # ---- HBM side ------------------------------------------------------------------------------------------------------
    peak_gbs = peaks["hbm_gbs"]
    alg_bytes = {"step": (20 * D + 16) * T + 4 * K * D, "encode": (4 * D + 8) * T, "roundtrip": (4 * D + 8) * T * 2}[mode]

    def cupti(fragment):
        if not cupti_us:
            return None
        hits = [v for k, v in cupti_us.items() if fragment in k and isinstance(v, float)]
        return hits[0] * 1e-3 if hits else None

    def hbm_kernel(ms_events, fragment, alg, iface, alg_what, iface_what):
        ms_cupti = cupti(fragment)
        ms = ms_cupti if ms_cupti else ms_events
        if not ms or ms <= 0:
            return None
        return {"avg_launch_ms": ms, "timed_by": "cupti" if ms_cupti else "cuda events (incl. ~5 us idle before the kernel)",
                "avg_launch_ms_events": ms_events,
                "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak_gbs,
                "algorithmic": alg_what,
                "interface_bytes": iface, "interface_gbs": iface / (ms * 1e-3) / 1e9,
                "interface_frac": iface / (ms * 1e-3) / 1e9 / peak_gbs, "interface": iface_what}

    kernels = {}
    if mode == "step" and tc16 and tok_major:
        kernels["k_prep_rows_fused (token rows + codebook)"] = hbm_kernel(
            slots["prep_tokens"], "k_prep_rows_fused", 4 * D * T + 4 * K * D, (10 * D + 8) * T + 14 * D * K,
            "read z 4D per token + the codebook 4KD (the unit rows it writes are an intermediate of this implementation)",
            "tokens: read z 4D; write zn32 4D, zn16 2D, row_sq + denom 8; codebook: read E 4D, write en32 + en32c 8D, en16 2D")
        kernels["k_exact_finish16 (exact rescoring + idx / z_q / loss / hist / segment sums)"] = hbm_kernel(
            exact_avg_ms, "k_exact_finish16", (8 * D + 8) * T, (8 * D + 8 + 16 + 4) * T,
            "SURVEY 8(d) forward gather/loss/STE: read zn 4D, write z_q 4D, write idx 8",
            "read zn32 4D + record 16 + row_sq 4, write z_q 4D + idx 8 (cells and code rows: ~1.5 KB/token of L2 reads)")
        kernels["k_backward_fused (grad_z + grad_E)"] = hbm_kernel(
            slots["backward_tokens"], "k_backward_fused", (12 * D + 8) * T + 4 * K * D, (12 * D + 12) * T + (16 * D + 8) * K,
            "SURVEY 8(d) backward: read G 4D, zn 4D, idx 8, write grad_z 4D; write grad_E 4KD once",
            "tokens: read G 4D, zn 4D, idx 8, denom 4; write grad_z 4D (code rows from L2); codebook: read seg sums 8D + 8, en 4D, "
            "write grad_E 4D")
    if tc_path and not tc16 and not tok_major:
        kernels["k_prep_nchw_fused (NCHW latents -> unit token rows fp32 + fp16, row norms: one launch)"] = hbm_kernel(
            slots["prep_tokens"], "k_prep_nchw_fused", 4 * D * T, (10 * D + 8) * T,
            "read z 4D per token (the unit rows it writes are an intermediate of this implementation)",
            "read z 4D; write zn32 4D, zn16 2D, row_sq + denom 8")
    iface_total = sum(k["interface_bytes"] for k in kernels.values() if k) + (2 * D * T + 16 * T if tc16 and mode == "step" else 0)
    non_filter_ms = max(ms_step - search_avg_ms, 1e-6)
    hbm = {"peak_gbs": peak_gbs, "algorithmic_bytes_per_step": alg_bytes,
           "algorithmic": {"step": "(20D + 16) per token + 4KD", "encode": "(4D + 8) per token",
                           "roundtrip": "(4D + 8) per token, encode + decode"}[mode],
           "kernels": {k: v for k, v in kernels.items() if v},
           "ms_step_minus_filter": non_filter_ms,
           "composite_frac": alg_bytes / (non_filter_ms * 1e-3) / 1e9 / peak_gbs,
           "composite_note": "algorithmic bytes of the whole step over ALL time outside the tensor-core filter (the exact/"
                             "finish kernel that writes 8D + 8 of them included)",
           "interface_over_algorithmic": (iface_total / alg_bytes) if iface_total else None,
           "note": "per kernel: in-stream duration of a separate 5-step pass after the timed region (CUPTI; event pairs as fallback)"}
    kernel_ms = dict(slots, search=search_avg_ms)

    # ---- cpu_baseline: oracle port on this box's host cores, bounded sample (N = 1 only) ------------------------------
    cpu = None
    if not args.skip_cpu and world == 1:
        tps, cores, tokens, sec, sample = cpu_reference_tokens_per_s(cfg, 131072, repeats=3, warmup=1, min_seconds=10.0)
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    workload = (cfg["desc"] + f"; {args.scaling} scaling: {shape[0]} batch items = {T} tokens per GPU, "
                f"{world * T} tokens over {world} GPU(s), fp32")
    line = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "name": cfg["name"], "mode": mode, "form": form, "K": K, "D": D,
                       "tokens_per_gpu": T, "global_tokens": T * world,
                       "parallelism": f"tokens sharded over {world} GPU(s), codebook replicated"
                                      + (f"; backward exchange: {'fused peer-memory kernel over NVLink (CUDA IPC)' if args.exchange == 'peer' else 'NCCL all-reduce of the packed int64 buffer'}" if world > 1 and mode == "step" else ""),
                       "l2": f"inputs rotate over {n_sets} resident sets ({n_sets * set_bytes >> 20} MiB) > 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.value) + int(replayed_launches),
            "roofline": roofline, "hbm_side": hbm, "kernel_ms_events": kernel_ms, "kernel_us_cupti": cupti_us,
            "module_path": module_path,
            "profile_sampling": f"timed region: CUDA-event pairs around the search and exact/finish kernels on every "
                                f"{args.profile_every}th step" + (" (those steps run eagerly, the others replay a CUDA graph "
                                "of the same launches)" if graphed else "") + "; other kernels: separate 5-step pass",
            "cuda_graph": bool(graphed),
            "cpu_baseline": cpu,
            "parity": parity}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", help="cfg1 | cfg2 | cfg2fwd | cfg3 (default) | cfg4 | sweep:T,K,D | cfg3pre (cfg3 behind the fused pre_quant Linear(512, 32), encode)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's batch on every GPU; strong: the config's batch split over the GPUs")
    ap.add_argument("--sets", type=int, default=4, help="resident input sets rotated between steps (at least; more for small shapes)")
    ap.add_argument("--exact-scan", action="store_true", help="force the exhaustive fp32 SIMT search")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"],
                    help="N > 1: fused peer-memory exchange kernel (default) or one NCCL all-reduce")
    ap.add_argument("--profile-every", type=int, default=10,
                    help="bracket the kernels with CUDA events on every n-th timed step (event records cost ~2%% of a step)")
    ap.add_argument("--no-graphs", action="store_true", help="launch every step eagerly (no CUDA-graph replay)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-sustained", action="store_true")
    ap.add_argument("--skip-module", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.config == "cfg3pre":
        # SURVEY.md 8(f) rank 1 (pre_quant fused into the quantiser): its own workload, measured to the same contract
        import bench_projected
        if args.impl == "reference":
            bench_projected.run_reference(args)
        else:
            bench_projected.run_b200(args, _peaks(), ClockSampler)
        return
    cfg = bi.resolve_config(args.config)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
