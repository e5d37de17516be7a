#!/usr/bin/env python
"""bench.py -- VQ codebook quantiser fwd+bwd throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[2] -- the ViT-VQGAN training-step quantiser, fwd + bwd
(straight-through + codebook gradient), K = 8192 codes x D = 32, batch 256 x 1024 tokens (262 144 tokens)
PER GPU, fp32, synthetic N(0,1) latents and upstream gradients, N(0,1) codebook.  Weak scaling: every
rank quantises its own 262 144 tokens against the replicated codebook; the only exchange is one packed
int64 all-reduce of the codebook-gradient segment sums, the usage histogram and the loss partial.

One "step" = one pass of the hot path over one batch: codebook prepare, forward (z_q, indices, loss),
backward (grad_z, grad_weight).  Inputs rotate over several resident sets so that every step reads
data that is not in L2 (a step's working set alone is > 126 MB).

Prints ONE JSON line (rank 0).  See the task contract for the keys; notes:
  value      tokens/s over all ranks with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the host-buffer C-ABI call (pinned host -> device -> pinned host inside
             the timed region), summed over ranks
  roofline   the nearest-code search kernel: algorithmic 2*K*D flop/token over its live CUDA-event time
  cpu_baseline  the oracle port of the reference (torch CPU ops, all host threads) on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import json
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

K_CODES, DIM, IMGS_PER_GPU, TOKENS_PER_IMG, BETA = 8192, 32, 256, 1024, 0.25
# dram__bytes_read.sum + dram__bytes_write.sum of one k_dist_tc16 launch (ncu --set full, profiles/r01_*): the
# fp16 token rows (16.8 MB) + the 0.5 MB fp16 codebook; the verdict records (12.6 MB) are still in L2 when the
# kernel ends, the codebook tiles stream from L2
DRAM_TRAFFIC_FILTER = 17.33e6
METRIC = "vq_tokens_per_sec_fwd_bwd_K8192_D32"
UNIT = "tokens/s"
WORKLOAD = ("cfg3: ViT-VQGAN quantiser fwd+bwd (STE + codebook grad), codebook 8192x32, "
            "256 img x 1024 tok = 262144 tokens per GPU, fp32")


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

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
                self.samples.append((float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=1)
        if not self.samples:
            return self._smi_once()
        sm = [s[0] for s in self.samples]
        reasons = sorted({name for _, mask in self.samples for name, bit in self.REASONS if mask & bit})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml polled during the timed region"}

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
def cpu_reference_tokens_per_s(sample_imgs: int, repeats: int, warmup: int):
    import torch
    from oracle import vq_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = vo.make_codebook("vit", K_CODES, DIM, 0)
    z = vo.make_latents((sample_imgs, TOKENS_PER_IMG, DIM), 3)
    up = vo.make_latents((sample_imgs, TOKENS_PER_IMG, DIM), 4)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        vo.quantise_step_chunked("vit", z, w, BETA, up, chunk_tokens=16384)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tokens = sample_imgs * TOKENS_PER_IMG
    return tokens / (sum(times) / len(times)), cores, tokens, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_imgs = 16
    tps, cores, tokens, sec = cpu_reference_tokens_per_s(sample_imgs, repeats=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    sample = (f"{tokens} tokens ({sample_imgs} img) of the workload per step, token-chunked 16384, torch CPU ops "
              f"restating models/vitvqgan.py:151-171 + autograd backward")
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K": K_CODES, "D": DIM, "sample_tokens_per_step": tokens},
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import vq_b200
    from vq_b200 import _lib
    from vq_b200 import dist as vq_dist
    from oracle import vq_oracle as vo   # seeded input generators + the cpu_baseline leg only

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

    T = IMGS_PER_GPU * TOKENS_PER_IMG
    n_sets = args.sets
    weight = vo.make_codebook("vit", K_CODES, DIM, 0).to(dev).requires_grad_(True)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    zs = [torch.randn(IMGS_PER_GPU, TOKENS_PER_IMG, DIM, device=dev, generator=g) for _ in range(n_sets)]
    ups = [torch.randn(IMGS_PER_GPU, TOKENS_PER_IMG, DIM, device=dev, generator=g) for _ in range(n_sets)]
    stepper = vq_dist.ShardedQuantiser("vit", BETA, world_size=world, exact_scan=args.exact_scan, exchange=args.exchange,
                                       graphs=not args.no_graphs)

    counter = [0]                                     # steps so far: input sets rotate without a break between phases

    def one_step(_unused=None, eager=False):
        i = counter[0]
        counter[0] += 1
        return stepper.step(zs[i % n_sets], ups[i % n_sets], weight, eager=eager)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # with graphs, every input set is seen twice before timing: one eager step, one capture
    for i in range(max(args.warmup, 2 * n_sets if stepper.graphs else 0)):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the step runs as a replayed CUDA graph; every profile_every-th step of the timed region runs eagerly instead, with
    # CUDA-event pairs around the filter and exact/finish kernels (events cannot be read out of a graph replay)
    search_mask = (1 << _lib.PROFILE_SEARCH) | (1 << _lib.PROFILE_EXACT_FINISH)
    graphed = stepper.graphs and stepper._graph_capable()
    _lib.check(lib.vq_profile_begin(1 if graphed else args.profile_every, search_mask))
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    replayed0 = stepper.graph_kernel_launches
    start.record()
    for i in range(args.steps):
        out = one_step(args.warmup + i, eager=graphed and (i % args.profile_every == 0))
    stop.record()
    barrier()
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
    exact_ms, exact_n = ctypes.c_double(slots["exact_finish"]), ctypes.c_int64(1 if slots["exact_finish"] > 0 else 0)
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * T / (ms_step * 1e-3)

    # ---- the other kernel families: a short pass after the timed region, every kernel bracketed (an event pair adds
    # ~5 us of idle time in front of its kernel, so these are upper bounds), and the same steps seen by CUPTI
    # (torch.profiler: in-stream kernel durations without that idle time)
    _lib.check(lib.vq_profile_begin(1, 0))
    for i in range(5):
        one_step(args.warmup + args.steps + i, eager=True)
    torch.cuda.synchronize()
    _lib.check(lib.vq_profile_end(None, None, None))
    slots.update(read_slots(["prep_codebook", "prep_tokens", "tail", "backward_tokens", "codebook_grad"]))
    cupti_us = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(5):
                one_step(args.warmup + args.steps + 5 + i, eager=True)
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

    # ---- e2e: host buffers through the C ABI (vq_host_step), copies inside the timed region ------
    e2e = None
    if not args.skip_e2e:
        arena_bytes = _lib.size_query("vq_host_step_arena_bytes", T, K_CODES, DIM)
        arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
        hz = [torch.randn(T, DIM).pin_memory() for _ in range(2)]
        hg = [torch.randn(T, DIM).pin_memory() for _ in range(2)]
        hw = weight.detach().cpu().pin_memory()
        o_zq, o_gz = torch.empty(T, DIM).pin_memory(), torch.empty(T, DIM).pin_memory()
        o_idx = torch.empty(T, dtype=torch.int64).pin_memory()
        o_loss, o_gw = torch.empty(1).pin_memory(), torch.empty(K_CODES, DIM).pin_memory()
        o_stats = torch.empty(_lib.STATS_LEN, dtype=torch.int64).pin_memory()
        stream = torch.cuda.current_stream(dev).cuda_stream

        def host_step(i):
            _lib.check(lib.vq_host_step(hz[i % 2].data_ptr(), hg[i % 2].data_ptr(), T, hw.data_ptr(), K_CODES, DIM, 0,
                                        BETA, o_zq.data_ptr(), o_idx.data_ptr(), o_loss.data_ptr(), o_gz.data_ptr(),
                                        o_gw.data_ptr(), o_stats.data_ptr(), arena.data_ptr(), arena_bytes, stream))

        e_steps = max(3, min(args.steps, 10))
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
        h2d = T * DIM * 4 * 2 + K_CODES * DIM * 4
        d2h = T * DIM * 4 * 2 + T * 8 + K_CODES * DIM * 4 + 4 + _lib.STATS_LEN * 8
        # the floor of this path: the same bytes as plain pinned copies, both directions at once, no kernels
        d_in, d_out = torch.empty(2 * T * DIM, device=dev), torch.empty(2 * T * DIM + 2 * T, device=dev)
        h_in, h_out = torch.empty(2 * T * DIM).pin_memory(), torch.empty(2 * T * DIM + 2 * T).pin_memory()
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
               "note": "vq_host_step: pinned host in/out, all outputs (z_q, idx, loss, grad_z, grad_weight) copied back every "
                       "step; pcie_floor_ms = the same bytes as plain pinned copies in both directions at once, no kernels"}

    if getattr(stepper, "trace", None):
        torch.cuda.synchronize()
        tr = stepper.trace[args.warmup + 2: args.warmup + args.steps]
        n_ = len(tr) - 1
        f = lambda a, b: sum(t[a].elapsed_time(t[b]) for t in tr[:-1]) / n_ * 1e3
        nxt = sum(tr[i][4].elapsed_time(tr[i + 1][1]) for i in range(n_)) / n_ * 1e3
        print(f"[rank {rank}] us: fwd_end->exchange_end {f(1, 2):.1f}  fwd_end->bwd_tokens_end {f(1, 3):.1f}  "
              f"fwd_end->join {f(1, 4):.1f}  join->next fwd_end {nxt:.1f}", file=sys.stderr, flush=True)
    peer_timeouts = int(out["stats"][_lib.STAT_PEER_TIMEOUT].item()) if world > 1 and args.exchange == "peer" else 0
    stepper.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert peer_timeouts == 0, "a peer never published its step (exchange kernel timed out)"

    # ---- roofline of the dominant kernel (nearest-code search) -------------------------------------
    flops_per_launch = 2.0 * K_CODES * DIM * T
    search_avg_ms = search_ms.value / max(1, search_n.value)
    achieved_tf = flops_per_launch / (search_avg_ms * 1e-3) / 1e12 if search_avg_ms > 0 else 0.0
    peak_tf = peaks["tf_sustained"]
    tc_path = not (args.exact_scan or not stepper.uses_tensor_cores(T, K_CODES, DIM))
    exact_avg_ms = exact_ms.value / max(1, exact_n.value)
    roofline = {"kernel": "k_dist_tc16: tcgen05 distance + running-maximum filter (z.C^T for every token x code)" if tc_path
                          else "k_scan_exact: exhaustive fp32 distance + argmin",
                "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": DRAM_TRAFFIC_FILTER if tc_path else None,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "avg_launch_ms": search_avg_ms, "share_of_step": search_avg_ms / ms_step,
                "algorithmic": "2*K*D = 524288 flop/token x 262144 tokens per launch",
                "path": "tcgen05 fp16 filter (fp16 accumulators) -> exact fp32 rescoring + finish kernel" if tc_path
                        else "exact fp32 SIMT scan",
                "traffic_note": "dram__bytes_read+write of one k_dist_tc16 launch, ncu --set full (profiles/)" if tc_path else None,
                "behind_the_filter": {"kernel": "k_exact_finish16: exact fp32 rescoring of the surviving cells + idx/z_q/loss/hist",
                                      "avg_launch_ms": exact_avg_ms, "bound": "L2 (scattered 128-byte code rows)",
                                      "search_total_tflops": flops_per_launch / ((search_avg_ms + exact_avg_ms) * 1e-3) / 1e12,
                                      "search_total_frac": flops_per_launch / ((search_avg_ms + exact_avg_ms) * 1e-3) / 1e12 / peak_tf}
                                     if exact_n.value else None}
    hbm_bytes_step = (20 * DIM + 16) * T + 4 * K_CODES * DIM
    non_search_ms = max(ms_step - search_avg_ms - exact_avg_ms, 1e-6)
    peak_gbs = peaks["hbm_gbs"]

    def cupti(fragment):
        if not cupti_us:
            return None
        hits = [v for k, v in cupti_us.items() if fragment in k and isinstance(v, float)]
        return hits[0] * 1e-3 if hits else None

    def hbm_kernel(ms_events, fragment, nbytes, what):
        ms_cupti = cupti(fragment)
        ms = ms_cupti if ms_cupti else ms_events
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms and ms > 0 else 0.0
        return {"avg_launch_ms": ms, "timed_by": "cupti" if ms_cupti else "cuda events (incl. ~5 us idle before the kernel)",
                "avg_launch_ms_events": ms_events, "dram_bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / peak_gbs, "bytes": what}

    hbm = {"peak_gbs": peak_gbs, "algorithmic_bytes_per_step": hbm_bytes_step,
           "kernels": {
               "k_prep_rows_fused (token rows + codebook)": hbm_kernel(
                   slots["prep_tokens"], "k_prep_rows_fused", (10 * DIM + 8) * T + 14 * DIM * K_CODES,
                   "tokens: read z 4D; write zn32 4D, zn16 2D, row_sq + denom 8; codebook: read E 4D, write en32 + en32c 8D, en16 2D"),
               "k_backward_fused (grad_z + grad_E)": hbm_kernel(
                   slots["backward_tokens"], "k_backward_fused", (12 * DIM + 12) * T + (16 * DIM + 8) * K_CODES,
                   "tokens: read G 4D, zn 4D, idx 8, denom 4; write grad_z 4D (code rows from L2); codebook: read seg sums "
                   "8D + 8, en 4D, write grad_E 4D"),
           },
           "non_search_ms": non_search_ms,
           "all_non_search_vs_20D+16": hbm_bytes_step / (non_search_ms * 1e-3) / 1e9 / peak_gbs,
           "note": "per kernel: DRAM bytes / in-stream duration of a separate 5-step pass after the timed region; the z_q / "
                   "idx / histogram / segment-sum writes of the forward are fused into k_exact_finish16 (roofline."
                   "behind_the_filter); all_non_search_vs_20D+16 = SURVEY 8(d) bytes of the step over all time outside the "
                   "filter and exact/finish kernels"}
    kernel_ms = dict(slots, search=search_avg_ms)

    # ---- cpu_baseline: oracle port on this box's host cores, bounded sample -------------------------
    cpu = None
    if not args.skip_cpu:
        tps, cores, tokens, sec = cpu_reference_tokens_per_s(sample_imgs=16, repeats=3, warmup=1)
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{tokens} tokens (16 img) of the same workload, fwd+bwd, token-chunked 16384, "
                         f"{sec:.2f} s per pass, mean of 3"}

    stats = out["stats"].tolist() if isinstance(out, dict) and "stats" in out else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K": K_CODES, "D": DIM, "tokens_per_gpu": T, "global_tokens": T * world,
                       "parallelism": f"tokens sharded over {world} GPU(s), codebook replicated"
                                      + (f"; backward exchange: {'fused peer-memory kernel over NVLink (CUDA IPC)' if args.exchange == 'peer' else 'NCCL all-reduce of the packed int64 buffer'}" if world > 1 else ""),
                       "l2": f"inputs rotate over {n_sets} resident sets ({n_sets * 2 * T * DIM * 4 >> 20} MiB) > 126 MB L2; "
                             "a step's own working set is 130 MB"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.value) + int(replayed_launches),
            "roofline": roofline, "hbm_side": hbm, "kernel_ms_events": kernel_ms, "kernel_us_cupti": cupti_us,
            "profile_sampling": f"timed region: CUDA-event pairs around the filter and exact/finish kernels on every "
                                f"{args.profile_every}th step" + (" (those steps run eagerly, the others replay a CUDA graph "
                                "of the same five launches)" if graphed else "") + "; other kernels: separate 5-step pass",
            "cuda_graph": bool(graphed),
            "cpu_baseline": cpu,
            "parity": {"near_tie_rows_last_step": stats[_lib.STAT_NEAR_TIE_ROWS] if stats else None,
                       "fallback_rows_last_step": stats[_lib.STAT_FALLBACK_ROWS] if stats else None}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sets", type=int, default=4, help="resident input sets rotated between steps")
    ap.add_argument("--exact-scan", action="store_true", help="force the exhaustive fp32 SIMT search")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"],
                    help="N > 1: fused peer-memory exchange kernel (default) or one NCCL all-reduce")
    ap.add_argument("--profile-every", type=int, default=10,
                    help="bracket the kernels with CUDA events on every n-th timed step (event records cost ~2%% of a step)")
    ap.add_argument("--no-graphs", action="store_true", help="launch every step eagerly (no CUDA-graph replay)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
