"""bench.py --config cfg3pre : the widened row of SURVEY.md section 8(f) rank 1, measured to the same contract.

Workload: ViTVQGAN.encode_imgs behind the encoder (/root/reference/models/vitvqgan.py:207-209) at BASELINE cfg3's size --
262 144 encoder rows x 512 features -> pre_quant Linear(512, 32) -> nearest code of 8192 x 32, indices only.  One step =
one pass over one batch of rows; the rows rotate over resident sets larger than the 126 MB L2.  No collective: N > 1 runs
the same batch on every rank (weak scaling, independent shards).

roofline: the new kernel of this path, k_prequant_prep (HBM-bound: it must read T*C*4 bytes of encoder rows), against the
measured copy bandwidth; the search kernel's tensor-core fraction is reported beside it.  Same keys as bench.py's line.
"""
from __future__ import annotations

import ctypes
import json
import os
import time

T_ROWS, C_IN, K_CODES, D_CODE = 262144, 512, 8192, 32
# dram__bytes_read.sum + dram__bytes_write.sum of one k_prequant_prep launch at this size (ncu --set full,
# profiles/r02b_prequant_ncu.md): 537.0 MB read (= the encoder rows, nothing re-read) + 43.8 MB of the 52.4 MB of outputs
# written back inside the launch (the rest is still in L2 when it ends)
DRAM_TRAFFIC_PREQUANT_CFG3 = 580.8e6
DESC = ("cfg3pre: ViTVQGAN.encode_imgs behind the encoder -- pre_quant Linear(512, 32) fused into the quantiser's token "
        "preparation + nearest code of 8192x32, 256 img x 1024 tok (262144 rows x 512 features), indices only")
METRIC = f"vq_tokens_per_sec_prequant_encode_K{K_CODES}_D{D_CODE}_C{C_IN}"
UNIT = "tokens/s"


def _cpu_port_tokens_per_s(sample_tokens: int, min_seconds: float):
    """oracle port: F.linear + the reference Codebook's forward (encode_imgs keeps the indices) on the host cores."""
    import torch
    from oracle import vq_oracle as vo
    import bench_inputs as bi
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = bi.make_codebook("vit", K_CODES, D_CODE, 0)
    w_pre, b_pre = vo.projection_inputs(C_IN, D_CODE, 1)
    x = bi.make_latents((sample_tokens, C_IN), 3)
    times = []
    i = 0
    while i < 2 or (sum(times) < min_seconds and len(times) < 200):
        t0 = time.perf_counter()
        with torch.no_grad():
            vo.quantise_projected(x, w_pre, b_pre, w, 0.25)
        if i >= 1:
            times.append(time.perf_counter() - t0)
        i += 1
    sec = sum(times) / len(times)
    sample = (f"{sample_tokens} rows of the workload, torch CPU ops restating pre_quant + the reference Codebook forward, "
              f"{sec:.3f} s per pass, {len(times)} timed passes = {sum(times):.1f} s of CPU work")
    return sample_tokens / sec, cores, sec, sample


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    tps, cores, sec, sample = _cpu_port_tokens_per_s(16384, 2.0 * max(1, args.steps) * 0.15)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": DESC, "K": K_CODES, "D": D_CODE, "C": C_IN, "sample_tokens_per_step": 16384},
                      "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), flush=True)


def run_b200(args, peaks, ClockSampler):
    import torch
    import torch.distributed as dist

    import bench_inputs as bi
    from vq_b200 import _lib, functional as F_vq, projected

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.load()
    T, C, K, D = T_ROWS, C_IN, K_CODES, D_CODE

    # seeded inputs (CPU generators: every rank the same codebook / projection; rows differ per rank)
    w = bi.make_codebook("vit", K, D, 0).to(dev)
    g = torch.Generator().manual_seed(1)
    bound = 1.0 / C ** 0.5
    w_pre = ((torch.rand(D, C, generator=g) * 2 - 1) * bound).to(dev)
    b_pre = ((torch.rand(D, generator=g) * 2 - 1) * bound).to(dev)
    gd = torch.Generator(device=dev).manual_seed(100 + rank)
    n_sets = max(2, args.sets - 1)                      # 3 x 537 MB by default
    xs = [torch.randn(T, C, device=dev, generator=gd) for _ in range(n_sets)]
    prepared = F_vq.prepare_codebook(w)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        return projected.encode_indices_projected(xs[i % n_sets], w_pre, b_pre, w, prepared=prepared)

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            idx = one_step(i)
        barrier()
        sampler = ClockSampler(local_rank).start()
        launches0 = lib.vq_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            idx = one_step(i)
        e1.record()
        barrier()
        clocks = sampler.stop()
        launches = lib.vq_kernel_launches() - launches0
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / args.steps
        value = world * T / (ms_step * 1e-3)

        # per-kernel times: a separate pass with every kernel family bracketed by CUDA events inside the library
        # (on the launching stream; an event pair adds a few us of idle time, so these are upper bounds)
        _lib.check(lib.vq_profile_begin(1, 0))
        for i in range(10):
            one_step(i)
        sm_, sn_ = ctypes.c_double(0), ctypes.c_int64(0)
        _lib.check(lib.vq_profile_end(ctypes.byref(sm_), ctypes.byref(sn_), None))

        def slot(s):
            ms, n = ctypes.c_double(0), ctypes.c_int64(0)
            _lib.check(lib.vq_profile_slot(s, ctypes.byref(ms), ctypes.byref(n)))
            return ms.value / max(1, n.value)
        prep_ms, exact_ms = slot(_lib.PROFILE_PREP_TOKENS), slot(_lib.PROFILE_EXACT_FINISH)
        search_ms = sm_.value / max(1, sn_.value)

        # ---- unfused composition on the same box (what the reference's call sites do: cuBLAS Linear, then the quantiser)
        def unfused(i):
            return F_vq.encode_indices(torch.nn.functional.linear(xs[i % n_sets], w_pre, b_pre), w, "vit", prepared=prepared)
        for i in range(3):
            idx_u = unfused(i)
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for i in range(args.steps):
            idx_u = unfused(i)
        u1.record()
        torch.cuda.synchronize()
        unfused_ms = u0.elapsed_time(u1) / args.steps

        # ---- parity, outside the timed region --------------------------------------------------------------------------
        last = (args.steps - 1) % n_sets
        bad = idx != idx_u
        n_bad = int(bad.sum())
        parity = {"rows": T, "rows_differing_from_cublas_linear_plus_quantiser": n_bad,
                  "histogram_sum_equals_tokens": int(torch.bincount(idx, minlength=K).sum()) == T}
        if n_bad:
            # rows that follow the GEMM's rounding: their two best codes must be (nearly) tied
            z_ref = torch.nn.functional.linear(xs[last][bad], w_pre, b_pre)
            zn, en = torch.nn.functional.normalize(z_ref, dim=-1), torch.nn.functional.normalize(w, dim=-1)
            d = (zn * zn).sum(1, keepdim=True) + (en * en).sum(1) - 2 * zn @ en.t()
            two = torch.topk(d, 2, dim=1, largest=False).values
            parity["max_top2_relative_gap_of_those_rows"] = float(((two[:, 1] - two[:, 0]) / two[:, 0].abs().clamp_min(1e-30)).max())
            parity["all_of_them_near_ties"] = parity["max_top2_relative_gap_of_those_rows"] < 1e-4
        sample = xs[last][:8192]
        parity["tensor_core_search_equals_exhaustive_on_sample"] = bool(torch.equal(
            projected.encode_indices_projected(sample, w_pre, b_pre, w, prepared=prepared),
            projected.encode_indices_projected(sample, w_pre, b_pre, w, prepared=prepared, exact_scan=True)))
        z64 = sample.double() @ w_pre.double().t() + b_pre.double()
        scale = sample.abs().double() @ w_pre.abs().double().t() + b_pre.abs().double()
        z_k = projected.quantise_projected(sample, w_pre, b_pre, w, prepared=prepared, return_z=True)[5]
        parity["gemm_max_err_over_sum_abs_terms"] = float(((z_k.double() - z64).abs() / scale).max())
        parity["cublas_fp32_max_err_over_sum_abs_terms"] = float(
            ((torch.nn.functional.linear(sample, w_pre, b_pre).double() - z64).abs() / scale).max())

        # ---- e2e: pinned host rows in, tokens out, copies inside the timed region ---------------------------------------
        e2e = None
        if not args.skip_e2e:
            hx = [torch.randn(T, C).pin_memory() for _ in range(2)]
            hidx = torch.empty(T, dtype=torch.int64).pin_memory()
            dx = torch.empty(T, C, device=dev)

            def host_step(i):
                dx.copy_(hx[i % 2], non_blocking=True)
                hidx.copy_(projected.encode_indices_projected(dx, w_pre, b_pre, w, prepared=prepared), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            host_step(0)
            barrier()
            e_steps = max(3, min(args.steps, 8))
            t0 = time.perf_counter()
            for i in range(e_steps):
                host_step(i)
            e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
            te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e_ms = float(te.item())
            e2e = {"value": world * T / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": T * C * 4, "d2h_bytes_per_step": T * 8,
                   "ms_per_step": e_ms,
                   "note": "pinned host rows (537 MB) -> device -> encode_indices_projected -> int64 tokens -> pinned host; "
                           "wall clock with a stream synchronize per step; PCIe-bound"}
            del hx, dx

    alg = T * C * 4 + D * C * 4
    iface = alg + T * (D * 4 + D * 2 + 8)
    ach = alg / (prep_ms * 1e-3) / 1e9 if prep_ms > 0 else 0.0
    roofline = {"kernel": "vq::k_prequant_prep (Linear 512 -> 32 as 3xTF32 warp MMAs + normalise + fp16 copy, one pass)",
                "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": DRAM_TRAFFIC_PREQUANT_CFG3, "peak_source": peaks["source"], "avg_launch_ms": prep_ms,
                "what_bounds_it": "not HBM yet (0.35 of the measured copy bandwidth): 1007 issued instructions per warp and "
                                  "32-column chunk around 96 warp MMAs (cvt.rna.tf32 is a 4-instruction sequence on sm_100a) "
                                  "with two warps per scheduler -- issue slots 57 % busy, tensor pipe 43 % of the legacy "
                                  "path's 8 cycles per m16n8k8 (profiles/r02b_prequant_ncu.md, r02_ubench_hmma.txt)",
                "algorithmic_bytes_per_launch": alg,
                "algorithmic_bytes_what": "T*C*4 encoder rows + D*C*4 weights (what any implementation reads); the z round "
                                          "trip of the unfused path is gone",
                "interface_bytes_per_launch": iface, "interface_frac": iface / (prep_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if prep_ms > 0 else 0.0,
                "tf32_tflops_issued": 3 * 2.0 * T * C * D / (prep_ms * 1e-3) / 1e12 if prep_ms > 0 else 0.0,
                "search": {"kernel": "vq::tc16::k_dist_tc16", "avg_launch_ms": search_ms,
                           "tflops": 2.0 * K * D * T / (search_ms * 1e-3) / 1e12 if search_ms > 0 else 0.0,
                           "frac_vs_burst_peak": (2.0 * K * D * T / (search_ms * 1e-3) / 1e12 / peaks["tf_burst"]) if search_ms > 0 else 0.0},
                "exact_finish_avg_launch_ms": exact_ms}
    cpu = None
    if rank == 0 and not args.skip_cpu:
        tps, cores, sec, sample = _cpu_port_tokens_per_s(16384, 10.0)
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": DESC, "K": K, "D": D, "C": C, "tokens_per_gpu": T,
                       "parallelism": f"rows sharded over {world} GPU(s), codebook and projection replicated, no collective",
                       "l2": f"encoder rows rotate over {n_sets} resident sets ({n_sets * T * C * 4 >> 20} MiB) > 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "unfused_same_box": {"ms_per_step": unfused_ms, "what": "F.linear (cuBLAS fp32, tf32 off) + encode_indices",
                                 "speedup_of_fused": unfused_ms / ms_step},
            "cpu_baseline": cpu, "parity": parity}), flush=True)
    if world > 1:
        dist.destroy_process_group()
