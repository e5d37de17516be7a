"""pre_quant / post_quant fusion: fused against unfused, same box, CUDA events.

  encode   F.linear (cuBLAS fp32) + vq encode_indices          vs  encode_indices_projected     (cfg3 rows x 512 features)
  forward  F.linear + quantise (no grad)                       vs  quantise_projected
  step     Linear + Codebook fwd + bwd (autograd)              vs  quantise_projected fwd + bwd
  decode   indices_to_embeddings + F.linear / conv2d 1x1       vs  ProjectedTable.gather        (both forms)
and the prep kernel alone (library event hooks) against the bytes it moves.
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
import bench_inputs as bi
from vq_b200 import functional as F_vq, projected, _lib

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
K, D, C, T = 8192, 32, 512, 262144
N_SETS = 3          # 3 x 537 MB of encoder rows: every pass reads rows that left the 126 MB L2 long ago
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
w = bi.make_codebook("vit", K, D, 0).to(dev)
w_pre = ((torch.rand(D, C, device=dev, generator=g) * 2 - 1) / C ** 0.5)
b_pre = ((torch.rand(D, device=dev, generator=g) * 2 - 1) / C ** 0.5)
w_post = ((torch.rand(C, D, device=dev, generator=g) * 2 - 1) / D ** 0.5)
b_post = ((torch.rand(C, device=dev, generator=g) * 2 - 1) / D ** 0.5)
xs = [torch.randn(T, C, device=dev, generator=g) for _ in range(N_SETS)]
prep = F_vq.prepare_codebook(w)
out = {}


def timed(fn, iters=12, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3       # us


def report(name, unfused, fused):
    out[name] = {"unfused_us": round(unfused, 1), "fused_us": round(fused, 1), "speedup": round(unfused / fused, 2)}
    print(f"{name:34s} unfused {unfused:8.1f} us   fused {fused:8.1f} us   x{unfused / fused:.2f}", flush=True)


with torch.no_grad():
    lin = timed(lambda i: torch.nn.functional.linear(xs[i % N_SETS], w_pre, b_pre))
    print(f"F.linear 262144 x 512 -> 32 alone (cuBLAS fp32): {lin:.1f} us", flush=True)
    out["linear_alone_us"] = round(lin, 1)
    report("encode (indices only)",
           timed(lambda i: F_vq.encode_indices(torch.nn.functional.linear(xs[i % N_SETS], w_pre, b_pre), w, "vit", prepared=prep)),
           timed(lambda i: projected.encode_indices_projected(xs[i % N_SETS], w_pre, b_pre, w, prepared=prep)))
    report("forward (z_q, idx, loss)",
           timed(lambda i: F_vq.quantise(torch.nn.functional.linear(xs[i % N_SETS], w_pre, b_pre), w, "vit", prepared=prep)),
           timed(lambda i: projected.quantise_projected(xs[i % N_SETS], w_pre, b_pre, w, prepared=prep)))

# the prep kernel alone: event pair around it inside the library
torch.cuda.synchronize()
lib.vq_profile_begin(1, 1 << _lib.PROFILE_PREP_TOKENS)
with torch.no_grad():
    for i in range(12):
        projected.encode_indices_projected(xs[i % N_SETS], w_pre, b_pre, w, prepared=prep)
ms, cnt, launches = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
lib.vq_profile_end(ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(launches))
pms, pcnt = ctypes.c_double(0), ctypes.c_int64(0)
lib.vq_profile_slot(_lib.PROFILE_PREP_TOKENS, ctypes.byref(pms), ctypes.byref(pcnt))
if pcnt.value:
    us = pms.value / pcnt.value * 1e3
    alg = T * C * 4 + D * C * 4                       # the encoder rows + W: what any implementation must read
    iface = alg + T * (D * 4 + D * 2 + 8)             # + unit rows fp32 / fp16, |zn|^2, |z| (this implementation's outputs)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    print(f"k_prequant_prep alone: {us:.1f} us; algorithmic {alg / 1e6:.1f} MB -> {alg / us / 1e3:.0f} GB/s; "
          f"interface {iface / 1e6:.1f} MB -> {iface / us / 1e3:.0f} GB/s; 3 x 8.6 GFLOP tf32 -> {3 * 2 * T * C * D / us / 1e6:.0f} TFLOP/s",
          flush=True)
    out["k_prequant_prep"] = {"us": round(us, 1), "algorithmic_bytes": alg, "interface_bytes": iface,
                              "algorithmic_GBps": round(alg / us / 1e3, 1), "interface_GBps": round(iface / us / 1e3, 1),
                              "measured_peaks": peaks}

# training step through autograd
pre = torch.nn.Linear(C, D).to(dev)
with torch.no_grad():
    pre.weight.copy_(w_pre); pre.bias.copy_(b_pre)
wg = w.clone().requires_grad_(True)
ups = [torch.randn(T, D, device=dev, generator=g) for _ in range(2)]
xg = [x.clone().requires_grad_(True) for x in xs[:2]]
prep_a, prep_b = F_vq.prepare_codebook(wg), F_vq.prepare_codebook(wg)


def step_unfused(i):
    z_q, idx, loss, _, _ = F_vq.quantise(pre(xg[i % 2]), wg, "vit", prepared=prep_a, always_refresh=True)
    torch.autograd.backward([z_q, loss], [ups[i % 2], torch.ones_like(loss)])


def step_fused(i):
    z_q, idx, loss, _, _ = projected.quantise_projected(xg[i % 2], pre.weight, pre.bias, wg, prepared=prep_b, always_refresh=True)
    torch.autograd.backward([z_q, loss], [ups[i % 2], torch.ones_like(loss)])


report("step (fwd + bwd, autograd)", timed(step_unfused, 8, 2), timed(step_fused, 8, 2))
del xg, ups

# decode_indices up to the decoder
with torch.no_grad():
    tok = torch.randint(0, K, (256, 1024), device=dev, generator=g)
    table = projected.ProjectedTable(w, w_post, b_post, "vit", prep)
    report("decode ViT 262144 tok -> 512",
           timed(lambda i: torch.nn.functional.linear(F_vq.indices_to_embeddings(tok, w, "vit", prepared=prep, check_indices=False), w_post, b_post)),
           timed(lambda i: table.gather(tok, check_indices=False)))
    tok16 = tok.to(torch.uint16)
    report("  ... from uint16 tokens",
           timed(lambda i: torch.nn.functional.linear(F_vq.indices_to_embeddings(tok16, w, "vit", prepared=prep, check_indices=False), w_post, b_post)),
           timed(lambda i: table.gather(tok16, check_indices=False)))
    Dc = 256
    wc = bi.make_codebook("vqgan", K, Dc, 0).to(dev)
    conv = torch.nn.Conv2d(Dc, Dc, 1).to(dev)
    tokc = torch.randint(0, K, (64, 256), device=dev, generator=g)
    tablec = projected.ProjectedTable(wc, conv.weight, conv.bias, "vqgan")
    report("decode VQGAN cfg2 16384 tok NCHW",
           timed(lambda i: conv(F_vq.indices_to_embeddings(tokc, wc, "vqgan", check_indices=False))),
           timed(lambda i: tablec.gather(tokc, check_indices=False)))
    t_build = timed(lambda i: projected.ProjectedTable(w, w_post, b_post, "vit", prep), 6, 2)
    print(f"building the (8192, 512) table: {t_build:.1f} us (once per codebook / post_quant state)")
    out["table_build_us"] = round(t_build, 1)
print(json.dumps(out))
