"""Throughput over BASELINE.json's configs (GPU only): encode (indices only), forward, fwd+bwd step, decode.
    python tools/sweep.py > profiles/rNN_sweep.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import functional as F_vq, dist as vq_dist

dev = torch.device("cuda:0")


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


CASES = [  # name, form, K, D, shape
    ("cfg1 ViT 2x1024", "vit", 8192, 32, (2, 1024, 32)),
    ("cfg2 VQGAN 64x16x16", "vqgan", 8192, 256, (64, 256, 16, 16)),
    ("cfg3 ViT 256x1024", "vit", 8192, 32, (256, 1024, 32)),
    ("cfg4 ViT 512x1024", "vit", 8192, 32, (512, 1024, 32)),
    ("cfg5 1M K1024 D32", "vit", 1024, 32, (1024, 1024, 32)),
    ("cfg5 1M K4096 D32", "vit", 4096, 32, (1024, 1024, 32)),
    ("cfg5 1M K16384 D32", "vit", 16384, 32, (1024, 1024, 32)),
    ("cfg5 4M K8192 D32", "vit", 8192, 32, (4096, 1024, 32)),
    ("cfg5 16M K8192 D32", "vit", 8192, 32, (16384, 1024, 32)),
    ("cfg5 1M K1024 D256", "vit", 1024, 256, (1024, 1024, 256)),
    ("cfg5 1M K8192 D256", "vit", 8192, 256, (1024, 1024, 256)),
    ("cfg5 1M K16384 D256", "vit", 16384, 256, (1024, 1024, 256)),
    ("cfg5 4M K8192 D256", "vit", 8192, 256, (4096, 1024, 256)),
]
print(f"{'case':24s} {'tokens':>9s} {'encode ms':>10s} {'Mtok/s':>8s} {'TFLOP/s':>8s} {'fwd ms':>8s} {'fwd+bwd ms':>10s} {'Mtok/s':>8s} {'decode ms':>9s} {'GB/s':>6s}")
for name, form, K, D, shape in CASES:
    w = vo.make_codebook(form, K, D, 0).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    z = torch.randn(*shape, device=dev, generator=g)
    up = torch.randn(*shape, device=dev, generator=g)
    T = z.numel() // D
    prep = F_vq.prepare_codebook(w)
    reps = 20 if T <= (1 << 20) else 5
    enc = timed(lambda: F_vq.encode_indices(z, w, form, prepared=prep), reps)
    with torch.no_grad():
        fwd = timed(lambda: F_vq.quantise(z, w, form, prepared=prep), reps)
    st = vq_dist.ShardedQuantiser(form, 0.25, world_size=1)
    step = timed(lambda: st.step(z, up, w), reps)
    idx = F_vq.encode_indices(z, w, form, prepared=prep).view(shape[0], -1)
    dec = timed(lambda: F_vq.indices_to_embeddings(idx, w, form, prepared=prep, check_indices=False), reps)
    print(f"{name:24s} {T:9d} {enc:10.3f} {T / enc / 1e3:8.1f} {2.0 * K * D * T / enc / 1e9:8.1f} {fwd:8.3f} {step:10.3f} "
          f"{T / step / 1e3:8.1f} {dec:9.3f} {T * (4 * D + 8) / dec / 1e6:6.0f}", flush=True)
    del z, up, st, idx
    torch.cuda.empty_cache()
