"""Per-kernel timeline (CUPTI via torch.profiler) of an indices-only encode, default cfg2's shape (VQGAN 8192 x 256,
64 x 256 x 16 x 16 NCHW latents).  GPU only.
    [FORM=vqgan K=8192 D=256 B=64] python tools/encode_timeline.py [steps]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench_inputs
from vq_b200 import functional as F_vq

dev = torch.device("cuda:0")
K, D, B = int(os.environ.get("K", 8192)), int(os.environ.get("D", 256)), int(os.environ.get("B", 64))
FORM = os.environ.get("FORM", "vqgan")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = bench_inputs.make_codebook(FORM, K, D, 0).to(dev)
shape = (B, D, 16, 16) if FORM != "vit" else (B, 256, D)
zs = [torch.randn(*shape, device=dev) for _ in range(8)]          # 8 x 16 MB: rotates through more than nothing, L2 stays warm-ish
prep = F_vq.prepare_codebook(w)
ref = F_vq.encode_indices(zs[0], w, FORM, prepared=prep, exact_scan=True)
got = F_vq.encode_indices(zs[0], w, FORM, prepared=prep)
print("mismatches vs exhaustive search:", int((ref != got).sum()), "of", got.numel())
for i in range(5):
    F_vq.encode_indices(zs[i % 8], w, FORM, prepared=prep)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        F_vq.encode_indices(zs[i % 8], w, FORM, prepared=prep)
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1; a[1] += e.time_range.end - e.time_range.start
busy = sum(a[1] for a in agg.values())
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"{steps} encodes: span {span / steps:.1f} us each, kernels busy {busy / steps:.1f} us, gaps {(span - busy) / steps:.1f} us")
for name, (n, t) in agg.items():
    print(f"{t / steps:8.2f} us  x{n / steps:4.1f}  {name}")
