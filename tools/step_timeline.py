"""Per-kernel timeline of the bench step (cfg3 shape) from CUPTI (torch.profiler): in-stream durations, warm L2,
no serialisation -- the numbers ncu's cold-cache launch list cannot give.  GPU only.
    python tools/step_timeline.py [steps]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import vq_oracle as vo
from vq_b200 import dist as vq_dist

dev = torch.device("cuda:0")
K, D, T = int(os.environ.get("K", 8192)), int(os.environ.get("D", 32)), int(os.environ.get("T", 262144))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
FORM = os.environ.get("FORM", "vit")
w = vo.make_codebook(FORM, K, D, 0).to(dev)
shape = (T // 1024, 1024, D) if FORM == "vit" else (T // 256, D, 16, 16)
zs = [torch.randn(*shape, device=dev) for _ in range(4)]
ups = [torch.randn(*shape, device=dev) for _ in range(4)]
st = vq_dist.ShardedQuantiser(FORM, 0.25, world_size=1)
for i in range(5):
    st.step(zs[i % 4], ups[i % 4], w)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        st.step(zs[i % 4], ups[i % 4], w)
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1; a[1] += e.time_range.end - e.time_range.start
busy = sum(a[1] for a in agg.values())
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"{steps} steps: span {span / steps:.1f} us/step, kernels+memsets busy {busy / steps:.1f} us/step, gaps {(span - busy) / steps:.1f} us/step")
for name, (n, t) in agg.items():
    print(f"{t / steps:8.2f} us/step  x{n / steps:4.1f}  {name}")
