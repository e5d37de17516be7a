"""Step time under skewed code usage (collapsed codebooks): tokens drawn around H hot codes.  GPU only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import dist as vq_dist
dev = torch.device("cuda:0")
K, D, T = 8192, 32, 262144
w = vo.make_codebook("vit", K, D, 0).to(dev)
en = torch.nn.functional.normalize(w, dim=-1)
st = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1)
for H in (8192, 1024, 128, 16, 1):
    g = torch.Generator(device=dev).manual_seed(H)
    pick = torch.randint(0, H, (T,), device=dev, generator=g) * (K // H)
    z = (en[pick] + 0.02 * torch.randn(T, D, device=dev, generator=g)).view(T // 1024, 1024, D)
    up = torch.randn_like(z)
    for _ in range(3):
        out = st.step(z, up, w)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = st.step(z, up, w)
    b.record()
    torch.cuda.synchronize()
    used = int((out["histogram"] > 0).sum())
    print(f"{H:5d} hot codes ({used} used): {a.elapsed_time(b) / 10 * 1e3:8.1f} us/step", flush=True)
