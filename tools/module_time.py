"""Drop-in module path (Codebook.forward + autograd backward) against the preallocated ShardedQuantiser.step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200.vitvqgan import Codebook
from vq_b200 import dist as vq_dist
dev = torch.device("cuda:0")
K, D, T = 8192, 32, 262144
m = Codebook(K, D).to(dev)
zs = [torch.randn(T // 1024, 1024, D, device=dev, requires_grad=True) for _ in range(4)]
ups = [torch.randn(T // 1024, 1024, D, device=dev) for _ in range(4)]


def mod_step(i):
    m.embedding.weight.grad = None
    zs[i % 4].grad = None
    with torch.no_grad():
        m.embedding.weight.add_(0.0)          # an optimizer step: the weight version changes, the codebook is re-prepared
    z_q, idx, loss = m(zs[i % 4])
    ((z_q * ups[i % 4]).sum() + loss).backward()


st = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1)
w = m.embedding.weight.detach()
for name, fn in (("module + autograd", mod_step), ("ShardedQuantiser.step", lambda i: st.step(zs[i % 4].detach(), ups[i % 4], w))):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    a.record(); t0 = time.perf_counter()
    for i in range(20):
        fn(i)
    host = (time.perf_counter() - t0) / 20 * 1e6
    b.record(); torch.cuda.synchronize()
    print(f"{name:24s}: {a.elapsed_time(b) / 20 * 1e3:7.1f} us/step (host issue {host:.0f} us)")
