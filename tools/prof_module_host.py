import os, sys, cProfile, pstats
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from vq_b200.vitvqgan import Codebook
dev = torch.device("cuda:0")
K, D, T = 8192, 32, 262144
m = Codebook(K, D).to(dev)
z = torch.randn(T // 1024, 1024, D, device=dev, requires_grad=True)
up = torch.randn(T // 1024, 1024, D, device=dev)
one = torch.ones((), device=dev)
def step():
    m.embedding.weight.grad = None; z.grad = None
    with torch.no_grad(): m.embedding.weight.add_(0.0)
    z_q, idx, loss = m(z)
    torch.autograd.backward([z_q, loss], [up, one])
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(22)
