"""One cfg3-sized fused encode (262 144 rows x 512 -> 32, K = 8192) for `ncu -k regex:k_prequant_prep -c 1 --set full`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
import bench_inputs as bi
from vq_b200 import functional as F_vq, projected
dev = torch.device("cuda:0")
K, D, C, T = 8192, 32, 512, 262144
g = torch.Generator(device=dev).manual_seed(0)
w = bi.make_codebook("vit", K, D, 0).to(dev)
w_pre = (torch.rand(D, C, device=dev, generator=g) * 2 - 1) / C ** 0.5
b_pre = (torch.rand(D, device=dev, generator=g) * 2 - 1) / C ** 0.5
x = torch.randn(T, C, device=dev, generator=g)
prep = F_vq.prepare_codebook(w)
for _ in range(2):
    idx = projected.encode_indices_projected(x, w_pre, b_pre, w, prepared=prep)
torch.cuda.synchronize()
print("ok", int(torch.bincount(idx, minlength=K).sum()))
