"""Tensor-core search vs exhaustive fp32 search for one token-major shape (env T, K, D).  GPU only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
import bench_inputs
from vq_b200 import functional as F_vq
dev = torch.device("cuda:0")
K, D, T = int(os.environ.get("K", 8192)), int(os.environ.get("D", 256)), int(os.environ.get("T", 16384))
w = bench_inputs.make_codebook("vit", K, D, 0).to(dev)
z = torch.randn(T, D, device=dev)
prep = F_vq.prepare_codebook(w)
a = F_vq.encode_indices(z, w, "vit", prepared=prep, exact_scan=True)
b = F_vq.encode_indices(z, w, "vit", prepared=prep)
torch.cuda.synchronize()
print(f"K={K} D={D} T={T}: mismatches {int((a != b).sum())}")
