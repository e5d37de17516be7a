import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import functional as F_vq
dev = torch.device("cuda:0")
K, D, T = int(os.environ.get("K", 8192)), 256, int(os.environ.get("T", 262144))
w = vo.make_codebook("vit", K, D, 0).to(dev)
z = torch.randn(T // 1024, 1024, D, device=dev)
prep = F_vq.prepare_codebook(w)
for i in range(3):
    idx, hist = F_vq.encode_indices(z, w, "vit", prepared=prep, want_hist=True)
torch.cuda.synchronize()
z_q, idx2, loss, hist, stats = F_vq.quantise(z, w, "vit", prepared=prep)
print("stats near_tie, multi, fallback:", stats[:3].tolist(), "of", T)
