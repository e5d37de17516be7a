"""Small shapes through the kernels added this round, as a quick functional pass (it was written for compute-sanitizer memcheck, which this pool does not allow): the clustered D = 256 filter, the
one-launch NCHW prep, the batched undecided-row search, the listed-row scan, narrow token formats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
import bench_inputs as bi
from vq_b200 import functional as F
dev = torch.device("cuda:0")
# D = 256, NCHW (VQGAN form): prep_nchw_fused + cluster filter + rescoring + finish
w = bi.make_codebook("vqgan", 1024, 256, 0).to(dev)
z = torch.randn(3, 256, 16, 16, device=dev)
p = F.prepare_codebook(w)
a = F.encode_indices(z, w, "vqgan", prepared=p)
b = F.encode_indices(z, w, "vqgan", prepared=p, exact_scan=True)
print("vqgan D=256 mismatches", int((a != b).sum()))
zq, idx, loss, hist, stats = F.quantise(z, w, "vqgan", prepared=p)
print("quantise ok", float(loss), int(hist.sum()))
# undecided rows: duplicated codebook (all rows listed: tiled scan), and a few rows only (batched search)
w2 = bi.make_codebook("vit", 256, 256, 1).repeat(4, 1).to(dev)
z2 = torch.randn(700, 256, device=dev)
p2 = F.prepare_codebook(w2)
print("dup codebook mismatches", int((F.encode_indices(z2, w2, "vit", prepared=p2) != F.encode_indices(z2, w2, "vit", prepared=p2, exact_scan=True)).sum()))
w3 = bi.make_codebook("vit", 2048, 256, 2)
w3[5::256] = w3[5]
z3 = torch.randn(600, 256); z3[10:40] = w3[5] * 2 + 0.01 * torch.randn(30, 256)
w3, z3 = w3.to(dev), z3.to(dev)
p3 = F.prepare_codebook(w3)
print("few undecided mismatches", int((F.encode_indices(z3, w3, "vit", prepared=p3) != F.encode_indices(z3, w3, "vit", prepared=p3, exact_scan=True)).sum()))
# D = 32 and narrow tokens
w4 = bi.make_codebook("vit", 1024, 32, 3).to(dev)
z4 = torch.randn(2, 512, 32, device=dev)
i16 = F.encode_indices(z4, w4, "vit", index_dtype=torch.uint16)
i64 = F.encode_indices(z4, w4, "vit")
print("narrow tokens equal", bool(torch.equal(F.convert_tokens(i16, torch.int64), i64)))
d = F.indices_to_embeddings(i16.view(2, -1), w4, "vit")
print("decode ok", tuple(d.shape))
torch.cuda.synchronize()
print("done")
