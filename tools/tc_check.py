"""Debug helper: tensor-core search vs exhaustive fp32 search on the same inputs (GPU only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import functional as F_vq, _lib

dev = torch.device("cuda:0")
cases = [("vit", 8192, 32, (2, 1024, 32)), ("vit", 8192, 32, (256, 1024, 32)), ("vit", 1024, 64, (8, 512, 64)),
         ("vit", 4096, 128, (8, 512, 128)), ("vit", 8192, 256, (16, 1024, 256)), ("vqgan", 8192, 256, (64, 256, 16, 16))]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for form, K, D, shape in cases:
    w = vo.make_codebook(form, K, D, 0).to(dev)
    z = vo.make_latents(shape, 1).to(dev)
    prep = F_vq.prepare_codebook(w)
    res = {}
    for name, exact in (("simt", True), ("tc", False)):
        torch.cuda.synchronize()
        idx, hist = F_vq.encode_indices(z, w, form, prepared=prep, exact_scan=exact, want_hist=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            idx, hist = F_vq.encode_indices(z, w, form, prepared=prep, exact_scan=exact, want_hist=True)
        torch.cuda.synchronize()
        res[name] = (idx, (time.perf_counter() - t0) / 3)
    z_q, idx2, loss, hist, stats = F_vq.quantise(z, w, form, prepared=prep)
    torch.cuda.synchronize()
    T = res["tc"][0].numel()
    mism = int((res["tc"][0] != res["simt"][0]).sum())
    print(f"{form} K={K} D={D} T={T}: mismatches tc-vs-simt={mism}  simt={res['simt'][1]*1e3:.3f} ms  tc={res['tc'][1]*1e3:.3f} ms "
          f"({T / res['tc'][1] / 1e9:.3f} Gtok/s)  stats near_tie={int(stats[0])} multi_cell={int(stats[1])} fallback={int(stats[2])}",
          flush=True)
