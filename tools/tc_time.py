"""Time the nearest-code search alone (vq_profile hooks) for cfg3's shape."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import functional as F_vq, _lib
dev = torch.device("cuda:0")
K, D = int(os.environ.get("K", 8192)), int(os.environ.get("D", 32))
T = int(os.environ.get("T", 262144))
w = vo.make_codebook("vit", K, D, 0).to(dev)
zs = [torch.randn(T, D, device=dev) for _ in range(4)]
prep = F_vq.prepare_codebook(w)
lib = _lib.load()
for i in range(3):
    F_vq.encode_indices(zs[i % 4], w, "vit", prepared=prep)
torch.cuda.synchronize()
lib.vq_profile_begin(1, 0)
n = 10
for i in range(n):
    F_vq.encode_indices(zs[i % 4], w, "vit", prepared=prep)
ms, cnt, launches = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
lib.vq_profile_end(ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(launches))
print(f"VQ_TC_DEBUG={os.environ.get('VQ_TC_DEBUG','0')} K={K} D={D} T={T}: search {ms.value / cnt.value * 1e3:.1f} us per call "
      f"({2.0 * K * D * T / (ms.value / cnt.value * 1e-3) / 1e12:.1f} TFLOP/s algorithmic)")
