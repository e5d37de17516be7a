"""exact/finish kernel time by mode (library event hooks): indices only, forward without segment sums, full training forward."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
import bench_inputs as bi
from vq_b200 import functional as F_vq, _lib
dev = torch.device("cuda:0")
K, D, T = 8192, 32, 262144
w = bi.make_codebook("vit", K, D, 0).to(dev)
zs = [torch.randn(T // 1024, 1024, D, device=dev) for _ in range(4)]
lib = _lib.load()
prep = F_vq.prepare_codebook(w)
wg = w.clone().requires_grad_(True)

def run(mode, i):
    if mode == "indices":
        F_vq.encode_indices(zs[i % 4], w, "vit", prepared=prep)
    elif mode == "forward":
        with torch.no_grad():
            F_vq.quantise(zs[i % 4], w, "vit", prepared=prep)
    else:
        F_vq.quantise(zs[i % 4].requires_grad_(True), wg, "vit")

for mode in ("indices", "forward", "train"):
    for i in range(4):
        run(mode, i)
    torch.cuda.synchronize()
    lib.vq_profile_begin(1, 0)
    for i in range(12):
        run(mode, i)
    ms, cnt, launches = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
    lib.vq_profile_end(ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(launches))
    ems, ecnt = ctypes.c_double(0), ctypes.c_int64(0)
    lib.vq_profile_slot(_lib.PROFILE_EXACT_FINISH, ctypes.byref(ems), ctypes.byref(ecnt))
    print(f"{mode:8s}: filter {ms.value / cnt.value * 1e3:.1f} us  exact+finish {ems.value / ecnt.value * 1e3:.1f} us")
