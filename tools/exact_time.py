"""Filter / exact-finish kernel times of the bench step (library event hooks), cfg3 shape.  GPU only."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import dist as vq_dist, _lib
dev = torch.device("cuda:0")
K, D, T = 8192, 32, 262144
w = vo.make_codebook("vit", K, D, 0).to(dev)
zs = [torch.randn(T // 1024, 1024, D, device=dev) for _ in range(4)]
ups = [torch.randn(T // 1024, 1024, D, device=dev) for _ in range(4)]
st = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1)
lib = _lib.load()
for i in range(5):
    st.step(zs[i % 4], ups[i % 4], w)
torch.cuda.synchronize()
prof = os.environ.get("NOPROF", "0") == "0"
if prof:
    lib.vq_profile_begin(1, 0)
n = 20
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
a.record()
t0 = time.perf_counter()
for i in range(n):
    st.step(zs[i % 4], ups[i % 4], w)
t_issue = (time.perf_counter() - t0) / n * 1e6
b.record()
torch.cuda.synchronize()
print(f"host issue time {t_issue:.1f} us/step")
if not prof:
    print(f"PDL={os.environ.get('VQ_PDL', '1')} no event hooks: step {a.elapsed_time(b) / n * 1e3:.1f} us")
    sys.exit(0)
ms, cnt, launches = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
lib.vq_profile_end(ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(launches))
ems, ecnt = ctypes.c_double(0), ctypes.c_int64(0)
lib.vq_profile_slot(_lib.PROFILE_EXACT_FINISH, ctypes.byref(ems), ctypes.byref(ecnt))
print(f"{os.environ.get('VQ_B200_LIB', 'default')}: step {a.elapsed_time(b) / n * 1e3:.1f} us  filter {ms.value / cnt.value * 1e3:.1f} us  "
      f"exact+finish {ems.value / ecnt.value * 1e3:.1f} us")
