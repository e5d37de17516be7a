"""Is a step capture-safe?  Capture ShardedQuantiser.step in a CUDA graph, replay it, compare with eager.  GPU only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from oracle import vq_oracle as vo
from vq_b200 import dist as vq_dist
dev = torch.device("cuda:0")
for (K, D, T) in ((8192, 32, 262144), (8192, 32, 2048), (8192, 256, 16384)):
    form = "vit"
    w = vo.make_codebook(form, K, D, 0).to(dev)
    z = torch.randn(T // 256, 256, D, device=dev)
    up = torch.randn_like(z)
    st = vq_dist.ShardedQuantiser(form, 0.25, world_size=1)
    for _ in range(3):
        eager = {k: v.clone() for k, v in st.step(z, up, w).items()}
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        st.step(z, up, w)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = st.step(z, up, w)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    same = all(torch.equal(out[k], eager[k]) for k in ("z_q", "indices", "loss", "grad_z", "grad_weight"))
    def timed(fn, n=50):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    print(f"K={K} D={D} T={T}: graph replay == eager: {same};  eager {timed(lambda: st.step(z, up, w)):.1f} us/step, graph {timed(g.replay):.1f} us/step")
