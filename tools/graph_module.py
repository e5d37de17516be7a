import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attention-models_b200"))
import torch
from vq_b200.vitvqgan import Codebook
dev = torch.device("cuda:0")
K, D = 8192, 32
m = Codebook(K, D).to(dev)
m_eager = Codebook(K, D).to(dev)
m_eager.load_state_dict(m.state_dict())
z = torch.randn(256, 1024, D, device=dev, requires_grad=True)
up = torch.randn(256, 1024, D, device=dev)
# graph first (autograd nodes of an earlier eager backward on the default stream would break the capture)
gm = torch.cuda.make_graphed_callables(m, (z.detach().clone().requires_grad_(True),))
z_q, idx, loss = m_eager(z)
(z_q * up).sum().add(loss).backward()
ref = (z_q.detach().clone(), idx.clone(), loss.detach().clone(), z.grad.clone(), m_eager.embedding.weight.grad.clone())
z2 = z.detach().clone().requires_grad_(True)
z_q, idx, loss = gm(z2)
(z_q * up).sum().add(loss).backward()
torch.cuda.synchronize()
print("graphed module == eager:", torch.equal(z_q, ref[0]), torch.equal(idx, ref[1]), torch.equal(loss, ref[2]),
      torch.equal(z2.grad, ref[3]), torch.equal(m.embedding.weight.grad, ref[4]))
def run(mod, zz):
    mod_z_q, _, l = mod(zz)
    torch.autograd.backward([mod_z_q, l], [up, torch.ones((), device=dev)])
for name, mod in (("eager", m_eager), ("graphed", gm)):
    for _ in range(3): run(mod, z2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): run(mod, z2)
    b.record(); torch.cuda.synchronize()
    print(name, f"{a.elapsed_time(b)/20*1e3:.1f} us/step")
