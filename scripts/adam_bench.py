"""Fused Adam + densification statistics vs torch.optim.Adam on the same tensors (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianip_b200.optim import FusedGaussianAdam
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shapes = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (K - 1, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
def groups():
    return [{"params": [torch.randn(P, *s, device=dev).requires_grad_(True)], "lr": 1e-3, "name": n} for n, s in shapes.items()]
def grads(gs):
    for g in gs: g["params"][0].grad = torch.randn_like(g["params"][0])
def timeit(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
a, b = groups(), groups(); grads(a); grads(b)
ours = FusedGaussianAdam(a, eps=1e-15); ref = torch.optim.Adam(b, lr=0.0, eps=1e-15)
vg = torch.randn(P, 3, device=dev); radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32)
acc, den, mr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
def ref_step():
    ref.step()
    f = radii > 0                                   # the reference's boolean-mask formulation
    mr[f] = torch.max(mr[f], radii[f].float()); acc[f] += torch.norm(vg[f, :2], dim=-1, keepdim=True); den[f] += 1
n_el = sum(g["params"][0].numel() for g in a)
t_ours = timeit(lambda: ours.step(densify=(acc, den, mr, vg, radii)))
t_ref = timeit(ref_step)
byt = 28 * n_el + 28 * P
print(f"P={P} K={K}: fused {t_ours:.1f} us ({byt / t_ours / 1e3:.0f} GB/s, {byt / t_ours / 1e3 / 6550.4:.2f} of measured HBM peak)"
      f"  torch.optim.Adam + masked stats {t_ref:.1f} us  speed-up {t_ref / t_ours:.1f}x")
