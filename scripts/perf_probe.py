"""Stage breakdown of fwd+bwd at the bench workload (run on the GPU box)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianip_b200 import synthetic, renderer, _lib, rasterizer

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--sh", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mode", default="two_level")
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
cl = synthetic.make_cloud(a.points, a.sh, 0).to(dev)
cl.active_sh_degree = a.sh
for t in (cl.xyz, cl.features_dc, cl.features_rest, cl.scaling, cl.rotation, cl.opacity):
    t.requires_grad_(not a.fwd_only)
cams = synthetic.ahds_cameras(a.views, a.res, a.res, seed=1, device=dev)
bg = torch.zeros(3, device=dev)
g = torch.Generator(device="cpu").manual_seed(2)
wts = [(torch.randn(3, a.res, a.res, generator=g).to(dev), torch.randn(1, a.res, a.res, generator=g).to(dev),
        torch.randn(1, a.res, a.res, generator=g).to(dev)) for _ in range(a.views)]
rasterizer.set_binning_mode(a.mode, dev)

def step():
    loss = 0
    for cam, w in zip(cams, wts):
        out = renderer.render(cam, cl, None, bg)
        if not a.fwd_only:
            loss = loss + (out["render"] * w[0]).sum() + (out["depth_3dgs"] * w[1]).sum() + (out["alpha_3dgs"] * w[2]).sum()
    if not a.fwd_only:
        loss.backward()
        for t in (cl.xyz, cl.features_dc, cl.features_rest, cl.scaling, cl.rotation, cl.opacity):
            t.grad = None

for _ in range(3):
    step()
torch.cuda.synchronize()
ws = rasterizer._workspace(dev)
print("D last view:", ws.last_num_rendered, "d_cap", ws.d_cap, "retries", ws.retries)
_lib.profile_enable(True)
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
for _ in range(a.iters):
    step()
e1.record(); torch.cuda.synchronize(); t1 = time.time()
ms = e0.elapsed_time(e1) / a.iters
print(f"step {ms:.3f} ms  ({a.views / ms * 1e3:.1f} views/s)  wall {(t1 - t0) / a.iters * 1e3:.3f} ms  launches/step {(_lib.launch_count() - l0) / a.iters:.0f}")
prof = _lib.profile_read()
tot = sum(v[0] for v in prof.values())
for k, (m, c) in prof.items():
    if c:
        print(f"  {k:16s} {m / c * 1e3:9.1f} us/call  x{c / a.iters:.0f}/step  {m / a.iters:8.3f} ms/step  {100 * m / tot:5.1f}%")
print(f"  kernels total {tot / a.iters:.3f} ms/step")
_lib.profile_enable(False)
