"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck): single-view forward + backward
on two parity scenes, a batched 2-view step with in-kernel activations, distCUDA2, one fused Adam step.
    compute-sanitizer --tool memcheck python scripts/sanitize_probe.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from gaussianip_b200 import renderer, synthetic
from gaussianip_b200.knn import distCUDA2
from gaussianip_b200.optim import FusedGaussianAdam
from tests import util

dev = torch.device("cuda", 0)
for kw in (dict(P=1500, H=64, W=80, sh_degree=0), dict(P=1000, H=48, W=48, sh_degree=3, scale_boost=3.0)):
    scene = util.humanoid_scene(**kw)
    w = util.loss_weights(scene.H, scene.W)
    got = util.run_gpu(scene, dev, grads=w, requires_grad=True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(got["color"]).all())

P = 2000
cl = synthetic.make_cloud(P, 1, 0)
names = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")
p = {k: getattr(cl, k).clone().to(dev).requires_grad_(True) for k in names}


class Model:
    active_sh_degree = 1
    _opacity = property(lambda s: p["opacity"])
    _scaling = property(lambda s: p["scaling"])
    _rotation = property(lambda s: p["rotation"])
    get_xyz = property(lambda s: p["xyz"])
    get_features = property(lambda s: torch.cat((p["features_dc"], p["features_rest"]), dim=1))


cams = synthetic.ahds_cameras(2, 64, 64, seed=3, device=dev)
out = renderer.render_views(cams, Model(), None, torch.zeros(3, device=dev), fused_activations=True)
(out["render"].sum() + out["depth_3dgs"].sum() + out["alpha_3dgs"].sum()).backward()
torch.cuda.synchronize()
opt = FusedGaussianAdam([{"params": [p[k]], "lr": 1e-3, "name": k} for k in names], eps=1e-15)
opt.step()
d = distCUDA2(p["xyz"].detach())
torch.cuda.synchronize()
assert bool(torch.isfinite(d).all())
print("sanitize probe ok")
