"""GPU: a whole rendering step captured in a CUDA graph (gaussianip_b200.graph.CapturedStep)
replays to the same results as the kernel-by-kernel path, follows new cameras / intrinsics / parameters written
into the fixed-address buffers it was captured on, and re-captures when a view outgrows the instance capacity."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")


def _setup(dev, P=20000, res=160, V=3, sh=1):
    from gaussianip_b200 import multiview, synthetic
    from gaussianip_b200.cameras import CameraBlock
    cl = synthetic.make_cloud(P, sh, 0)
    leaves = {k: getattr(cl, k).clone().to(dev).requires_grad_(True) for k in NAMES}

    class Model:
        active_sh_degree = sh
        _opacity = property(lambda s: leaves["opacity"])
        _scaling = property(lambda s: leaves["scaling"])
        _rotation = property(lambda s: leaves["rotation"])
        get_xyz = property(lambda s: leaves["xyz"])
        get_features = property(lambda s: torch.cat((leaves["features_dc"], leaves["features_rest"]), dim=1))
        get_opacity = property(lambda s: torch.sigmoid(leaves["opacity"]))
        get_scaling = property(lambda s: torch.exp(leaves["scaling"]))
        get_rotation = property(lambda s: torch.nn.functional.normalize(leaves["rotation"]))

    block = CameraBlock(V, res, res, device=dev)
    g = torch.Generator().manual_seed(3)
    w = [torch.randn(V, c, res, res, generator=g).to(dev) for c in (3, 1, 1)]
    vp = multiview.ViewParallel(leaves, P)
    return leaves, Model(), block, w, vp


def _specs(seed, V):
    from gaussianip_b200.cameras import look_at_c2w, orbit_position
    rng = np.random.default_rng(seed)
    return ([look_at_c2w(orbit_position(rng.uniform(-180, 180), rng.uniform(-30, 30), rng.uniform(1.3, 1.7)))
             for _ in range(V)], [math.radians(rng.uniform(40, 70)) for _ in range(V)])


def _step_fn(model, block, w, vp, V):
    from gaussianip_b200 import renderer
    bg = torch.zeros(3, device=block.device)

    def step():
        def rv(views, vsp, exchange=None):
            return renderer.render_views([block.cameras[v] for v in views], model, None, bg, screenspace_points=vsp,
                                         fused_activations=True)

        def loss(views, out):
            return (out["render"] * w[0]).sum() + (out["depth_3dgs"] * w[1]).sum() + (out["alpha_3dgs"] * w[2]).sum()
        out = vp.step_batched(V, rv, loss, views=range(V))
        out["render"] = None
        return out
    return step


def _snapshot(out):
    return {"loss": out["loss"].clone(), "radii": out["radii"].clone(), "vs": out["viewspace_grad"].clone(),
            **{k: out["grads"][k].clone() for k in NAMES}}


def test_graph_replay_equals_eager_and_follows_cameras(cuda_device):
    from gaussianip_b200.graph import CapturedStep
    dev, V = cuda_device, 3
    leaves, model, block, w, vp = _setup(dev, V=V)
    step = _step_fn(model, block, w, vp, V)
    c2ws, fovs = _specs(1, V)
    block.update(c2ws, fovs)
    cs = CapturedStep(step, device=dev, max_forwards=8).capture()
    assert cs.launches_per_capture >= 10 * V
    for seed in (1, 2, 3):                      # seed 1 = the cameras of the capture, then two other sets (other fovs)
        c2ws, fovs = _specs(seed, V)
        block.update(c2ws, fovs)
        got = _snapshot(cs.run())
        block.update(c2ws, fovs)
        ref = _snapshot(step())
        assert torch.equal(got["radii"], ref["radii"]), seed
        assert float(ref["radii"].max()) > 0
        for k in ref:
            if k == "radii":
                continue
            scale = float(ref[k].abs().max())
            assert float((got[k] - ref[k]).abs().max()) <= 2e-5 * max(scale, 1e-12), (seed, k)   # atomic order only
    # new parameter VALUES at the captured addresses are picked up
    with torch.no_grad():
        leaves["xyz"].mul_(0.9)
    got = _snapshot(cs.run())
    ref = _snapshot(step())
    assert torch.equal(got["radii"], ref["radii"])
    assert float((got["loss"] - ref["loss"]).abs()) <= 2e-5 * float(ref["loss"].abs())
    assert cs.captures == 1


def test_graph_recaptures_when_a_view_outgrows_the_capacity(cuda_device):
    from gaussianip_b200 import rasterizer
    from gaussianip_b200.graph import CapturedStep
    dev, V = cuda_device, 2
    leaves, model, block, w, vp = _setup(dev, P=30000, V=V, sh=0)
    step = _step_fn(model, block, w, vp, V)
    c2ws, fovs = _specs(5, V)
    block.update(c2ws, fovs)
    cs = CapturedStep(step, device=dev, max_forwards=8).capture()
    cs.run()
    # blow the splats up: D grows far beyond the capacity the graph was captured with
    with torch.no_grad():
        leaves["scaling"].add_(2.5)
    out = cs.run()                                           # replays, detects the overflow, re-captures, replays
    assert cs.captures >= 2
    got = _snapshot(out)
    ref = _snapshot(step())
    assert torch.equal(got["radii"], ref["radii"])
    assert float((got["loss"] - ref["loss"]).abs()) <= 2e-5 * float(ref["loss"].abs())
    assert rasterizer.stats()["num_rendered"] > 4 * 30000
