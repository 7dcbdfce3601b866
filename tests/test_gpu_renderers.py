"""GPU: the render wrappers behind the reference's call sites — ``render_deformed``
(gaussian_renderer/__init__.py:195-265, called positionally as ``(cam, feats, means3D, ...)`` at
GaussianIP_anim.py:511 and avatar/__init__.py:377), ``Renderer(sh_degree, white_background).render`` with a
``MiniCam`` (gs_renderer.py:853-1014), and repeated backward through one graph."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


class _Model:
    def __init__(self, cl, dev):
        self.cl = cl
        self.active_sh_degree = cl.sh_degree
        self.max_sh_degree = cl.sh_degree
        self.get_xyz = cl.get_xyz().to(dev).requires_grad_(True)
        self.get_features = cl.get_features().to(dev).requires_grad_(True)
        self.get_opacity = cl.get_opacity().to(dev).requires_grad_(True)
        self.get_scaling = cl.get_scaling().to(dev).requires_grad_(True)
        self.get_rotation = cl.get_rotation().to(dev).requires_grad_(True)


def _setup(dev, sh_degree, P=3000, res=96):
    from gaussianip_b200 import synthetic
    cl = synthetic.make_cloud(P, sh_degree, 0)
    cam = synthetic.ahds_cameras(1, res, res, seed=3, device=dev)[0]
    return _Model(cl, dev), cam


@pytest.mark.parametrize("sh_degree", [0, 2])
def test_render_deformed_sh_branch_equals_render(cuda_device, sh_degree):
    from gaussianip_b200 import renderer
    dev = cuda_device
    m, cam = _setup(dev, sh_degree)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    ref = renderer.render(cam, m, None, bg)
    # exactly the reference's positional call: (viewpoint_camera, feats, means3D, opacity, scales, rotations, deg, pipe, bg)
    out = renderer.render_deformed(cam, m.get_features, m.get_xyz, m.get_opacity, m.get_scaling, m.get_rotation,
                                   m.active_sh_degree, None, bg)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}      # no depth / alpha entries
    assert torch.equal(out["render"], ref["render"])
    assert torch.equal(out["radii"], ref["radii"])
    assert torch.equal(out["visibility_filter"], ref["radii"] > 0)
    w = torch.randn(3, cam.image_height, cam.image_width, device=dev, generator=torch.Generator(dev).manual_seed(0))
    leaves = (m.get_xyz, m.get_features, m.get_opacity, m.get_scaling, m.get_rotation)
    g_ref = torch.autograd.grad((ref["render"] * w).sum(), leaves + (ref["viewspace_points"],))
    g_out = torch.autograd.grad((out["render"] * w).sum(), leaves + (out["viewspace_points"],))
    for a, b in zip(g_out, g_ref):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-5 * max(scale, 1e-12)     # same kernels, atomic order only
    assert out["viewspace_points"].shape == m.get_xyz.shape


def test_render_deformed_rgb_branch_is_colors_precomp(cuda_device):
    """feats with two dimensions are precomputed colours (``len(feats.shape) == 2``, __init__.py:233-236)."""
    from gaussianip_b200 import renderer
    dev = cuda_device
    m, cam = _setup(dev, 0)
    bg = torch.zeros(3, device=dev)
    rgb = torch.rand(m.get_xyz.shape[0], 3, device=dev, generator=torch.Generator(dev).manual_seed(1)).requires_grad_(True)
    out = renderer.render_deformed(cam, rgb, m.get_xyz, m.get_opacity, m.get_scaling, m.get_rotation, 0, None, bg,
                                   1.0, None)                               # scaling_modifier, override_color positional
    ref = renderer.render(cam, m, None, bg, override_color=rgb)
    assert torch.equal(out["render"], ref["render"]) and torch.equal(out["radii"], ref["radii"])
    (out["render"].sum()).backward()
    assert rgb.grad is not None and float(rgb.grad.abs().sum()) > 0
    # against the oracle: colours_precomp path
    scene = util.Scene(cam.image_height, cam.image_width, 0, cam.tanfovx, cam.tanfovy, bg.cpu(),
                       cam.world_view_transform.cpu(), cam.full_proj_transform.cpu(), cam.camera_center.cpu(),
                       m.get_xyz.detach().cpu(), m.get_opacity.detach().cpu(), None, rgb.detach().cpu(),
                       m.get_scaling.detach().cpu(), m.get_rotation.detach().cpu(), None)
    o = util.run_oracle(scene)
    assert float((out["render"].detach().cpu() - o["color"]).abs().max()) <= 1e-5
    assert torch.equal(out["radii"].cpu(), o["radii"])


def test_renderer_positional_signature_and_minicam(cuda_device):
    """Renderer(sh_degree, white_background) as at animation.py / gs_renderer.py:882, rendering a MiniCam view."""
    from gaussianip_b200 import renderer
    from gaussianip_b200.cameras import MiniCam, look_at_c2w, orbit_position
    dev = cuda_device
    m, _ = _setup(dev, 1)
    r = renderer.Renderer(1, True)
    assert r.sh_degree == 1 and r.white_background is True and torch.equal(r.bg_color.cpu(), torch.ones(3))
    r.gaussians = m
    c2w = look_at_c2w(orbit_position(40.0, 10.0, 1.6))
    cam = MiniCam(c2w, 128, 96, 0.8, 1.0, 0.01, 100.0, data_device=dev)
    out = r.render(cam)
    assert set(out) == {"image", "depth", "alpha", "viewspace_points", "visibility_filter", "radii"}
    assert out["image"].shape == (3, 96, 128) and out["depth"].shape == (1, 96, 128)
    assert float(out["image"].min()) >= 0.0 and float(out["image"].max()) <= 1.0
    scene = util.Scene(96, 128, 1, cam.tanfovx, cam.tanfovy, torch.ones(3), cam.world_view_transform.cpu(),
                       cam.full_proj_transform.cpu(), cam.camera_center.cpu(), m.get_xyz.detach().cpu(),
                       m.get_opacity.detach().cpu(), m.get_features.detach().cpu(), None,
                       m.get_scaling.detach().cpu(), m.get_rotation.detach().cpu(), None)
    o = util.run_oracle(scene)
    assert float((out["image"].detach().cpu() - o["color"].clamp(0, 1)).abs().max()) <= 1e-5
    assert float((out["alpha"].detach().cpu() - o["alpha"]).abs().max()) <= 1e-5
    assert torch.equal(out["radii"].cpu(), o["radii"])


def test_backward_twice_through_one_graph(cuda_device):
    """retain_graph=True followed by a second backward works, as with the reference operator (its saved buffers
    stay alive with the autograd context)."""
    dev = cuda_device
    scene = util.humanoid_scene(P=2000, H=64, W=64, sh_degree=1)
    w = util.loss_weights(64, 64)
    r = util.run_gpu(scene, dev, requires_grad=True)
    loss = sum((r[k] * wk.to(dev)).sum() for k, wk in zip(("color", "depth", "alpha"), w))
    leaves = [v for v in r["inputs"].values() if v is not None]
    g1 = torch.autograd.grad(loss, leaves, retain_graph=True)
    g2 = torch.autograd.grad(loss, leaves)
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) <= 1e-5 * max(float(a.abs().max()), 1e-12)
