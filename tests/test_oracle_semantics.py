"""CPU: the vectorised oracle against (i) a scalar pure-Python restatement of the blend loop,
(ii) closed-form cases, (iii) fp64 finite differences of its own forward."""
import math

import numpy as np
import torch

from oracle import splat_torch as O
from tests import util


def test_vectorised_blend_equals_sequential_loop():
    scene = util.humanoid_scene(P=800, H=48, W=48, sh_degree=0, bg=(0.1, 0.5, 0.9), scale_boost=2.0)
    r = util.run_oracle(scene)
    g, b, img = r["geom"], r["binning"], r["image"]
    xy, con = g.xy.numpy(), g.conic.numpy()
    op, rgb, dz = g.opacity.numpy(), g.rgb.numpy(), g.depth.numpy()
    gx = b.grid[0]
    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(60):
        px, py = int(rng.integers(0, 48)), int(rng.integers(0, 48))
        s, e = b.ranges[(py // 16) * gx + px // 16]
        C, Dz, A, last, T = O.render_pixel_sequential(px, py, b.point_list[s:e], xy, con, op, rgb, dz, scene.bg.numpy())
        if bool(img.marginal[py, px]):
            continue
        checked += 1
        assert last == int(img.n_contrib[py, px])
        np.testing.assert_allclose(np.array(C, dtype=np.float32), r["color"][:, py, px].numpy(), atol=2e-6)
        assert abs(float(Dz) - float(r["depth"][0, py, px])) < 5e-6
        assert abs(float(A) - float(r["alpha"][0, py, px])) < 2e-6
        assert abs(float(T) - float(img.final_T[py, px])) < 2e-6
    assert checked > 40


def test_closed_form_single_gaussian():
    from gaussianip_b200.cameras import Camera, look_at_c2w
    H = W = 65
    cam = Camera(look_at_c2w((2.0, 0.0, 0.0)), math.radians(60), H, W, data_device="cpu")
    s, o, z = 0.02, 0.8, 2.0
    st = O.Settings(H, W, cam.tanfovx, cam.tanfovy, torch.tensor([0.1, 0.2, 0.3]), 1.0, cam.world_view_transform,
                    cam.full_proj_transform, 0, cam.camera_center)
    col = torch.tensor([[0.9, 0.5, 0.2]])
    color, radii, depth, alpha = O.rasterize(st, torch.zeros(1, 3), torch.zeros(1, 3), torch.tensor([[o]]),
                                             colors_precomp=col, scales=torch.full((1, 3), s),
                                             rotations=torch.tensor([[1.0, 0, 0, 0]]))
    focal = W / (2 * cam.tanfovx)
    assert int(radii[0]) == math.ceil(3 * math.sqrt((s * focal / z) ** 2 + 0.3))
    c = H // 2
    assert abs(float(alpha[0, c, c]) - o) < 1e-6
    assert abs(float(depth[0, c, c]) - o * z) < 1e-5
    exp = o * col[0] + (1 - o) * torch.tensor([0.1, 0.2, 0.3])
    assert (color[:, c, c] - exp).abs().max().item() < 1e-6
    # alpha is capped at 0.99
    _, _, _, alpha2 = O.rasterize(st, torch.zeros(1, 3), torch.zeros(1, 3), torch.tensor([[1.0]]),
                                  colors_precomp=col, scales=torch.full((1, 3), s),
                                  rotations=torch.tensor([[1.0, 0, 0, 0]]))
    assert abs(float(alpha2[0, c, c]) - 0.99) < 1e-6


def test_exactly_one_of_errors():
    import pytest
    st = O.Settings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3))
    z = lambda *s: torch.zeros(*s)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        O.rasterize(st, z(2, 3), z(2, 3), z(2, 1), scales=z(2, 3), rotations=z(2, 4))
    with pytest.raises(Exception, match="scale/rotation pair"):
        O.rasterize(st, z(2, 3), z(2, 3), z(2, 1), colors_precomp=z(2, 3), scales=z(2, 3))


def test_autograd_matches_fp64_finite_differences():
    """The autograd backward of the oracle is the derivative of its forward (fp64, tiny scene
    without active thresholds or clamps), i.e. the three deliberate deviations are inactive here."""
    torch.manual_seed(0)
    scene = util.humanoid_scene(P=40, H=32, W=32, sh_degree=1, scale_boost=25.0)
    dt = torch.float64
    st = scene.oracle_settings()
    base = {k: getattr(scene, k).to(dt) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    base["opacities"] = base["opacities"] * 0.3 + 0.2          # keep alpha away from 0.99
    w = tuple(t.to(dt) for t in util.loss_weights(32, 32))

    def loss_of(inp):
        color, radii, depth, alpha = O.rasterize(st, inp["means3D"], torch.zeros_like(inp["means3D"]),
                                                 inp["opacities"], shs=inp["shs"], scales=inp["scales"],
                                                 rotations=inp["rotations"])
        return (color * w[0]).sum() + (depth * w[1]).sum() + (alpha * w[2]).sum()

    leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    loss_of(leaves).backward()
    rng = np.random.default_rng(1)
    for k in leaves:
        flat = base[k].reshape(-1)
        worst = 0.0
        for idx in rng.choice(flat.numel(), size=min(6, flat.numel()), replace=False):
            h = 1e-6 * max(1.0, abs(float(flat[idx])))
            vals = []
            for sgn in (+1, -1):
                pert = {kk: vv.clone() for kk, vv in base.items()}
                pert[k].reshape(-1)[idx] += sgn * h
                vals.append(float(loss_of(pert)))
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(leaves[k].grad.reshape(-1)[idx])
            worst = max(worst, abs(fd - an) / max(abs(fd), abs(an), 1e-3))
        assert worst < 2e-3, (k, worst)
