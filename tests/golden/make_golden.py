"""Generates tests/golden/*.npz.  RUN IN THE AUTHORING CONTAINER ONLY (needs /root/reference).

Two kinds of fixtures:
* ``ref_twins.npz`` — outputs of the REFERENCE's own Python functions, imported from
  /root/reference (the only executable pieces of the hot path that exist in the tree):
    gaussiansplatting/utils/sh_utils.py        eval_sh, RGB2SH, SH2RGB
    gaussiansplatting/utils/general_utils.py   build_rotation, build_scaling_rotation, strip_symmetric
    gaussiansplatting/utils/graphics_utils.py  getProjectionMatrix, fov2focal, focal2fov
    gaussiansplatting/scene/cameras.py         Camera (world_view_transform, full_proj_transform, camera_center)
  They hard-code device="cuda"; this script maps those requests to the CPU while it runs them.
* ``oracle_scene_*.npz`` — SELF-GENERATED snapshots of oracle/splat_torch.py on small scenes
  (inputs + every output).  They pin the oracle against drift and give the GPU tests fixed
  vectors that travel to the GPU box; they are NOT reference outputs (the rasterizer the
  reference calls is not in its tree — parity at that boundary is unpinned).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cpu_only():
    """The reference asks for device='cuda' / .cuda(); run the same code on the CPU."""
    for fn_name in ("zeros", "ones", "tensor", "empty"):
        orig = getattr(torch, fn_name)

        def wrapped(*a, __orig=orig, **k):
            if "device" in k:
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, fn_name, wrapped)
    torch.Tensor.cuda = lambda self, *a, **k: self
    _dev = torch.device

    class _Dev:
        def __new__(cls, *a, **k):
            return _dev("cpu")
    torch.device = _Dev


def make_ref_twins():
    sys.path.insert(0, REF)
    sh_utils = _load(f"{REF}/gaussiansplatting/utils/sh_utils.py", "ref_sh_utils")
    gu = _load(f"{REF}/gaussiansplatting/utils/general_utils.py", "ref_general_utils")
    gr = _load(f"{REF}/gaussiansplatting/utils/graphics_utils.py", "gaussiansplatting.utils.graphics_utils")
    sys.modules["gaussiansplatting.utils.graphics_utils"] = gr
    import types
    for pkg in ("gaussiansplatting", "gaussiansplatting.utils"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    cams = _load(f"{REF}/gaussiansplatting/scene/cameras.py", "ref_cameras")
    g = torch.Generator().manual_seed(1234)
    out = {}
    # SH
    dirs = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=1)
    sh = torch.randn(64, 3, 16, generator=g)
    out["sh_dirs"], out["sh_coeffs"] = dirs.numpy(), sh.numpy()
    for deg in range(4):
        out[f"sh_rgb_deg{deg}"] = sh_utils.eval_sh(deg, sh, dirs).numpy()
    rgb = torch.rand(16, 3, generator=g)
    out["rgb"], out["rgb2sh"] = rgb.numpy(), sh_utils.RGB2SH(rgb).numpy()
    out["sh2rgb"] = sh_utils.SH2RGB(sh[:, :, 0]).numpy()
    # covariance
    s = torch.rand(64, 3, generator=g) * 0.1 + 0.001
    q = torch.randn(64, 4, generator=g)
    out["cov_scales"], out["cov_quats"] = s.numpy(), q.numpy()
    out["rot_matrices"] = gu.build_rotation(q).numpy()
    for mod in (1.0, 1.7):
        L = gu.build_scaling_rotation(mod * s, q)
        out[f"cov6_mod{mod}"] = gu.strip_symmetric(L @ L.transpose(1, 2)).numpy()
    # projection + cameras
    fovs = [(0.6, 0.9), (1.2, 1.2), (0.3, 0.2)]
    out["proj_fovs"] = np.array(fovs)
    out["proj_mats"] = np.stack([gr.getProjectionMatrix(0.01, 100.0, fx, fy).numpy() for fx, fy in fovs])
    out["fov2focal"] = np.array([gr.fov2focal(0.8, 512), gr.focal2fov(700.0, 1024)])
    from gaussianip_b200.cameras import look_at_c2w, orbit_position
    c2ws, wv, fp, cc, fx = [], [], [], [], []
    for az, el, dist, fovy, (h, w) in [(30, 10, 1.5, 0.9, (512, 512)), (-120, -25, 1.3, 1.1, (1024, 1024)),
                                        (170, 28, 1.7, 0.75, (96, 160))]:
        c2w = torch.tensor(look_at_c2w(orbit_position(az, el, dist)), dtype=torch.float32)
        cam = cams.Camera(c2w, fovy, h, w)
        c2ws.append(c2w.numpy()); wv.append(cam.world_view_transform.numpy())
        fp.append(cam.full_proj_transform.numpy()); cc.append(cam.camera_center.numpy())
        fx.append([cam.FoVx, cam.FoVy, h, w])
    out["cam_c2w"], out["cam_world_view"] = np.stack(c2ws), np.stack(wv)
    out["cam_full_proj"], out["cam_center"], out["cam_fov_hw"] = np.stack(fp), np.stack(cc), np.array(fx)
    np.savez_compressed(os.path.join(HERE, "ref_twins.npz"), **out)
    print("ref_twins.npz:", sorted(out))


def make_oracle_scenes():
    from tests import util
    specs = {"a": dict(P=600, H=64, W=80, sh_degree=1, bg=(0.2, 0.3, 0.4)),
             "b": dict(P=400, H=48, W=48, sh_degree=3, scale_boost=4.0, bg=(1.0, 1.0, 1.0))}
    for name, kw in specs.items():
        scene = util.humanoid_scene(**kw)
        w = util.loss_weights(scene.H, scene.W)
        r = util.run_oracle(scene, grads=w, requires_grad=True)
        out = {"spec": np.array(repr(kw))}
        for k in ("means3D", "opacities", "shs", "scales", "rotations", "bg", "viewmatrix", "projmatrix", "campos"):
            out["in_" + k] = getattr(scene, k).numpy()
        out["in_scalars"] = np.array([scene.H, scene.W, scene.sh_degree, scene.tanfovx, scene.tanfovy,
                                      scene.scale_modifier], dtype=np.float64)
        out["w_color"], out["w_depth"], out["w_alpha"] = (t.numpy() for t in w)
        for k in ("color", "depth", "alpha"):
            out["out_" + k] = r[k].detach().numpy()
        out["out_radii"] = r["radii"].numpy()
        out["out_keys"] = r["binning"].keys
        out["out_point_list"] = r["binning"].point_list
        out["out_ranges"] = r["binning"].ranges
        out["out_n_contrib"] = r["image"].n_contrib.numpy()
        out["out_marginal"] = r["image"].marginal.numpy()
        for k, v in r["grads"].items():
            if v is not None:
                out["grad_" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"oracle_scene_{name}.npz"), **out)
        print(f"oracle_scene_{name}.npz: D={len(r['binning'].keys)}")


if __name__ == "__main__":
    make_oracle_scenes()      # before the cpu-only patching of torch
    _cpu_only()
    make_ref_twins()
