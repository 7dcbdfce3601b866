"""Generates tests/golden/ref_minicam.npz.  RUN IN THE AUTHORING CONTAINER ONLY (needs /root/reference).

Imports the REFERENCE's own ``MiniCam`` (gs_renderer.py:853-879) and ``getProjectionMatrix``
(gs_renderer.py:829-850) and records their outputs for a few poses.  gs_renderer.py imports packages that are not
in the image (plyfile, kiui) and CUDA-only extensions; they are stubbed with empty modules — MiniCam touches none of
them — and ``.cuda()`` is mapped to the CPU while it runs.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def main():
    for name, attrs in (("plyfile", ("PlyData", "PlyElement")), ("kiui", ()), ("kiui.sh", ("eval_sh", "SH2RGB", "RGB2SH")),
                        ("kiui.mesh", ("Mesh",)), ("diff_gaussian_rasterization", ("GaussianRasterizationSettings",
                                                                                   "GaussianRasterizer")),
                        ("simple_knn", ()), ("simple_knn._C", ("distCUDA2",))):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, object)
        sys.modules[name] = m
    torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("ref_gs_renderer", f"{REF}/gs_renderer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from gaussianip_b200.cameras import look_at_c2w, orbit_position
    out = {"c2w": [], "params": [], "world_view": [], "proj": [], "full_proj": [], "center": []}
    for az, el, dist, fovy, fovx, (h, w), zn, zf in [(20, 5, 2.0, 0.86, 0.86, (512, 512), 0.01, 100.0),
                                                      (-140, -20, 1.4, 1.0, 0.7, (1024, 768), 0.1, 50.0),
                                                      (95, 40, 3.0, 0.5, 0.9, (96, 160), 0.01, 100.0)]:
        c2w = look_at_c2w(orbit_position(az, el, dist)).astype(np.float32)
        cam = mod.MiniCam(c2w.copy(), w, h, fovy, fovx, zn, zf)
        out["c2w"].append(c2w)
        out["params"].append([w, h, fovy, fovx, zn, zf])
        out["world_view"].append(cam.world_view_transform.numpy())
        out["proj"].append(cam.projection_matrix.numpy())
        out["full_proj"].append(cam.full_proj_transform.numpy())
        out["center"].append(cam.camera_center.numpy())
    np.savez_compressed(os.path.join(HERE, "ref_minicam.npz"), **{k: np.stack([np.asarray(x) for x in v]) for k, v in out.items()})
    print("ref_minicam.npz written")


if __name__ == "__main__":
    main()
