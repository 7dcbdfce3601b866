"""Shared scene builders for the parity tests (oracle vs CUDA path on identical inputs)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from gaussianip_b200 import synthetic
from oracle import splat_torch as O


@dataclass
class Scene:
    H: int
    W: int
    sh_degree: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    means3D: torch.Tensor
    opacities: torch.Tensor
    shs: Optional[torch.Tensor]
    colors: Optional[torch.Tensor]
    scales: Optional[torch.Tensor]
    rotations: Optional[torch.Tensor]
    cov3D: Optional[torch.Tensor]
    scale_modifier: float = 1.0

    def oracle_settings(self):
        return O.Settings(self.H, self.W, self.tanfovx, self.tanfovy, self.bg, self.scale_modifier,
                          self.viewmatrix, self.projmatrix, self.sh_degree, self.campos)

    def inputs(self, device="cpu", requires_grad=False):
        d = {}
        for k in ("means3D", "opacities", "shs", "colors", "scales", "rotations", "cov3D"):
            t = getattr(self, k)
            if t is not None:
                t = t.detach().clone().to(device)
                if requires_grad:
                    t.requires_grad_(True)
            d[k] = t
        d["means2D"] = torch.zeros_like(d["means3D"], requires_grad=requires_grad)
        return d


def humanoid_scene(P=4000, H=128, W=128, sh_degree=0, seed=0, cam_seed=1, bg=(0.0, 0.0, 0.0),
                   scale_boost=1.0, K=None, precomp_color=False, precomp_cov=False, scale_modifier=1.0,
                   camera=None) -> Scene:
    cl = synthetic.make_cloud(P, sh_degree if K is None else int(math.isqrt(K)) - 1, seed)
    cl.sh_degree = sh_degree
    cam = camera if camera is not None else synthetic.ahds_cameras(1, H, W, seed=cam_seed)[0]
    scales = cl.get_scaling() * scale_boost
    rots = cl.get_rotation()
    shs = cl.get_features()
    colors = None
    cov3D = None
    if precomp_color:
        colors = torch.rand(P, 3, generator=torch.Generator().manual_seed(seed + 7))
        shs = None
    if precomp_cov:
        cov3D = O.cov3d_from_scale_rot(scales, rots, scale_modifier)
        scales = rots = None
    return Scene(H, W, sh_degree, cam.tanfovx, cam.tanfovy, torch.tensor(bg, dtype=torch.float32),
                 cam.world_view_transform.cpu(), cam.full_proj_transform.cpu(), cam.camera_center.cpu(),
                 cl.get_xyz(), cl.get_opacity(), shs, colors, scales, rots, cov3D, scale_modifier)


def run_oracle(scene: Scene, grads=None, requires_grad=False):
    inp = scene.inputs("cpu", requires_grad)
    out = O.rasterize(scene.oracle_settings(), inp["means3D"], inp["means2D"], inp["opacities"], shs=inp["shs"],
                      colors_precomp=inp["colors"], scales=inp["scales"], rotations=inp["rotations"],
                      cov3D_precomp=inp["cov3D"], return_aux=True)
    color, radii, depth, alpha, g, b, img = out
    res = dict(color=color, radii=radii, depth=depth, alpha=alpha, geom=g, binning=b, image=img, inputs=inp)
    if grads is not None:
        gc, gd, ga = grads
        loss = (color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum()
        loss.backward()
        res["grads"] = {k: (v.grad if v is not None else None) for k, v in inp.items()}
    return res


def run_gpu(scene: Scene, device, grads=None, requires_grad=False, debug=False, mode=None):
    from gaussianip_b200 import rasterizer as R
    if mode is not None:
        R.set_binning_mode(mode, device)
    inp = scene.inputs(device, requires_grad)
    rs = R.GaussianRasterizationSettings(
        image_height=scene.H, image_width=scene.W, tanfovx=scene.tanfovx, tanfovy=scene.tanfovy,
        bg=scene.bg.to(device), scale_modifier=scene.scale_modifier, viewmatrix=scene.viewmatrix.to(device),
        projmatrix=scene.projmatrix.to(device), sh_degree=scene.sh_degree, campos=scene.campos.to(device),
        prefiltered=False, debug=debug)
    rast = R.GaussianRasterizer(rs)
    if requires_grad:
        inp["means2D"].retain_grad()
    color, radii, depth, alpha = rast(means3D=inp["means3D"], means2D=inp["means2D"], shs=inp["shs"],
                                      colors_precomp=inp["colors"], opacities=inp["opacities"],
                                      scales=inp["scales"], rotations=inp["rotations"], cov3D_precomp=inp["cov3D"])
    res = dict(color=color, radii=radii, depth=depth, alpha=alpha, inputs=inp)
    if grads is not None:
        gc, gd, ga = (t.to(device) for t in grads)
        loss = (color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum()
        loss.backward()
        res["grads"] = {k: (v.grad if v is not None else None) for k, v in inp.items()}
    if mode is not None:
        R.set_binning_mode("two_level", device)
    return res


def loss_weights(H, W, seed=2):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g), torch.randn(1, H, W, generator=g))
