"""Render wrappers with the reference's call signatures and result dictionaries.

* ``render``            — gaussiansplatting/gaussian_renderer/__init__.py:18-104 (also its
                           twin ``render_with_smaller_scale`` :106-193, same logic)
* ``render_deformed``   — same file :195-265 (explicit tensors; dict without depth/alpha)
* ``Renderer.render``   — gs_renderer.py:923-1014 (animation path; clamps the image to [0,1])

``pc`` is anything with the GaussianModel getters (gaussian_model.py:84-107): ``get_xyz``,
``get_features``, ``get_opacity``, ``get_scaling``, ``get_rotation``, ``active_sh_degree``,
``max_sh_degree`` and optionally ``get_covariance``.  ``viewpoint_camera`` is a reference
``Camera`` / ``MiniCam`` or the host-built ones in ``gaussianip_b200.cameras``.

Difference by design: ``tanfovx/tanfovy`` come from Python floats (the reference calls
``math.tan`` on what may be a 0-d CUDA tensor, a device sync per view).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_views

SH_C0 = 0.28209479177387814


def _get(obj, name):
    v = getattr(obj, name)
    return v() if callable(v) else v


def _python_sh_colors(pc, campos, degree):
    """convert_SHs_python branch (gaussian_renderer/__init__.py:73-78) — kept for parity of the
    switch; evaluated with torch ops on the device the model lives on."""
    from .sh import eval_sh
    feats = _get(pc, "get_features")
    shs_view = feats.transpose(1, 2).reshape(feats.shape[0], 3, -1)
    xyz = _get(pc, "get_xyz")
    d = xyz - campos[None, :]
    d = d / d.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(degree, shs_view, d) + 0.5, 0.0)


def _settings(cam, bg_color, scaling_modifier, sh_degree):
    fx, fy = float(cam.FoVx), float(cam.FoVy)
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(fx * 0.5), tanfovy=math.tan(fy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=int(sh_degree), campos=cam.camera_center,
        prefiltered=False, debug=False)


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           screenspace_points=None):
    """``screenspace_points`` (additive, optional): a caller-owned zero-valued [P,3] leaf to use as
    the screen-space gradient carrier instead of a fresh tensor per view, so the per-view
    gradients sum in place (gaussianip_b200.multiview.GradBucket)."""
    xyz = _get(pc, "get_xyz")
    if screenspace_points is None:
        screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    sh_degree = getattr(pc, "active_sh_degree", getattr(pc, "sh_degree", 0))
    rasterizer = GaussianRasterizer(_settings(viewpoint_camera, bg_color, scaling_modifier, sh_degree))
    scales = rotations = cov3D_precomp = None
    if pipe is not None and getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = _get(pc, "get_scaling"), _get(pc, "get_rotation")
    shs = colors_precomp = None
    if override_color is None:
        if pipe is not None and getattr(pipe, "convert_SHs_python", False):
            colors_precomp = _python_sh_colors(pc, viewpoint_camera.camera_center, sh_degree)
        else:
            shs = _get(pc, "get_features")
    else:
        colors_precomp = override_color
    f = lambda t: None if t is None else t.float()
    rendered_image, radii, depth, alpha = rasterizer(
        means3D=xyz.float(), means2D=screenspace_points.float(), shs=f(shs), colors_precomp=colors_precomp,
        opacities=_get(pc, "get_opacity").float(), scales=f(scales), rotations=f(rotations),
        cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "depth_3dgs": depth, "alpha_3dgs": alpha}


render_with_smaller_scale = render


def render_views(cameras, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                 screenspace_points=None, exchange=None, fused_activations: bool = False):
    """All views of one optimisation step in ONE call (additive API; the reference loops
    ``render`` over the views in Python, threestudio/systems/GaussianIP.py:154-159, 305-307).

    Same per-view semantics as ``render``; the differences are host-side only: the activation
    getters (exp / sigmoid / normalize / cat, gaussian_model.py:84-107) run once per step instead
    of once per view, there is one autograd node, and ``viewspace_points.grad`` receives the sum
    over views directly.  Returns the ``render`` dictionary with a leading view axis:
    render [V,3,H,W], depth_3dgs / alpha_3dgs [V,1,H,W], radii_per_view [V,P], radii = max over
    views [P], visibility_filter = radii > 0.  ``exchange``: see rasterize_views (multi-GPU, gradients reduced
    over ranks inside the backward kernel).  ``fused_activations``: hand the model's RAW parameters
    (``pc._opacity``, ``pc._scaling``, ``pc._rotation`` — the reference's attribute names) to the kernels, which apply
    sigmoid / exp / normalize themselves (gaussian_model.py:84-107) and return raw-parameter gradients: same
    values as the getters + autograd to fp32 rounding, ~15 fewer torch kernels per step."""
    xyz = _get(pc, "get_xyz")
    if screenspace_points is None:
        screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    sh_degree = getattr(pc, "active_sh_degree", getattr(pc, "sh_degree", 0))
    settings = [_settings(cam, bg_color, scaling_modifier, sh_degree) for cam in cameras]
    scales = rotations = cov3D_precomp = None
    raw = 0
    if pipe is not None and getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    elif fused_activations:
        scales, rotations = pc._scaling, pc._rotation
        raw |= _lib.RAW_SCALE | _lib.RAW_ROTATION
    else:
        scales, rotations = _get(pc, "get_scaling"), _get(pc, "get_rotation")
    if fused_activations:
        opacities = pc._opacity
        raw |= _lib.RAW_OPACITY
    else:
        opacities = _get(pc, "get_opacity")
    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    elif pipe is not None and getattr(pipe, "convert_SHs_python", False):
        raise ValueError("convert_SHs_python is per view (colours depend on the camera); use render()")
    else:
        shs = _get(pc, "get_features")
    f = lambda t: None if t is None else t.float()
    image, radii_v, depth, alpha = rasterize_views(
        settings, means3D=xyz.float(), means2D=screenspace_points.float(), shs=f(shs),
        colors_precomp=colors_precomp, opacities=opacities.float(), scales=f(scales),
        rotations=f(rotations), cov3D_precomp=cov3D_precomp, exchange=exchange, raw_inputs=raw,
        tanfov_dev=[cam.tanfov_dev for cam in cameras] if all(hasattr(cam, "tanfov_dev") for cam in cameras) else None)
    radii = radii_v.max(dim=0).values
    return {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "radii_per_view": radii_v, "depth_3dgs": depth, "alpha_3dgs": alpha}


def render_deformed(viewpoint_camera, feats, means3D, opacity, scales, rotations, active_sh_degree, pipe,
                    bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """gaussian_renderer/__init__.py:195-265, same positional order (the callers pass ``(cam, feats, means3D, ...)``
    positionally: GaussianIP_anim.py:511, avatar/__init__.py:377).  ``feats`` [P,3] are precomputed colours,
    [P,K,3] spherical harmonics; the dictionary has no depth / alpha entries; ``override_color`` is accepted and
    ignored exactly as the reference does."""
    screenspace_points = torch.zeros_like(means3D, dtype=torch.float32, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    rasterizer = GaussianRasterizer(_settings(viewpoint_camera, bg_color, scaling_modifier, active_sh_degree))
    shs = colors_precomp = None
    if len(feats.shape) == 2:
        colors_precomp = feats
    else:
        shs = feats
    rendered_image, radii, _depth, _alpha = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
        scales=scales, rotations=rotations)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}


class Renderer:
    """gs_renderer.Renderer.render (gs_renderer.py:923-1014) over any object with the getters."""

    def __init__(self, sh_degree=3, white_background=True, gaussians=None, device="cuda"):
        """Positional signature of the reference, ``Renderer(sh_degree, white_background)`` (gs_renderer.py:882).  The
        reference constructs its own GaussianModel; here the model (anything with the getters) is handed in
        with ``gaussians=`` or assigned to ``self.gaussians`` afterwards."""
        self.sh_degree = sh_degree
        self.white_background = white_background
        self.gaussians = gaussians
        self.bg_color = torch.tensor([1, 1, 1] if white_background else [0, 0, 0], dtype=torch.float32,
                                     device=device)

    def render(self, viewpoint_camera, scaling_modifier=1.0, bg_color=None, override_color=None,
               compute_cov3D_python=False, convert_SHs_python=False):
        class _Pipe:
            pass
        pipe = _Pipe()
        pipe.compute_cov3D_python, pipe.convert_SHs_python = compute_cov3D_python, convert_SHs_python
        out = render(viewpoint_camera, self.gaussians, pipe, self.bg_color if bg_color is None else bg_color,
                     scaling_modifier, override_color)
        return {"image": out["render"].clamp(0, 1), "depth": out["depth_3dgs"], "alpha": out["alpha_3dgs"],
                "viewspace_points": out["viewspace_points"], "visibility_filter": out["visibility_filter"],
                "radii": out["radii"]}
