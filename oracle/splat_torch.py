"""CPU oracle for the Gaussian-splatting rasterizer hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``gaussianip_b200``) never does: it fails loudly when the CUDA library is missing.

PARITY UNPINNED.  The operator GaussianIP calls (``diff_gaussian_rasterization``, the
ashawkey depth+alpha fork) is an external dependency whose source is NOT under
``/root/reference`` (README.md:22-24 clones it from GitHub at an unpinned HEAD;
``gaussiansplatting/.gitmodules:4-6`` is stale) and the reference ships no tests, golden
vectors or fixtures for it.  This file restates the *published*
algorithm of that dependency and is anchored on what IS in the
reference tree:

* call sites / argument meaning: gaussiansplatting/gaussian_renderer/__init__.py:18-104,
  gs_renderer.py:923-1014;
* Python twins of op sub-stages: ``build_rotation`` / ``build_scaling_rotation`` /
  ``strip_symmetric`` gaussiansplatting/utils/general_utils.py:64-110, ``eval_sh``
  gaussiansplatting/utils/sh_utils.py:26-112 — ``tests/test_oracle_twins.py`` checks the
  oracle against independent restatements of those twins;
* camera algebra: gaussiansplatting/scene/cameras.py:17-51,
  gaussiansplatting/utils/graphics_utils.py:73-99.

Design of the oracle
--------------------
* ``preprocess`` uses ONLY separately-rounded fp32 elementwise ops in a fixed order (no
  matmul, no fused multiply-add), so the CUDA preprocess kernel (compiled with
  ``-fmad=false``) reproduces depth bits, pixel centres, radii and tile rectangles
  bit-for-bit.
* ``bin_and_sort`` builds the 64-bit ``(tile << 32) | depth_bits`` keys and sorts them
  stably (numpy ``kind='stable'``) — the order a stable LSD radix sort produces.
* ``blend_tiles`` is the per-tile front-to-back alpha blend, vectorised as
  ``[256 pixels x n instances]`` with a sequential ``cumprod`` transmittance, written with
  differentiable torch ops so ``autograd`` supplies the backward pass.  Three places where
  the dependency's hand-written backward deliberately differs from the true derivative are
  reproduced explicitly (straight-through 0.99 alpha cap, zero gradient through the
  tan-fov clamp, ``1/(det^2+1e-7)`` in the conic backward).
* ``render_pixel_sequential`` is a scalar pure-Python restatement of the blend loop used to
  pin the vectorised version on small cases.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import NamedTuple, Optional

import numpy as np
import torch

BLOCK_X = 16
BLOCK_Y = 16
NEAR_Z = 0.2            # frustum cull threshold on view-space z (independent of znear)
LOWPASS = 0.3           # screen-space dilation added to the 2D covariance diagonal
ALPHA_CAP = 0.99
ALPHA_MIN = 1.0 / 255.0
T_MIN = 1e-4

# gaussiansplatting/utils/sh_utils.py:26-43
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


class Settings(NamedTuple):
    """Mirror of GaussianRasterizationSettings (gaussian_renderer/__init__.py:36-49)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


@dataclass
class Geom:
    """Per-Gaussian results of preprocess (what the CUDA kernel stores per Gaussian)."""
    visible: torch.Tensor      # [P] bool
    depth: torch.Tensor        # [P] view-space z
    xy: torch.Tensor           # [P,2] pixel centre
    conic: torch.Tensor        # [P,3] inverse 2D covariance (A, B, C)
    opacity: torch.Tensor      # [P]
    rgb: torch.Tensor          # [P,3]
    radii: torch.Tensor        # [P] int32
    rect_min: torch.Tensor     # [P,2] int32 (x, y)
    rect_max: torch.Tensor     # [P,2] int32
    tiles_touched: torch.Tensor  # [P] int32
    cov3d: torch.Tensor        # [P,6]
    cov2d: torch.Tensor        # [P,3] (a, b, c) after low-pass


@dataclass
class Binning:
    keys: np.ndarray           # [D] uint64 sorted
    point_list: np.ndarray     # [D] int64 Gaussian index per sorted instance
    ranges: np.ndarray         # [T,2] int64  [start, end)
    grid: tuple                # (tiles_x, tiles_y)


class _ConicFromCov(torch.autograd.Function):
    """conic = inverse of [[a,b],[b,c]]; backward follows the dependency's cov2D backward,
    which uses 1/(det^2 + 1e-7) instead of 1/det^2 (backward
    preprocess)."""

    @staticmethod
    def forward(ctx, a, b, c):
        det = a * c - b * b
        det_inv = 1.0 / det
        ctx.save_for_backward(a, b, c)
        return c * det_inv, (-b) * det_inv, a * det_inv

    @staticmethod
    def backward(ctx, gA, gB, gC):
        a, b, c = ctx.saved_tensors
        denom = a * c - b * b
        d2 = 1.0 / (denom * denom + 1e-7)
        # d conic / d (a, b, c): conic = (c, -b, a) / denom
        ga = d2 * (-c * c * gA + b * c * gB + (denom - a * c) * gC)
        gc = d2 * (-a * a * gC + a * b * gB + (denom - a * c) * gA)
        gb = d2 * (2 * b * c * gA - (denom + 2 * b * b) * gB + 2 * a * b * gC)
        return ga, gb, gc


def _f(x, dtype):
    return torch.tensor(x, dtype=dtype)


def quat_to_rot(q: torch.Tensor):
    """Rotation matrix entries from a quaternion (r, x, y, z) used AS GIVEN (the kernel
    does not renormalise; callers pass normalised).  general_utils.py:78-98 order."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R00 = 1 - 2 * (y * y + z * z)
    R01 = 2 * (x * y - r * z)
    R02 = 2 * (x * z + r * y)
    R10 = 2 * (x * y + r * z)
    R11 = 1 - 2 * (x * x + z * z)
    R12 = 2 * (y * z - r * x)
    R20 = 2 * (x * z - r * y)
    R21 = 2 * (y * z + r * x)
    R22 = 1 - 2 * (x * x + y * y)
    return ((R00, R01, R02), (R10, R11, R12), (R20, R21, R22))


def cov3d_from_scale_rot(scales: torch.Tensor, rots: torch.Tensor, scale_modifier: float):
    """Sigma = (R S)(R S)^T, upper triangle [xx, xy, xz, yy, yz, zz]
    (general_utils.py:64-110; gaussian_model.py:16-20)."""
    dt = scales.dtype
    mod = _f(scale_modifier, dt)
    s = (mod * scales[:, 0], mod * scales[:, 1], mod * scales[:, 2])
    R = quat_to_rot(rots)
    m = [[R[i][k] * s[k] for k in range(3)] for i in range(3)]

    def dot(i, j):
        return (m[i][0] * m[j][0] + m[i][1] * m[j][1]) + m[i][2] * m[j][2]

    return torch.stack([dot(0, 0), dot(0, 1), dot(0, 2), dot(1, 1), dot(1, 2), dot(2, 2)], dim=1)


def eval_sh_rgb(deg: int, shs: torch.Tensor, dirs: torch.Tensor):
    """SH -> RGB before the +0.5 / clamp.  shs [P,K,3] (coefficient-major, RGB-minor, the
    layout get_features produces, gaussian_model.py:96-100); dirs [P,3] unit.  Same
    polynomial and term order as sh_utils.py:57-99."""
    dt = shs.dtype
    c = lambda v: _f(v, dt)
    sh = lambda k: shs[:, k, :]
    result = c(SH_C0) * sh(0)
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        result = result - c(SH_C1) * y * sh(1) + c(SH_C1) * z * sh(2) - c(SH_C1) * x * sh(3)
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + c(SH_C2[0]) * xy * sh(4) + c(SH_C2[1]) * yz * sh(5)
                      + c(SH_C2[2]) * (2.0 * zz - xx - yy) * sh(6)
                      + c(SH_C2[3]) * xz * sh(7) + c(SH_C2[4]) * (xx - yy) * sh(8))
            if deg > 2:
                result = (result
                          + c(SH_C3[0]) * y * (3.0 * xx - yy) * sh(9)
                          + c(SH_C3[1]) * xy * z * sh(10)
                          + c(SH_C3[2]) * y * (4.0 * zz - xx - yy) * sh(11)
                          + c(SH_C3[3]) * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh(12)
                          + c(SH_C3[4]) * x * (4.0 * zz - xx - yy) * sh(13)
                          + c(SH_C3[5]) * z * (xx - yy) * sh(14)
                          + c(SH_C3[6]) * x * (xx - 3.0 * yy) * sh(15))
    return result


def preprocess(settings: Settings, means3D, opacities, shs=None, colors_precomp=None,
               scales=None, rotations=None, cov3D_precomp=None, means2D=None) -> Geom:
    """Per-Gaussian projection (forward steps 1-10).  Every fp32
    operation is separately rounded and ordered exactly as in csrc/preprocess.cu."""
    dt = means3D.dtype
    H, W = int(settings.image_height), int(settings.image_width)
    V = settings.viewmatrix.to(dt).reshape(16)
    M = settings.projmatrix.to(dt).reshape(16)
    cam = settings.campos.to(dt).reshape(3)
    tanx = _f(settings.tanfovx, dt)
    tany = _f(settings.tanfovy, dt)
    px, py, pz = means3D[:, 0], means3D[:, 1], means3D[:, 2]

    def xf(m, r):   # row r of the column-major 4x4 times [p, 1]
        return ((m[r] * px + m[4 + r] * py) + m[8 + r] * pz) + m[12 + r]

    tx, ty, tz = xf(V, 0), xf(V, 1), xf(V, 2)
    in_front = tz > NEAR_Z
    hx, hy, hw = xf(M, 0), xf(M, 1), xf(M, 3)
    pw = 1.0 / (hw + _f(1e-7, dt))
    projx, projy = hx * pw, hy * pw
    if means2D is not None:
        # means2D is the zero-valued gradient carrier (gaussian_renderer/__init__.py:26-30):
        # its .grad is dL/d(ndc xy).  Adding it here changes no value.
        projx = projx + means2D[:, 0]
        projy = projy + means2D[:, 1]

    if cov3D_precomp is not None:
        cov3d = cov3D_precomp
    else:
        cov3d = cov3d_from_scale_rot(scales, rotations, settings.scale_modifier)
    S00, S01, S02, S11, S12, S22 = [cov3d[:, i] for i in range(6)]
    Sig = ((S00, S01, S02), (S01, S11, S12), (S02, S12, S22))

    # EWA projection
    limx = _f(1.3, dt) * tanx
    limy = _f(1.3, dt) * tany
    txtz, tytz = tx / tz, ty / tz
    cx = torch.minimum(limx, torch.maximum(-limx, txtz))
    cy = torch.minimum(limy, torch.maximum(-limy, tytz))
    # zero gradient through a clamped coordinate (the dependency multiplies dL/dt.x by 0)
    txc = torch.where(cx == txtz, txtz * tz, (cx * tz).detach())
    tyc = torch.where(cy == tytz, tytz * tz, (cy * tz).detach())
    focal_x = _f(W, dt) / (_f(2.0, dt) * tanx)
    focal_y = _f(H, dt) / (_f(2.0, dt) * tany)
    tz2 = tz * tz
    J00 = focal_x / tz
    J02 = -(focal_x * txc) / tz2
    J11 = focal_y / tz
    J12 = -(focal_y * tyc) / tz2
    Wm = lambda i, j: V[i + 4 * j]           # world->view rotation, row i col j
    T0 = [J00 * Wm(0, j) + J02 * Wm(2, j) for j in range(3)]
    T1 = [J11 * Wm(1, j) + J12 * Wm(2, j) for j in range(3)]

    def tsig(Tr, j):
        return (Tr[0] * Sig[0][j] + Tr[1] * Sig[1][j]) + Tr[2] * Sig[2][j]

    U0 = [tsig(T0, j) for j in range(3)]
    U1 = [tsig(T1, j) for j in range(3)]
    c00 = (U0[0] * T0[0] + U0[1] * T0[1]) + U0[2] * T0[2]
    c01 = (U0[0] * T1[0] + U0[1] * T1[1]) + U0[2] * T1[2]
    c11 = (U1[0] * T1[0] + U1[1] * T1[1]) + U1[2] * T1[2]
    a = c00 + _f(LOWPASS, dt)
    b = c01
    c = c11 + _f(LOWPASS, dt)
    det = a * c - b * b
    det_ok = det != 0
    A_, B_, C_ = _ConicFromCov.apply(a, b, c)
    mid = _f(0.5, dt) * (a + c)
    disc = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam = torch.maximum(mid + disc, mid - disc)
    radius_f = torch.ceil(_f(3.0, dt) * torch.sqrt(lam))
    pixx = ((projx + 1.0) * _f(W, dt) - 1.0) * _f(0.5, dt)
    pixy = ((projy + 1.0) * _f(H, dt) - 1.0) * _f(0.5, dt)

    gx = (W + BLOCK_X - 1) // BLOCK_X
    gy = (H + BLOCK_Y - 1) // BLOCK_Y

    def rect(lo, hi, g, blk):
        with torch.no_grad():
            mn = torch.clamp(torch.trunc(lo / blk), 0, g)
            mx = torch.clamp(torch.trunc(hi / blk), 0, g)
            mn = torch.nan_to_num(mn, nan=0.0)
            mx = torch.nan_to_num(mx, nan=0.0)
        return mn.to(torch.int32), mx.to(torch.int32)

    with torch.no_grad():
        rminx, rmaxx = rect(pixx - radius_f, pixx + radius_f + (BLOCK_X - 1), gx, BLOCK_X)
        rminy, rmaxy = rect(pixy - radius_f, pixy + radius_f + (BLOCK_Y - 1), gy, BLOCK_Y)
        tiles = (rmaxx - rminx) * (rmaxy - rminy)
        visible = in_front & det_ok & (tiles > 0)
        radii = torch.where(visible, torch.nan_to_num(radius_f, nan=0.0, posinf=2**30).to(torch.int32),
                            torch.zeros((), dtype=torch.int32))
        tiles = torch.where(visible, tiles, torch.zeros_like(tiles))

    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - cam[None, :]
        ln = torch.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        dirs = d / ln[:, None]
        rgb = torch.clamp_min(eval_sh_rgb(settings.sh_degree, shs, dirs) + 0.5, 0.0)

    return Geom(visible=visible, depth=tz, xy=torch.stack([pixx, pixy], 1),
                conic=torch.stack([A_, B_, C_], 1), opacity=opacities.reshape(-1), rgb=rgb,
                radii=radii, rect_min=torch.stack([rminx, rminy], 1),
                rect_max=torch.stack([rmaxx, rmaxy], 1), tiles_touched=tiles,
                cov3d=cov3d, cov2d=torch.stack([a, b, c], 1))


def bin_and_sort(g: Geom, H: int, W: int) -> Binning:
    """duplicateWithKeys + stable (tile|depth) sort + identifyTileRanges (Appendix A,
    'Binning').  Emission order: Gaussian index ascending, then tile y outer, x inner."""
    gx = (W + BLOCK_X - 1) // BLOCK_X
    gy = (H + BLOCK_Y - 1) // BLOCK_Y
    vis = g.visible.numpy()
    idx = np.nonzero(vis)[0]
    mn = g.rect_min.numpy()[idx].astype(np.int64)
    mx = g.rect_max.numpy()[idx].astype(np.int64)
    wx = mx[:, 0] - mn[:, 0]
    wy = mx[:, 1] - mn[:, 1]
    cnt = wx * wy
    D = int(cnt.sum())
    owner = np.repeat(np.arange(len(idx)), cnt)
    first = np.cumsum(cnt) - cnt
    local = np.arange(D) - np.repeat(first, cnt)
    lx = local % np.maximum(wx[owner], 1)
    ly = local // np.maximum(wx[owner], 1)
    tile = (mn[owner, 1] + ly) * gx + (mn[owner, 0] + lx)
    depth_bits = g.depth.detach().to(torch.float32).numpy().view(np.uint32)[idx][owner].astype(np.uint64)
    keys = (tile.astype(np.uint64) << np.uint64(32)) | depth_bits
    order = np.argsort(keys, kind="stable")
    keys_sorted = keys[order]
    point_list = idx[owner][order]
    tile_sorted = (keys_sorted >> np.uint64(32)).astype(np.int64)
    T = gx * gy
    starts = np.searchsorted(tile_sorted, np.arange(T), side="left")
    ends = np.searchsorted(tile_sorted, np.arange(T), side="right")
    ranges = np.stack([starts, ends], 1)
    ranges[starts == ends] = 0          # untouched tiles keep the zero-initialised range
    return Binning(keys=keys_sorted, point_list=point_list, ranges=ranges, grid=(gx, gy))


@dataclass
class Image:
    color: torch.Tensor        # [3,H,W]
    depth: torch.Tensor        # [1,H,W]
    alpha: torch.Tensor        # [1,H,W]
    n_contrib: torch.Tensor    # [H,W] int32
    final_T: torch.Tensor      # [H,W]
    marginal: torch.Tensor     # [H,W] bool: some threshold decision was within fp noise


def blend_tiles(g: Geom, b: Binning, bg: torch.Tensor, H: int, W: int, chunk: int = 2048,
                margin: float = 2e-5, tile_rows=None) -> Image:
    """Per-tile front-to-back alpha blend (Appendix A, 'Forward blend'), differentiable.
    tile_rows=(phase, stride) blends only tile rows r with r % stride == phase (a bounded
    sample for CPU timing); the other tiles keep the background."""
    dt = g.xy.dtype
    gx, gy = b.grid
    color = torch.zeros(3, gy * BLOCK_Y, gx * BLOCK_X, dtype=dt)
    depth = torch.zeros(gy * BLOCK_Y, gx * BLOCK_X, dtype=dt)
    alpha = torch.zeros(gy * BLOCK_Y, gx * BLOCK_X, dtype=dt)
    finalT = torch.ones(gy * BLOCK_Y, gx * BLOCK_X, dtype=dt)
    ncontrib = torch.zeros(gy * BLOCK_Y, gx * BLOCK_X, dtype=torch.int32)
    marginal = torch.zeros(gy * BLOCK_Y, gx * BLOCK_X, dtype=torch.bool)
    ly, lx = torch.meshgrid(torch.arange(BLOCK_Y), torch.arange(BLOCK_X), indexing="ij")
    ly, lx = ly.reshape(-1), lx.reshape(-1)
    pl = torch.from_numpy(b.point_list)
    half = _f(-0.5, dt)
    color_t, depth_t, alpha_t, T_t = {}, {}, {}, {}
    for t in range(gx * gy):
        s, e = int(b.ranges[t, 0]), int(b.ranges[t, 1])
        if e <= s:
            continue
        ty_, tx_ = divmod(t, gx)
        if tile_rows is not None and ty_ % tile_rows[1] != tile_rows[0]:
            continue
        pxf = (tx_ * BLOCK_X + lx).to(dt)
        pyf = (ty_ * BLOCK_Y + ly).to(dt)
        Tcur = torch.ones(256, dtype=dt)
        Cacc = torch.zeros(256, 3, dtype=dt)
        Dacc = torch.zeros(256, dtype=dt)
        Aacc = torch.zeros(256, dtype=dt)
        done = torch.zeros(256, dtype=torch.bool)
        ncon = torch.zeros(256, dtype=torch.int64)
        marg = torch.zeros(256, dtype=torch.bool)
        for c0 in range(s, e, chunk):
            ids = pl[c0:min(e, c0 + chunk)]
            n = ids.numel()
            xy = g.xy[ids]
            con = g.conic[ids]
            dx = xy[:, 0][None, :] - pxf[:, None]
            dy = xy[:, 1][None, :] - pyf[:, None]
            power = half * ((con[:, 0][None, :] * dx) * dx + (con[:, 2][None, :] * dy) * dy) \
                - (con[:, 1][None, :] * dx) * dy
            G = torch.exp(power)
            a_raw = g.opacity[ids][None, :] * G
            # min(0.99f, x) in CUDA is fminf: a NaN product (NaN opacity) yields 0.99, not NaN
            a_cap = torch.fmin(a_raw.detach(), torch.tensor(ALPHA_CAP, dtype=dt))
            a = torch.where(torch.isnan(a_raw.detach()), a_cap, a_raw + (a_cap - a_raw.detach()))
            valid = (power <= 0) & (a >= ALPHA_MIN)
            a_eff = torch.where(valid, a, torch.zeros((), dtype=dt))
            cp = torch.cumprod(torch.cat([Tcur[:, None], 1 - a_eff], 1), dim=1)
            T_before, test_T = cp[:, :-1], cp[:, 1:]
            stop_here = valid & (test_T < T_MIN)
            stopped = torch.cummax(stop_here.to(torch.int8), dim=1).values.bool()
            blended = valid & ~stopped & ~done[:, None]
            w = torch.where(blended, a * T_before, torch.zeros((), dtype=dt))
            Cacc = Cacc + w @ g.rgb[ids]
            Dacc = Dacc + w @ g.depth[ids]
            Aacc = Aacc + w.sum(1)
            with torch.no_grad():
                pos = torch.arange(1, n + 1)[None, :] + (c0 - s)
                ncon = torch.maximum(ncon, (pos * blended).max(1).values)
                live = ~done[:, None] & ~(stopped & ~stop_here)   # decisions actually taken
                near = ((a.detach() - ALPHA_MIN).abs() < margin * ALPHA_MIN) \
                    | (((test_T.detach() - T_MIN).abs() < 50 * margin * T_MIN) & valid) \
                    | ((power.detach().abs() < 1e-12) & (power.detach() != 0))
                marg |= (near & live).any(1)
            any_stop = stop_here.any(1)
            first = torch.argmax(stop_here.to(torch.int8), dim=1)
            T_stop = cp.gather(1, first[:, None])[:, 0]
            Tnew = torch.where(any_stop, T_stop, cp[:, -1])
            Tcur = torch.where(done, Tcur, Tnew)
            done = done | any_stop
            if bool(done.all()):
                break
        color_t[t], depth_t[t], alpha_t[t], T_t[t] = Cacc, Dacc, Aacc, Tcur
        ys, xs = ty_ * BLOCK_Y, tx_ * BLOCK_X
        ncontrib[ys:ys + BLOCK_Y, xs:xs + BLOCK_X] = ncon.reshape(BLOCK_Y, BLOCK_X).to(torch.int32)
        marginal[ys:ys + BLOCK_Y, xs:xs + BLOCK_X] = marg.reshape(BLOCK_Y, BLOCK_X)
    # assemble differentiably
    tiles = sorted(color_t)
    Hp, Wp = gy * BLOCK_Y, gx * BLOCK_X
    if tiles:
        ty_idx = torch.tensor([t // gx for t in tiles])
        tx_idx = torch.tensor([t % gx for t in tiles])
        Cst = torch.stack([color_t[t] for t in tiles])            # [n,256,3]
        Dst = torch.stack([depth_t[t] for t in tiles])
        Ast = torch.stack([alpha_t[t] for t in tiles])
        Tst = torch.stack([T_t[t] for t in tiles])
        rows = (ty_idx[:, None] * BLOCK_Y + ly[None, :]).reshape(-1)
        cols = (tx_idx[:, None] * BLOCK_X + lx[None, :]).reshape(-1)
        flat = rows * Wp + cols
        color = color.reshape(3, -1).index_copy(1, flat, Cst.reshape(-1, 3).t()).reshape(3, Hp, Wp)
        depth = depth.reshape(-1).index_copy(0, flat, Dst.reshape(-1)).reshape(Hp, Wp)
        alpha = alpha.reshape(-1).index_copy(0, flat, Ast.reshape(-1)).reshape(Hp, Wp)
        finalT = finalT.reshape(-1).index_copy(0, flat, Tst.reshape(-1)).reshape(Hp, Wp)
    color = color + finalT[None] * bg.to(dt).reshape(3, 1, 1)
    return Image(color=color[:, :H, :W], depth=depth[None, :H, :W], alpha=alpha[None, :H, :W],
                 n_contrib=ncontrib[:H, :W], final_T=finalT[:H, :W], marginal=marginal[:H, :W])


def rasterize(settings: Settings, means3D, means2D, opacities, shs=None, colors_precomp=None,
              scales=None, rotations=None, cov3D_precomp=None, return_aux: bool = False, tile_rows=None):
    """The operator: same argument meaning and return tuple as
    GaussianRasterizer.forward (called at gaussian_renderer/__init__.py:85-93)."""
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    H, W = int(settings.image_height), int(settings.image_width)
    g = preprocess(settings, means3D, opacities, shs, colors_precomp, scales, rotations,
                   cov3D_precomp, means2D)
    b = bin_and_sort(g, H, W)
    img = blend_tiles(g, b, settings.bg, H, W, tile_rows=tile_rows)
    if return_aux:
        return img.color, g.radii, img.depth, img.alpha, g, b, img
    return img.color, g.radii, img.depth, img.alpha


def render_pixel_sequential(px: int, py: int, ids, xy, conic, opacity, rgb, depth, bg):
    """Scalar fp32 restatement of the blend loop for ONE pixel (small cases only).  numpy
    float32 scalars, one rounding per operation, same order as csrc/render_fwd.cu."""
    f = np.float32
    T = f(1.0)
    C = [f(0), f(0), f(0)]
    Dz = f(0)
    A = f(0)
    last = 0
    for pos, i in enumerate(ids, start=1):
        dx = f(xy[i][0]) - f(px)
        dy = f(xy[i][1]) - f(py)
        power = f(-0.5) * ((f(conic[i][0]) * dx) * dx + (f(conic[i][2]) * dy) * dy) \
            - (f(conic[i][1]) * dx) * dy
        if power > 0:
            continue
        alpha = min(f(ALPHA_CAP), f(opacity[i]) * f(np.exp(power, dtype=f)))
        if alpha < f(ALPHA_MIN):
            continue
        test_T = T * (f(1) - alpha)
        if test_T < f(T_MIN):
            break
        w = alpha * T
        for ch in range(3):
            C[ch] = C[ch] + f(rgb[i][ch]) * w
        Dz = Dz + f(depth[i]) * w
        A = A + w
        T = test_T
        last = pos
    return [C[ch] + T * f(bg[ch]) for ch in range(3)], Dz, A, last, T
