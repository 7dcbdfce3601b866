"""Synthetic workloads of SURVEY.md §8(d): an SMPL-X-shaped Gaussian cloud and the
reference's camera distributions.  Host-side (CPU torch/numpy); callers move tensors to
the device.  No SMPL-X files exist offline, so the body is a capsule humanoid on an
18-joint / 17-bone OpenPose-style skeleton with the proportions of the one the reference
hard-codes (animation.py:69-91 joints, 119-140 bones), normalised the way the reference
normalises SMPL-X: bbox max-extent 0.6 * 1.1**10, centred, z-up
(threestudio/utils/poser.py:808-821 with scale(-10) at threestudio/systems/GaussianIP.py:128).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from .cameras import Camera, MiniCam, look_at_c2w, orbit_position

SH_C0 = 0.28209479177387814

# A-pose humanoid, y-up before the y/z swap; proportions follow the OpenPose-18 layout
# (nose, neck, r-shoulder, r-elbow, r-wrist, l-shoulder, l-elbow, l-wrist, r-hip, r-knee,
#  r-ankle, l-hip, l-knee, l-ankle, r-eye, l-eye, r-ear, l-ear).
_JOINTS = np.array([
    [0.000, 0.166, 0.054], [0.000, 0.109, -0.005],
    [-0.068, 0.104, -0.006], [-0.114, 0.040, 0.000], [-0.156, -0.029, 0.033],
    [0.060, 0.107, -0.001], [0.104, 0.045, -0.007], [0.154, -0.023, 0.031],
    [0.039, -0.040, 0.002], [0.040, -0.157, -0.002], [0.046, -0.268, -0.002],
    [-0.051, -0.049, 0.002], [-0.049, -0.166, -0.001], [-0.048, -0.275, -0.001],
    [-0.031, 0.194, 0.020], [0.017, 0.196, 0.027], [-0.054, 0.173, -0.013], [0.037, 0.169, -0.009],
], dtype=np.float64)
_BONES = [(0, 1), (1, 2), (2, 3), (3, 4), (1, 5), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10),
          (1, 11), (11, 12), (12, 13), (0, 14), (14, 16), (0, 15), (15, 17)]
# capsule radii before rescale (torso .11, head .09, upper limbs .05, lower .04), in units of
# the 0.6-extent body: scaled by 0.6/1.556 so they are in the same frame as _JOINTS.
_R = 0.6 / 1.556
_BONE_RADIUS = [0.09, 0.05, 0.05, 0.04, 0.05, 0.05, 0.04, 0.11, 0.05, 0.04, 0.11, 0.05, 0.04,
                0.03, 0.03, 0.03, 0.03]
BODY_EXTENT = 0.6 * 1.1 ** 10


@dataclass
class Cloud:
    """Raw (pre-activation) parameters, same fields as GaussianModel
    (gaussiansplatting/scene/gaussian_model.py:36-48)."""
    xyz: torch.Tensor            # [P,3]
    features_dc: torch.Tensor    # [P,1,3]
    features_rest: torch.Tensor  # [P,K-1,3]
    scaling: torch.Tensor        # [P,3] log-scale
    rotation: torch.Tensor       # [P,4] unnormalised (r,x,y,z)
    opacity: torch.Tensor        # [P,1] logit
    sh_degree: int

    def to(self, device):
        return Cloud(*(t.to(device) if torch.is_tensor(t) else t for t in
                       (self.xyz, self.features_dc, self.features_rest, self.scaling,
                        self.rotation, self.opacity)), self.sh_degree)

    # activations: gaussian_model.py:22-34, getters 84-107
    def get_xyz(self):
        return self.xyz

    def get_features(self):
        return torch.cat((self.features_dc, self.features_rest), dim=1)

    def get_opacity(self):
        return torch.sigmoid(self.opacity)

    def get_scaling(self):
        return torch.exp(self.scaling)

    def get_rotation(self):
        return torch.nn.functional.normalize(self.rotation)


def _capsule_surface(rng: np.random.Generator, a, b, r, n):
    """n points uniform on the surface of the capsule with axis a->b and radius r."""
    axis = b - a
    L = np.linalg.norm(axis)
    axis = axis / max(L, 1e-12)
    helper = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
    u = np.cross(axis, helper); u /= np.linalg.norm(u)
    v = np.cross(axis, u)
    area_cyl, area_caps = 2 * math.pi * r * L, 4 * math.pi * r * r
    on_cyl = rng.random(n) < area_cyl / (area_cyl + area_caps)
    phi = rng.random(n) * 2 * math.pi
    t = rng.random(n) * L
    ring = np.cos(phi)[:, None] * u + np.sin(phi)[:, None] * v
    pts_cyl = a + t[:, None] * axis + r * ring
    nrm_cyl = ring
    z = rng.random(n) * 2 - 1
    s = np.sqrt(np.maximum(0, 1 - z * z))
    nrm_cap = s[:, None] * ring + z[:, None] * axis
    base = np.where((z > 0)[:, None], b, a)
    pts_cap = base + r * nrm_cap
    pts = np.where(on_cyl[:, None], pts_cyl, pts_cap)
    nrm = np.where(on_cyl[:, None], nrm_cyl, nrm_cap)
    return pts, nrm


def make_cloud(num_points: int, sh_degree: int = 0, seed: int = 0) -> Cloud:
    rng = np.random.default_rng(seed)
    lens = np.array([np.linalg.norm(_JOINTS[j] - _JOINTS[i]) for i, j in _BONES])
    radii = np.array(_BONE_RADIUS) * _R
    areas = 2 * math.pi * radii * lens + 4 * math.pi * radii ** 2
    counts = np.floor(areas / areas.sum() * num_points).astype(np.int64)
    counts[np.argmax(counts)] += num_points - counts.sum()
    pts, nrms = [], []
    for (i, j), r, n in zip(_BONES, radii, counts):
        p, nr = _capsule_surface(rng, _JOINTS[i], _JOINTS[j], r, int(n))
        pts.append(p); nrms.append(nr)
    pts = np.concatenate(pts); nrms = np.concatenate(nrms)
    perm = rng.permutation(num_points)       # no spatial order in memory, like surface sampling
    pts, nrms = pts[perm], nrms[perm]
    vmin, vmax = pts.min(0), pts.max(0)
    pts = (pts - (vmax + vmin) / 2) * (BODY_EXTENT / np.max(vmax - vmin))
    pts = pts + nrms * rng.normal(0, 0.004, size=(num_points, 1))
    pts = pts[:, [0, 2, 1]]                  # opengl -> blender (y/z swap), z-up
    s0 = math.sqrt(1.5 / num_points)
    scaling = np.log(s0 * np.exp(rng.normal(0, 0.3, size=(num_points, 3))))
    rotation = rng.normal(0, 1, size=(num_points, 4))
    o = rng.uniform(0.05, 0.95, size=(num_points, 1))
    opacity = np.log(o / (1 - o))
    K = (sh_degree + 1) ** 2
    f_dc = (rng.uniform(0, 1, size=(num_points, 1, 3)) - 0.5) / SH_C0
    f_rest = rng.normal(0, 0.05, size=(num_points, K - 1, 3))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32).contiguous()
    return Cloud(t(pts), t(f_dc), t(f_rest), t(scaling), t(rotation), t(opacity), sh_degree)


def ahds_cameras(batch: int, height: int, width: int, seed: int = 1, device="cpu") -> List[Camera]:
    """Stage-1 random orbit cameras: azimuth batch-stratified over [-180,180), elevation
    U(-30,30) deg, distance U(1.3,1.7), fovy U(40,70) deg (configs/exp.yaml:29-34,43-45;
    camera_data.py:349-364, 444-454)."""
    rng = np.random.default_rng(seed)
    cams = []
    for i in range(batch):
        az = (rng.random() + i) / batch * 360.0 - 180.0
        el = rng.uniform(-30, 30)
        dist = rng.uniform(1.3, 1.7)
        fovy = math.radians(rng.uniform(40, 70))
        c2w = look_at_c2w(orbit_position(az, el, dist))
        cams.append(Camera(c2w, fovy, height, width, data_device=device))
    return cams


def vcr_cameras(n_views: int = 64, height: int = 1024, width: int = 1024, device="cpu") -> List[Camera]:
    """Stage-2/3 refinement views: n azimuths linspace(-180,180,n+1)[:n], elevation 17 deg,
    distance 1.5, fovy 70 deg (threestudio/systems/GaussianIP.py:83-87, 232-281)."""
    cams = []
    for az in np.linspace(-180, 180, n_views + 1)[:n_views]:
        c2w = look_at_c2w(orbit_position(float(az), 17.0, 1.5))
        cams.append(Camera(c2w, math.radians(70.0), height, width, data_device=device))
    return cams


def playback_cameras(n_frames: int = 136, height: int = 1024, width: int = 1024, device="cpu") -> List[MiniCam]:
    """animation.py playback: orbit radius 2.5, fovy 50 deg, 1024^2 (animation.py:600-604),
    slowly rotating so successive frames differ."""
    cams = []
    fovy = math.radians(50.0)
    fovx = 2 * math.atan(math.tan(fovy / 2) * width / height)
    for i in range(n_frames):
        c2w = look_at_c2w(orbit_position(-90.0 + 360.0 * i / n_frames, 0.0, 2.5))
        cams.append(MiniCam(c2w, width, height, fovy, fovx, 0.01, 100.0, data_device=device))
    return cams


def playback_sway(xyz: torch.Tensor, frame: int, n_frames: int = 136) -> torch.Tensor:
    """Stand-in for the LBS re-posing of animation.py:375-388: only _xyz changes per frame;
    a smooth +-0.05 sinusoidal sway that grows with distance from the body axis."""
    phase = 2 * math.pi * frame / n_frames
    lateral = xyz[:, 0:1].abs() / (BODY_EXTENT * 0.5)
    off = torch.zeros_like(xyz)
    off[:, 1:2] = 0.05 * math.sin(phase) * lateral
    off[:, 2:3] = 0.05 * math.cos(phase) * lateral * 0.5
    return xyz + off
