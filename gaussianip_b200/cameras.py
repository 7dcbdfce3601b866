"""Host-side camera algebra feeding GaussianRasterizationSettings.

Mirrors the two camera classes of the reference so callers can switch unchanged:

* ``Camera``  — gaussiansplatting/scene/cameras.py:17-51 (c2w in OpenGL/NeRF convention,
  ``world_view_transform = (flip . c2w^-1)^T``, ``full_proj_transform = view @ proj``,
  ``camera_center = view^-1[3, :3]``; znear 0.01 / zfar 100).
* ``MiniCam`` — gs_renderer.py:853-879 (same flip; ``camera_center = -c2w[:3, 3]`` as the
  reference does, which is what its SH evaluation sees).

Difference by design: the 4x4 algebra is done on the HOST in float64 and
uploaded once, instead of two ``torch.inverse`` launches on the GPU per view; ``tanfovx`` /
``tanfovy`` are plain Python floats, so building settings never synchronises the device.
Matrices keep the reference's row-vector convention: the flat memory is the column-major
world->view matrix, ``x_view = m[0]*x + m[4]*y + m[8]*z + m[12]``.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def fov2focal(fov: float, pixels: int) -> float:
    """gaussiansplatting/utils/graphics_utils.py:95-96."""
    return pixels / (2.0 * math.tan(fov / 2.0))


def focal2fov(focal: float, pixels: int) -> float:
    """gaussiansplatting/utils/graphics_utils.py:98-99."""
    return 2.0 * math.atan(pixels / (2.0 * focal))


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """Perspective matrix of graphics_utils.py:73-93 (symmetric frustum, z_sign = +1)."""
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 1.0 / math.tan(fovx / 2.0)
    P[1, 1] = 1.0 / math.tan(fovy / 2.0)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def _w2c_from_c2w(c2w) -> np.ndarray:
    c2w = np.asarray(c2w.detach().cpu() if torch.is_tensor(c2w) else c2w, dtype=np.float64)
    w2c = np.linalg.inv(c2w)
    w2c[1:3, :3] *= -1.0          # OpenGL -> COLMAP axis flip ("rectify", cameras.py:25-27)
    w2c[:3, 3] *= -1.0
    return w2c


class Camera:
    """Drop-in for gaussiansplatting.scene.cameras.Camera (cameras.py:17-51)."""

    def __init__(self, c2w, FoVy, height, width, trans=None, scale=1.0, data_device="cuda"):
        FoVy = float(FoVy)
        self.FoVx = focal2fov(fov2focal(FoVy, height), width)
        self.FoVy = FoVy
        self.image_height = int(height)
        self.image_width = int(width)
        self.zfar = 100.0
        self.znear = 0.01
        self.trans = torch.zeros(3) if trans is None else trans.float()
        self.scale = scale
        self.data_device = torch.device(data_device)
        w2c = _w2c_from_c2w(c2w)
        # contiguous copies: torch.tensor keeps numpy's transposed strides, and a non-contiguous matrix would cost
        # every render call a device copy before its pointer can be handed to the kernels
        view = np.ascontiguousarray(w2c.T)
        proj = np.ascontiguousarray(projection_matrix(self.znear, self.zfar, self.FoVx, self.FoVy).T)
        full = view.astype(np.float32).astype(np.float64) @ proj.astype(np.float32).astype(np.float64)
        center = np.linalg.inv(view)[3, :3]
        dev = self.data_device
        self.world_view_transform = torch.tensor(view, dtype=torch.float32).to(dev)
        self.projection_matrix = torch.tensor(proj, dtype=torch.float32).to(dev)
        self.full_proj_transform = torch.tensor(full, dtype=torch.float32).to(dev)
        self.camera_center = torch.tensor(center, dtype=torch.float32).to(dev)

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)


def _camera_rows(c2ws, fovys, height, width):
    """Host algebra for n cameras at once (one LAPACK call for all inversions): rows of 53 floats
    [world_view 16 | projection 16 | full_proj 16 | centre 3 | tanfovx, tanfovy] with values identical to
    Camera(c2w, fovy, height, width); also returns (FoVx, FoVy) per camera."""
    n = len(c2ws)
    fovys = [float(f) for f in fovys]
    c2w = np.stack([np.asarray(c.detach().cpu() if torch.is_tensor(c) else c, dtype=np.float64) for c in c2ws])
    w2c = np.linalg.inv(c2w)
    w2c[:, 1:3, :3] *= -1.0
    w2c[:, :3, 3] *= -1.0
    view = np.ascontiguousarray(np.swapaxes(w2c, 1, 2))
    fovxs = [focal2fov(fov2focal(f, height), width) for f in fovys]
    proj = np.stack([projection_matrix(0.01, 100.0, fx, fy).T for fx, fy in zip(fovxs, fovys)])
    full = view.astype(np.float32).astype(np.float64) @ proj.astype(np.float32).astype(np.float64)
    center = np.linalg.inv(view)[:, 3, :3]
    tanfov = np.array([[math.tan(fx * 0.5), math.tan(fy * 0.5)] for fx, fy in zip(fovxs, fovys)])
    rows = np.concatenate((view.reshape(n, 16), proj.reshape(n, 16), full.reshape(n, 16), center, tanfov), axis=1)
    return rows.astype(np.float32), fovxs, fovys


ROW_FLOATS = 53


def _bind_rows(cams, dev_buf):
    for i, cam in enumerate(cams):
        cam.world_view_transform = dev_buf[i, 0:16].view(4, 4)
        cam.projection_matrix = dev_buf[i, 16:32].view(4, 4)
        cam.full_proj_transform = dev_buf[i, 32:48].view(4, 4)
        cam.camera_center = dev_buf[i, 48:51]
        cam.tanfov_dev = dev_buf[i, 51:53]     # device copy of (tanfovx, tanfovy), see rasterize_views(tanfov_dev=)


def _blank_cameras(fovxs, fovys, height, width, device):
    cams = []
    for fovx, fovy in zip(fovxs, fovys):
        cam = Camera.__new__(Camera)
        cam.FoVx, cam.FoVy = fovx, fovy
        cam.image_height, cam.image_width = int(height), int(width)
        cam.zfar, cam.znear, cam.trans, cam.scale = 100.0, 0.01, torch.zeros(3), 1.0
        cam.data_device = torch.device(device)
        cams.append(cam)
    return cams


def cameras_from_c2w(c2ws, fovys, height, width, device="cuda"):
    """Build the Camera objects of one step with ONE pinned staging buffer and ONE asynchronous
    host->device copy (the reference builds each camera separately with two GPU inversions,
    threestudio/systems/GaussianIP.py:155; a pageable per-tensor upload would also block the host
    until the stream drains).  Values are identical to Camera(c2w, fovy, height, width)."""
    rows, fovxs, fovys = _camera_rows(c2ws, fovys, height, width)
    stage = torch.from_numpy(rows)
    if torch.device(device).type == "cuda":
        stage = stage.pin_memory()
    cams = _blank_cameras(fovxs, fovys, height, width, device)
    dev_buf = stage.to(device, non_blocking=True)
    _bind_rows(cams, dev_buf)
    for cam in cams:
        cam._stage = stage            # keep the pinned buffer alive until the copy has run
    return cams


class CameraBlock:
    """n cameras whose matrices live at FIXED device addresses (one [n, 53] tensor), refreshed per step with one
    asynchronous copy from a ring of pinned staging buffers.  A step captured in a CUDA graph
    (gaussianip_b200.graph.CapturedStep) reads its cameras from here, so replaying it with new cameras is
    ``block.update(c2ws, fovys)`` followed by the replay; the intrinsics travel in the same rows
    (``cam.tanfov_dev``) because kernel arguments passed by value are frozen at capture time."""

    def __init__(self, n: int, height: int, width: int, device="cuda", ring: int = 4):
        self.n, self.height, self.width = int(n), int(height), int(width)
        self.device = torch.device(device)
        self.dev = torch.zeros(self.n, ROW_FLOATS, dtype=torch.float32, device=self.device)
        self.stage = [torch.zeros(self.n, ROW_FLOATS, dtype=torch.float32) for _ in range(ring)]
        if self.device.type == "cuda":
            self.stage = [t.pin_memory() for t in self.stage]
        self.slot = 0
        self.cameras = _blank_cameras([1.0] * self.n, [1.0] * self.n, height, width, self.device)
        _bind_rows(self.cameras, self.dev)

    def update(self, c2ws, fovys):
        """Host algebra + ONE async H2D on the current stream.  The host may run at most ``ring - 1`` updates
        ahead of the device (the caller's per-step validation / loss read-back bounds that)."""
        if len(c2ws) != self.n:
            raise ValueError(f"expected {self.n} cameras")
        rows, fovxs, fovys = _camera_rows(c2ws, fovys, self.height, self.width)
        st = self.stage[self.slot]
        self.slot = (self.slot + 1) % len(self.stage)
        st.copy_(torch.from_numpy(rows))
        self.dev.copy_(st, non_blocking=True)
        for cam, fx, fy in zip(self.cameras, fovxs, fovys):
            cam.FoVx, cam.FoVy = fx, fy
        return self.cameras

    def nbytes(self) -> int:
        return self.n * ROW_FLOATS * 4


class MiniCam:
    """Drop-in for gs_renderer.MiniCam (gs_renderer.py:853-879)."""

    def __init__(self, c2w, width, height, fovy, fovx, znear, zfar, data_device="cuda"):
        self.image_width = int(width)
        self.image_height = int(height)
        self.FoVy = float(fovy)
        self.FoVx = float(fovx)
        self.znear = znear
        self.zfar = zfar
        c2w = np.asarray(c2w, dtype=np.float64)
        w2c = _w2c_from_c2w(c2w)
        view = np.ascontiguousarray(w2c.T)
        proj = np.ascontiguousarray(projection_matrix(znear, zfar, self.FoVx, self.FoVy).T)
        full = view.astype(np.float32).astype(np.float64) @ proj.astype(np.float32).astype(np.float64)
        dev = torch.device(data_device)
        self.world_view_transform = torch.tensor(view, dtype=torch.float32).to(dev)
        self.projection_matrix = torch.tensor(proj, dtype=torch.float32).to(dev)
        self.full_proj_transform = torch.tensor(full, dtype=torch.float32).to(dev)
        self.camera_center = -torch.tensor(c2w[:3, 3], dtype=torch.float32).to(dev)

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)


def look_at_c2w(position, center=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """c2w of a camera at ``position`` looking at ``center``, built as the reference's
    random-camera collate does (threestudio/data/camera_data.py:444-454):
    columns = right, up, -lookat, position."""
    position = np.asarray(position, dtype=np.float64)
    look = np.asarray(center, dtype=np.float64) - position
    look /= np.linalg.norm(look)
    right = np.cross(look, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, look)
    upv /= np.linalg.norm(upv)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, -look, position
    return c2w


def orbit_position(azimuth_deg: float, elevation_deg: float, distance: float) -> np.ndarray:
    """z-up orbit position (camera_data.py:356-364)."""
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    return np.array([distance * math.cos(el) * math.cos(az),
                     distance * math.cos(el) * math.sin(az),
                     distance * math.sin(el)])
