"""B200-native differentiable Gaussian-splatting rasterizer (drop-in for the
``diff_gaussian_rasterization`` operator GaussianIP calls).  See DESIGN.md."""
from .cameras import Camera, MiniCam  # noqa: F401

__all__ = ["Camera", "MiniCam"]
