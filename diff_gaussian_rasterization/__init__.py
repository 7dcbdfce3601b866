"""Import shim: lets the reference's unchanged imports

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

(gs_renderer.py:10-13, gaussiansplatting/gaussian_renderer/__init__.py:14) resolve to the
B200-native implementation in ``gaussianip_b200``."""
from gaussianip_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                        rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
