"""`distCUDA2` — mean squared distance to the 3 nearest neighbours.

Drop-in for ``simple_knn._C.distCUDA2`` (gaussiansplatting/submodules/simple-knn/spatial.cu:15-26):
takes a float [P,3] CUDA tensor, returns float32 [P].  GaussianIP calls it once per model to
initialise the scales (gaussian_model.py:123, gs_renderer.py:387).  The kernels are in
csrc/knn.cu; there is no CPU path — a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import torch

from . import _lib

_scratch = {}


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("distCUDA2 expects a [P,3] tensor")
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 runs on the GPU only (no CPU fallback); got a %s tensor" % points.device)
    lib = _lib.load()
    pts = points.detach().to(torch.float32).contiguous()
    P = pts.shape[0]
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)   # the reference fills with 0 (spatial.cu:21)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        stream = torch.cuda.current_stream()
        need = int(lib.gsb_knn_scratch_bytes(P))
        key = (pts.device.index, stream.cuda_stream)
        buf = _scratch.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.uint8, device=pts.device)
            _scratch[key] = buf
        _lib.check(lib.gsb_knn_dist2(P, pts.data_ptr(), out.data_ptr(), buf.data_ptr(), buf.numel(),
                                     stream.cuda_stream), "gsb_knn_dist2")
    return out


def release_scratch() -> None:
    _scratch.clear()
