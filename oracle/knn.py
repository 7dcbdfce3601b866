"""TEST INFRASTRUCTURE ONLY — loaders for the two 3-NN checkers (never imported by the product):

* ``oracle_dist2``     the plain-C brute-force restatement, oracle/knn_oracle.c -> oracle/_ref/libknn_oracle.so;
* ``reference_dist2``  the REFERENCE ITSELF (simple-knn compiled from its own sources by oracle/Makefile into
                       oracle/_ref/libsimple_knn_ref.so); needs a GPU; ``None`` when the library was not built.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_C_SO = _HERE / "_ref" / "libknn_oracle.so"
_REF_SO = _HERE / "_ref" / "libsimple_knn_ref.so"


def build(with_reference: bool = True) -> None:
    """Compile the checkers (gcc for the C port; nvcc + the reference sources when they are present)."""
    targets = ["_ref/libknn_oracle.so"] + (["ref"] if with_reference else [])
    r = subprocess.run(["make", "-C", str(_HERE), *targets], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def _c_lib():
    if not _C_SO.exists():
        build(with_reference=False)
    lib = ctypes.CDLL(str(_C_SO))
    lib.knn_oracle_dist2.restype = ctypes.c_int
    lib.knn_oracle_dist2.argtypes = [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib


def oracle_dist2(points, threads: int | None = None) -> np.ndarray:
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
    assert pts.ndim == 2 and pts.shape[1] == 3
    out = np.empty(pts.shape[0], dtype=np.float32)
    threads = threads or min(os.cpu_count() or 1, 32)
    rc = _c_lib().knn_oracle_dist2(pts.shape[0], pts.ctypes.data, out.ctypes.data, threads)
    if rc:
        raise RuntimeError("knn_oracle_dist2 failed")
    return out


def reference_available() -> bool:
    return _REF_SO.exists()


def reference_dist2(points_cuda):
    """Run the reference's own SimpleKNN::knn on a CUDA tensor [P,3] fp32; returns a CUDA tensor [P]."""
    import torch
    lib = ctypes.CDLL(str(_REF_SO))
    lib.ref_simple_knn_dist2.restype = ctypes.c_int
    lib.ref_simple_knn_dist2.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    pts = points_cuda.detach().to(torch.float32).contiguous()
    out = torch.zeros(pts.shape[0], dtype=torch.float32, device=pts.device)
    torch.cuda.synchronize()
    with torch.cuda.device(pts.device):
        rc = lib.ref_simple_knn_dist2(pts.shape[0], pts.data_ptr(), out.data_ptr())
    if rc:
        raise RuntimeError(f"reference simple-knn failed with CUDA error {rc}")
    return out
