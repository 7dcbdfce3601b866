"""Writes tests/golden/knn_golden.npz: seeded point sets and their 3-NN mean squared distances from the C oracle
(oracle/knn_oracle.c).  The oracle itself is pinned bit-for-bit against the reference's compiled simple-knn on the
GPU box (tests/test_gpu_knn.py::test_oracle_matches_reference_build); the GPU test
test_golden_matches_reference_build re-derives these very vectors with the reference.
Run from the repo root:  python tests/golden/make_knn_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import knn as oknn  # noqa: E402
from tests import util  # noqa: E402


def main():
    rng = np.random.default_rng(20240607)
    sets = {}
    sets["humanoid"] = util.humanoid_scene(P=3000, H=64, W=64, sh_degree=0).means3D.detach().cpu().numpy().astype(np.float32)
    sets["uniform"] = rng.random((2500, 3)).astype(np.float32) * np.array([2.0, 1.0, 0.01], dtype=np.float32)
    centres = rng.standard_normal((12, 3)).astype(np.float32) * 3.0
    cl = centres[rng.integers(0, 12, 2000)] + rng.standard_normal((2000, 3)).astype(np.float32) * 0.02
    cl[::97] = cl[1::97][: len(cl[::97])]          # exact duplicates
    sets["clustered"] = cl.astype(np.float32)
    out = {}
    for k, p in sets.items():
        out[k + "_pts"] = p
        out[k + "_dist2"] = oknn.oracle_dist2(p)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "knn_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
