"""CPU checks of the 3-NN oracle (oracle/knn_oracle.c) — SURVEY.md §8 (f3).

The oracle is pinned against the reference itself on the GPU box (tests/test_gpu_knn.py: the reference's
simple-knn, compiled from its own sources into oracle/_ref).  Here, without a GPU: the committed golden
vectors (outputs of that pinned oracle), an independent float64 brute force, and the edge cases."""
import os

import numpy as np
import pytest

from oracle import knn as oknn

GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn_golden.npz")


def brute64(pts):
    p = pts.astype(np.float64)
    d = ((p[:, None, :] - p[None]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    return np.sort(d, axis=1)[:, :3].mean(1)


@pytest.mark.parametrize("P", [4, 5, 33, 257, 1500])
def test_oracle_matches_float64_brute_force(P):
    rng = np.random.default_rng(P)
    pts = (rng.standard_normal((P, 3)) * np.array([1.0, 0.3, 2.0])).astype(np.float32)
    got = oknn.oracle_dist2(pts)
    ref = brute64(pts)
    assert np.allclose(got, ref, rtol=2e-6, atol=1e-12)


def test_oracle_thread_count_invariant():
    rng = np.random.default_rng(1)
    pts = rng.random((3000, 3)).astype(np.float32)
    assert np.array_equal(oknn.oracle_dist2(pts, threads=1), oknn.oracle_dist2(pts, threads=7))


def test_oracle_duplicates_and_tiny_inputs():
    # duplicates count as neighbours at distance 0 (the reference excludes only the query's own index)
    pts = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], dtype=np.float32)
    got = oknn.oracle_dist2(pts)
    assert got[0] == np.float32((0.0 + 1.0 + 4.0) / 3.0)
    assert got[2] == np.float32((1.0 + 1.0 + 5.0) / 3.0)
    # fewer than 4 points: the untouched FLT_MAX slots stay in the sum, as in the reference (:161,185)
    fm = np.finfo(np.float32).max
    two = oknn.oracle_dist2(np.array([[0, 0, 0], [1, 0, 0]], dtype=np.float32))
    with np.errstate(over="ignore"):
        expect = (np.float32(1.0) + np.float32(fm) + np.float32(fm)) / np.float32(3.0)
    assert np.isinf(expect) and np.isinf(two).all()
    three = oknn.oracle_dist2(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32))
    assert np.all(three > 1e37) and np.isfinite(three).all()
    assert oknn.oracle_dist2(np.zeros((0, 3), np.float32)).shape == (0,)


def test_oracle_golden_vectors():
    g = np.load(GOLD)
    for name in ("humanoid", "uniform", "clustered"):
        got = oknn.oracle_dist2(g[name + "_pts"])
        assert np.array_equal(got, g[name + "_dist2"]), name


def test_shim_resolves_and_refuses_cpu():
    import torch
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(8, 3))
    with pytest.raises(ValueError):
        distCUDA2(torch.zeros(8, 2))
