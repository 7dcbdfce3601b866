"""GPU parity of distCUDA2 (csrc/knn.cu) through the C ABI — SURVEY.md §8 (f3).

Three-way, all bit-exact: product == C oracle (small/medium sizes), product == the REFERENCE ITSELF
(oracle/_ref/libsimple_knn_ref.so, simple-knn compiled unmodified from /root/reference by oracle/Makefile
in the build container) up to the 1 M points of the benchmark configuration, and oracle == reference
(which is what pins the oracle)."""
import os

import numpy as np
import pytest
import torch

from oracle import knn as oknn
from tests import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn_golden.npz")


def _dev():
    return torch.device("cuda", 0)


def _gpu(pts):
    from simple_knn._C import distCUDA2          # the import the reference uses (gaussian_model.py:9)
    out = distCUDA2(torch.from_numpy(np.ascontiguousarray(pts)).to(_dev()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _sets():
    rng = np.random.default_rng(7)
    s = {}
    s["gauss_20k"] = rng.standard_normal((20000, 3)).astype(np.float32)
    s["humanoid_30k"] = util.humanoid_scene(P=30000, H=64, W=64, sh_degree=0).means3D.detach().cpu().numpy()
    s["plane_5k"] = (rng.random((5000, 3)) * np.array([1, 1, 0])).astype(np.float32)        # degenerate extent in z
    s["line_3k"] = (rng.random((3000, 1)) * np.array([[1, 2, 3]])).astype(np.float32)
    g = np.stack(np.meshgrid(np.arange(24), np.arange(24), np.arange(24)), -1).reshape(-1, 3).astype(np.float32)
    s["lattice_ties"] = g                                                                    # massive distance ties
    d = rng.standard_normal((4000, 3)).astype(np.float32)
    d[1000:3000] = d[:2000]                                                                  # every point doubled
    s["duplicates"] = d
    s["all_same"] = np.ones((1000, 3), np.float32) * 0.25
    s["far_offset"] = (rng.standard_normal((6000, 3)) * 1e-3 + 1000.0).astype(np.float32)   # cancellation-heavy
    s["ragged_1025"] = rng.standard_normal((1025, 3)).astype(np.float32)                     # one full level + 1
    s["ragged_33"] = rng.standard_normal((33, 3)).astype(np.float32)
    return s


@pytest.mark.parametrize("name", list(_sets()))
def test_matches_oracle_bit_exact(name):
    pts = np.ascontiguousarray(_sets()[name], dtype=np.float32)
    got = _gpu(pts)
    ref = oknn.oracle_dist2(pts)
    assert np.array_equal(got, ref), f"{name}: {(got != ref).sum()} of {len(ref)} differ"


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 31, 32])
def test_tiny_inputs(P):
    rng = np.random.default_rng(P)
    pts = rng.standard_normal((P, 3)).astype(np.float32)
    got = _gpu(pts)
    ref = oknn.oracle_dist2(pts)
    assert np.array_equal(got, ref)


def test_empty_and_noncontiguous_and_dtype():
    from simple_knn._C import distCUDA2
    assert distCUDA2(torch.zeros(0, 3, device=_dev())).shape == (0,)
    rng = np.random.default_rng(3)
    base = torch.from_numpy(rng.standard_normal((3, 5000)).astype(np.float32)).to(_dev())
    pts = base.t()                                                       # non-contiguous view
    got = distCUDA2(pts).cpu().numpy()
    assert np.array_equal(got, oknn.oracle_dist2(pts.cpu().numpy()))
    got64 = distCUDA2(pts.double()).cpu().numpy()                        # callers pass .float(); be lenient
    assert np.array_equal(got64, got)


def test_golden_vectors():
    g = np.load(GOLD)
    for name in ("humanoid", "uniform", "clustered"):
        assert np.array_equal(_gpu(g[name + "_pts"]), g[name + "_dist2"]), name


needs_ref = pytest.mark.skipif(not oknn.reference_available(),
                               reason="oracle/_ref/libsimple_knn_ref.so not built (needs /root/reference at build time)")


@needs_ref
def test_oracle_matches_reference_build():
    """Pins the oracle: C restatement == the reference's own compiled simple-knn, bit for bit."""
    for name, pts in _sets().items():
        if len(pts) < 4:
            continue
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        ref = oknn.reference_dist2(torch.from_numpy(pts).to(_dev())).cpu().numpy()
        assert np.array_equal(oknn.oracle_dist2(pts), ref), name


@needs_ref
def test_golden_matches_reference_build():
    g = np.load(GOLD)
    for name in ("humanoid", "uniform", "clustered"):
        ref = oknn.reference_dist2(torch.from_numpy(g[name + "_pts"]).to(_dev())).cpu().numpy()
        assert np.array_equal(ref, g[name + "_dist2"]), name


@needs_ref
@pytest.mark.parametrize("P", [100_000, 1_000_000])
def test_matches_reference_build_full_size(P):
    """BASELINE.json's cloud size: product vs the reference itself, bit for bit."""
    cloud = util.humanoid_scene(P=P, H=64, W=64, sh_degree=0).means3D.detach().to(torch.float32).contiguous()
    pts = cloud.to(_dev())
    from simple_knn._C import distCUDA2
    got = distCUDA2(pts)
    ref = oknn.reference_dist2(pts)
    assert torch.equal(got, ref), f"{(got != ref).sum().item()} of {P} differ"


def test_full_size_properties():
    """Size-independent properties at 1 M points (no reference needed): permutation equivariance,
    translation by a power of two, and scaling by 2 (all exact in fp32)."""
    from simple_knn._C import distCUDA2
    P = 1_000_000
    g = torch.Generator().manual_seed(5)
    pts = (torch.randn(P, 3, generator=g) * torch.tensor([0.3, 0.2, 0.9])).to(_dev())
    base = distCUDA2(pts)
    perm = torch.randperm(P, generator=g).to(_dev())
    assert torch.equal(distCUDA2(pts[perm]), base[perm])
    assert torch.equal(distCUDA2(pts * 2.0), base * 4.0)
    assert bool((base > 0).all()) and bool(torch.isfinite(base).all())
    # exact subset check against the oracle's arithmetic: brute force for 256 random queries
    q = torch.randint(0, P, (256,), generator=g).to(_dev())
    d = pts[q][:, None, :] - pts[None, :, :]            # other - self has the opposite sign; squares agree
    dx, dy, dz = (pts[None, :, 0] - pts[q][:, None, 0]), (pts[None, :, 1] - pts[q][:, None, 1]), \
        (pts[None, :, 2] - pts[q][:, None, 2])
    del d
    d2 = torch.addcmul(torch.addcmul(dx * dx, dy, dy), dz, dz)          # not fused: allow 1 ulp-level slack below
    d2[torch.arange(256, device=_dev()), q] = float("inf")
    best = torch.topk(d2, 3, dim=1, largest=False).values
    approx = (best[:, 0] + best[:, 1] + best[:, 2]) / 3.0
    assert torch.allclose(base[q], approx, rtol=1e-5, atol=0)
