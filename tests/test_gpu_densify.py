"""GPU parity of the fused density control (gaussianip_b200/densify.py, csrc/compact.cu) — SURVEY.md §8 (f4).

* against tests/golden/ref_densify.npz = outputs of the reference's own GaussianModel methods (CPU run): moved rows
  bit-exact, recomputed values (children of a split, reset opacities) to 1e-6 (CPU vs GPU transcendental rounding);
* against the pinned oracle (oracle/densify_torch.py) executed on the same GPU — the reference's own tensor-op
  sequence — bit-exact for everything, up to 1 M points."""
import os
import types

import numpy as np
import pytest
import torch
from torch import nn

from oracle import densify_torch as OD
from tests.test_oracle_densify import load_case

pytestmark = pytest.mark.gpu
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation"}


def _dev():
    return torch.device("cuda", 0)


def make_model(state, fused=False):
    """A GaussianModel-shaped object (reference attribute names) over an oracle-layout state on the GPU."""
    from gaussianip_b200.optim import FusedGaussianAdam
    m = types.SimpleNamespace()
    groups = []
    for n in OD.NAMES:
        p = nn.Parameter(state["p_" + n].clone().to(_dev()).requires_grad_(True))
        setattr(m, ATTR[n], p)
        groups.append({"params": [p], "lr": 1e-3, "name": n})
    m.optimizer = FusedGaussianAdam(groups, lr=0.0, eps=1e-15) if fused else torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for g in m.optimizer.param_groups:
        p = g["params"][0]
        m.optimizer.state[p] = {"step": 7 if fused else torch.tensor(7.0),
                                "exp_avg": state["m_" + g["name"]].clone().to(_dev()),
                                "exp_avg_sq": state["v_" + g["name"]].clone().to(_dev())}
    for k in OD.STATS:
        setattr(m, k, state[k].clone().to(_dev()))
    m.percent_dense = 0.01
    return m


def model_state(m):
    out = {}
    for g in m.optimizer.param_groups:
        p = g["params"][0]
        assert getattr(m, ATTR[g["name"]]) is p and p.requires_grad and isinstance(p, nn.Parameter)
        st = m.optimizer.state[p]
        out["p_" + g["name"]], out["m_" + g["name"]], out["v_" + g["name"]] = p.detach(), st["exp_avg"], st["exp_avg_sq"]
    assert len(m.optimizer.state) == len(m.optimizer.param_groups)        # no stale keys left behind
    for k in OD.STATS:
        out[k] = getattr(m, k)
    return out


def run_product(op, m, samples):
    from gaussianip_b200 import densify as D
    s = None if samples is None else samples.to(_dev())
    if op == "densify_and_prune":
        D.densify_and_prune(m, 2e-4, 0.05, 1.2, 20, 0.1, split_samples=s)
    elif op == "densify_and_prune_no_screen":
        D.densify_and_prune(m, 2e-4, 0.05, 1.2, None, 0.1, split_samples=s)
    elif op == "prune_only":
        D.prune_only(m, min_opacity=0.05, max_world_size=0.05)
    elif op == "reset_opacity":
        D.reset_opacity(m)
    return model_state(m)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_matches_reference_golden(tag, fused):
    op, inp, ref, samples = load_case(tag)
    got = run_product(op, make_model(OD.state_from_inputs(inp), fused), samples)
    for k, v in ref.items():
        g = got[k].cpu().numpy()
        assert g.shape == v.shape, f"{tag}/{k}: {g.shape} vs {v.shape}"
        exact = np.array_equal(g, v)
        if not exact:
            # only values the operation RECOMPUTES may differ (GPU vs CPU exp/log/sigmoid), never moved rows
            assert k in ("p_xyz", "p_scaling", "p_opacity"), f"{tag}/{k} moved rows differ"
            np.testing.assert_allclose(g, v, rtol=2e-6, atol=2e-6, err_msg=f"{tag}/{k}")


def big_state(P, deg, seed):
    g = torch.Generator().manual_seed(seed)
    K = (deg + 1) ** 2
    r = lambda *s: torch.randn(*s, generator=g)
    inp = {"xyz": r(P, 3) * 0.4, "f_dc": r(P, 1, 3), "f_rest": r(P, K - 1, 3) * 0.1, "opacity": r(P, 1) * 2.0,
           "scaling": r(P, 3) * 0.6 - 3.5, "rotation": r(P, 4)}
    for k in list(inp):
        inp["m_" + k] = r(*inp[k].shape) * 1e-3
        inp["v_" + k] = (r(*inp[k].shape) * 1e-3) ** 2
    inp["xyz_gradient_accum"] = (torch.rand(P, 1, generator=g) * 6e-3) * (torch.rand(P, 1, generator=g) > 0.3)
    inp["denom"] = torch.randint(0, 12, (P, 1), generator=g).float()
    inp["max_radii2D"] = torch.rand(P, generator=g) * 40.0
    return OD.state_from_inputs(inp)


@pytest.mark.parametrize("P,deg,op", [(50_000, 3, "densify_and_prune"), (1_000_000, 0, "densify_and_prune"),
                                      (1_000_000, 0, "prune_only"), (2049, 1, "densify_and_prune_no_screen"),
                                      (20_000, 2, "reset_opacity")])
def test_matches_oracle_on_gpu_bit_exact(P, deg, op):
    st = big_state(P, deg, 100 + deg)
    # same RNG stream for both: the product calls torch.normal with the reference's shapes
    torch.manual_seed(1234)
    ref = OD.run_case(op, {k: v.to(_dev()) for k, v in st.items()})
    torch.manual_seed(1234)
    got = run_product(op, make_model(st, fused=True), None)
    for k, v in ref.items():
        assert got[k].shape == v.shape, f"{k}: {tuple(got[k].shape)} vs {tuple(v.shape)}"
        assert torch.equal(got[k], v), k


def test_primitives_edge_cases():
    from gaussianip_b200 import densify as D
    dev = _dev()
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 7, 2048, 2049, 100_003):
        for p in (0.0, 0.3, 1.0):
            mask = (torch.rand(n, generator=g) < p).to(dev) if p < 1.0 else torch.ones(n, dtype=torch.bool, device=dev)
            idx = D.mask_to_index(mask)
            assert torch.equal(idx, torch.nonzero(mask).squeeze(1)), (n, p)
    # 30 tensors of different widths in one call (two launches of <= 24), with an index and an offset
    n = 5000
    srcs = [torch.randn(n, w, generator=g).to(dev) if w > 1 else torch.randn(n, generator=g).to(dev) for w in
            [1, 3, 4, 45, 1, 3] * 5]
    idx = torch.randint(0, n, (1234,), generator=g).to(dev)
    dsts = [torch.full((2000,) + tuple(s.shape[1:]), -1.0, device=dev) for s in srcs]
    srcs[7] = None
    D.gather_rows(srcs, dsts, 1234, idx, dst_row0=100)
    for s, d in zip(srcs, dsts):
        assert bool((d[:100] == -1).all()) and bool((d[1334:] == -1).all())
        want = torch.zeros_like(d[100:1334]) if s is None else s[idx]
        assert torch.equal(d[100:1334], want)
    with pytest.raises(RuntimeError):
        D.mask_to_index(torch.zeros(4, dtype=torch.bool))


def test_training_continues_after_density_control():
    """FusedGaussianAdam steps on the re-shaped parameters and the carried-over moments."""
    from gaussianip_b200 import densify as D
    m = make_model(big_state(30_000, 1, 5), fused=True)
    D.densify_and_prune(m, 2e-4, 0.05, 1.2, 20, 0.1)
    P = m._xyz.shape[0]
    before = m._xyz.detach().clone()
    for g in m.optimizer.param_groups:
        p = g["params"][0]
        assert p.shape[0] == P
        p.grad = torch.ones_like(p)
    m.optimizer.step()
    torch.cuda.synchronize()
    assert not torch.equal(before, m._xyz.detach()) and bool(torch.isfinite(m._xyz).all())
