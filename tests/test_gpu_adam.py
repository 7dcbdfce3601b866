"""GPU: fused Adam + densification statistics against the reference's optimizer,
torch.optim.Adam(eps=1e-15) with the six parameter groups of gaussian_model.py:138-159, run on the CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu

LRS = {"xyz": 0.00005 * 5.0, "f_dc": 0.0125, "f_rest": 0.0125 / 20.0, "opacity": 0.01, "scaling": 0.005,
       "rotation": 0.001}
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}


def _groups(P, device, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"params": [torch.randn(P, *SHAPES[n], generator=g).to(device).requires_grad_(True)], "lr": LRS[n],
             "name": n} for n in LRS]


def test_fused_adam_matches_torch_adam(cuda_device):
    from gaussianip_b200.optim import FusedGaussianAdam
    P = 5003                                             # not a multiple of the vector width
    ref_groups, our_groups = _groups(P, "cpu"), _groups(P, cuda_device)
    ref = torch.optim.Adam(ref_groups, lr=0.0, eps=1e-15)
    ours = FusedGaussianAdam(our_groups, lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(1)
    for it in range(6):
        for rg, og in zip(ref_groups, our_groups):
            grad = torch.randn(rg["params"][0].shape, generator=g) * (10.0 ** (it - 3))
            rg["params"][0].grad = grad.clone()
            og["params"][0].grad = grad.to(cuda_device)
        if it == 3:                                       # the reference rewrites the xyz lr every step
            ref.param_groups[0]["lr"] = ours.param_groups[0]["lr"] = 1.2345e-4
        ref.step(); ours.step()
    for rg, og in zip(ref_groups, our_groups):
        rp, op = rg["params"][0], og["params"][0]
        torch.testing.assert_close(op.detach().cpu(), rp.detach(), rtol=2e-6, atol=1e-7)
        rs, os_ = ref.state[rp], ours.state[op]
        # moments: within fp32 rounding of the largest entry (cancellation m + 0.1 (g - m) near zero)
        for key in ("exp_avg", "exp_avg_sq"):
            scale = float(rs[key].abs().max())
            assert float((os_[key].cpu() - rs[key]).abs().max()) <= 2e-6 * scale, key
        assert os_["step"] == 6
    ours.zero_grad()
    assert all(g["params"][0].grad is None for g in our_groups)


def test_densification_statistics_in_the_same_launch(cuda_device):
    from gaussianip_b200.optim import FusedGaussianAdam
    P = 4097
    dev = cuda_device
    groups = _groups(P, dev)
    opt = FusedGaussianAdam(groups, eps=1e-15)
    g = torch.Generator().manual_seed(4)
    for gr in groups:
        gr["params"][0].grad = torch.randn(gr["params"][0].shape, generator=g).to(dev)
    vgrad = torch.randn(P, 3, generator=g).to(dev)
    radii = ((torch.rand(P, generator=g) * 40).to(torch.int32) * (torch.rand(P, generator=g) > 0.4)).to(torch.int32).to(dev)
    acc = torch.rand(P, 1, generator=g).to(dev); den = torch.ones(P, 1, device=dev); mr = (torch.rand(P, generator=g) * 20).to(dev)
    acc0, den0, mr0 = acc.clone(), den.clone(), mr.clone()
    before = groups[0]["params"][0].detach().clone()
    opt.step(densify=(acc, den, mr, vgrad, radii))
    f = radii > 0                                          # GaussianIP.py:456-457, gaussian_model.py:420-422
    exp_acc, exp_den, exp_mr = acc0.clone(), den0.clone(), mr0.clone()
    exp_acc[f] += torch.norm(vgrad[f, :2], dim=-1, keepdim=True)
    exp_den[f] += 1
    exp_mr[f] = torch.max(mr0[f], radii[f].float())
    torch.testing.assert_close(acc, exp_acc, rtol=1e-6, atol=1e-7)
    assert torch.equal(den, exp_den) and torch.equal(mr, exp_mr)
    assert not torch.equal(groups[0]["params"][0].detach(), before)      # Adam ran in the same launch


def test_amp_unscale_and_state_roundtrip(cuda_device):
    from gaussianip_b200.optim import FusedGaussianAdam
    P = 1000
    a, b = _groups(P, cuda_device), _groups(P, cuda_device)
    oa, ob = FusedGaussianAdam(a, eps=1e-15), FusedGaussianAdam(b, eps=1e-15)
    g = torch.Generator().manual_seed(2)
    for it in range(3):
        for ga, gb in zip(a, b):
            grad = torch.randn(ga["params"][0].shape, generator=g).to(cuda_device)
            ga["params"][0].grad = grad
            gb["params"][0].grad = grad * 65536.0
        oa.step(); ob.step(grad_scale=1.0 / 65536.0)      # GradScaler-scaled gradients, unscaled in the kernel
        if it == 1:
            sd = ob.state_dict(); ob.load_state_dict(sd)
    for ga, gb in zip(a, b):
        torch.testing.assert_close(ga["params"][0], gb["params"][0], rtol=1e-6, atol=1e-8)
    cpu_p = torch.zeros(3, requires_grad=True)
    cpu_p.grad = torch.zeros(3)
    with pytest.raises(ValueError, match="CUDA"):        # no CPU fallback
        FusedGaussianAdam([{"params": [cpu_p], "lr": 1e-3, "name": "cpu"}]).step()


def test_state_dict_interchanges_with_torch_adam(cuda_device):
    """GaussianModel.capture()/restore() round-trip optimizer.state_dict() (gaussian_model.py:62-82): a checkpoint
    written by the reference's torch.optim.Adam loads here and continues identically, and the reverse."""
    from gaussianip_b200.optim import FusedGaussianAdam
    P, dev = 777, cuda_device
    ref_groups, our_groups = _groups(P, dev, seed=3), _groups(P, dev, seed=3)
    ref = torch.optim.Adam(ref_groups, lr=0.0, eps=1e-15)
    ours = FusedGaussianAdam(our_groups, lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(5)

    def grads():
        for rg, og in zip(ref_groups, our_groups):
            gr = torch.randn(rg["params"][0].shape, generator=g).to(dev)
            rg["params"][0].grad, og["params"][0].grad = gr.clone(), gr.clone()

    for _ in range(3):
        grads(); ref.step(); ours.step()
    sd_t, sd_o = ref.state_dict(), ours.state_dict()
    assert set(sd_o) == {"state", "param_groups"} and set(sd_o["state"]) == set(sd_t["state"])
    for i in sd_t["state"]:
        assert set(sd_t["state"][i]) <= set(sd_o["state"][i])
        assert float(sd_o["state"][i]["step"]) == float(sd_t["state"][i]["step"]) == 3.0
    assert [pg["params"] for pg in sd_o["param_groups"]] == [pg["params"] for pg in sd_t["param_groups"]]
    assert [pg["name"] for pg in sd_o["param_groups"]] == list(LRS)
    # cross-load: torch checkpoint -> fused, fused checkpoint -> torch, then two more steps each
    ref2 = torch.optim.Adam(ref_groups, lr=0.0, eps=1e-15)
    ours2 = FusedGaussianAdam(our_groups, lr=0.0, eps=1e-15)
    ref2.load_state_dict(sd_o)
    ours2.load_state_dict(sd_t)
    for _ in range(2):
        grads(); ref2.step(); ours2.step()
    for rg, og in zip(ref_groups, our_groups):
        torch.testing.assert_close(og["params"][0], rg["params"][0], rtol=2e-6, atol=1e-7)
        assert int(ours2.state[og["params"][0]]["step"]) == 5


def test_group_without_gradient_keeps_its_own_step_count(cuda_device):
    """torch.optim.Adam keeps `step` per parameter: a group whose .grad was None for two steps is bias-corrected
    with ITS count when it joins."""
    from gaussianip_b200.optim import FusedGaussianAdam
    P, dev = 1030, cuda_device
    ref_groups, our_groups = _groups(P, dev, seed=8), _groups(P, dev, seed=8)
    ref = torch.optim.Adam(ref_groups, lr=0.0, eps=1e-15)
    ours = FusedGaussianAdam(our_groups, lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        for k, (rg, og) in enumerate(zip(ref_groups, our_groups)):
            if k == 2 and it < 2:                       # f_rest: no gradient on the first two steps
                rg["params"][0].grad = og["params"][0].grad = None
                continue
            gr = torch.randn(rg["params"][0].shape, generator=g).to(dev)
            rg["params"][0].grad, og["params"][0].grad = gr.clone(), gr.clone()
        ref.step(); ours.step()
    for rg, og in zip(ref_groups, our_groups):
        torch.testing.assert_close(og["params"][0], rg["params"][0], rtol=2e-6, atol=1e-7)
    assert int(ours.state[our_groups[2]["params"][0]]["step"]) == 3
    assert int(ours.state[our_groups[0]["params"][0]]["step"]) == 5
