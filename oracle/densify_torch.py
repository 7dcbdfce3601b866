"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference's adaptive density control
(gaussiansplatting/scene/gaussian_model.py:216-219, 266-428) as pure functions over a state dict.

State: {"p_<name>", "m_<name>", "v_<name>" for name in xyz, f_dc, f_rest, opacity, scaling, rotation}
(parameter, Adam exp_avg, exp_avg_sq) + "xyz_gradient_accum" [P,1], "denom" [P,1], "max_radii2D" [P].
Plain boolean indexing and torch.cat, exactly the tensor operations the reference performs, on whatever device
the tensors live on.  PINNED: tests/test_oracle_densify.py checks it against tests/golden/ref_densify.npz,
which holds outputs of the reference's own GaussianModel methods (tests/golden/make_densify_golden.py)."""
from __future__ import annotations

import torch

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
STATS = ("xyz_gradient_accum", "denom", "max_radii2D")


def build_rotation(r):
    """gaussiansplatting/utils/general_utils.py:78-100."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def prune_points(s, mask):
    """:299-312 with _prune_optimizer :281-297."""
    valid = ~mask
    out = {}
    for n in NAMES:
        for pre in ("p_", "m_", "v_"):
            out[pre + n] = s[pre + n][valid]
    for k in STATS:
        out[k] = s[k][valid]
    return out


def postfix(s, new):
    """:345-365 with cat_tensors_to_optimizer :314-343."""
    out = {}
    for n in NAMES:
        ext = new[n]
        out["p_" + n] = torch.cat((s["p_" + n], ext), dim=0)
        out["m_" + n] = torch.cat((s["m_" + n], torch.zeros_like(ext)), dim=0)
        out["v_" + n] = torch.cat((s["v_" + n], torch.zeros_like(ext)), dim=0)
    P = out["p_xyz"].shape[0]
    dev = out["p_xyz"].device
    out["xyz_gradient_accum"] = torch.zeros((P, 1), device=dev)
    out["denom"] = torch.zeros((P, 1), device=dev)
    out["max_radii2D"] = torch.zeros((P,), device=dev)
    return out


def densify_and_clone(s, grads, thr, extent, percent_dense):
    """:391-403."""
    sel = torch.norm(grads, dim=-1) >= thr
    sel = torch.logical_and(sel, torch.max(torch.exp(s["p_scaling"]), dim=1).values <= percent_dense * extent)
    return postfix(s, {n: s["p_" + n][sel] for n in NAMES})


def densify_and_split(s, grads, thr, extent, percent_dense, samples=None, N=2):
    """:367-389.  `samples` replaces the torch.normal draw."""
    n_init = s["p_xyz"].shape[0]
    dev = s["p_xyz"].device
    padded = torch.zeros((n_init,), device=dev)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = padded >= thr
    scaling = torch.exp(s["p_scaling"])
    sel = torch.logical_and(sel, torch.max(scaling, dim=1).values > percent_dense * extent)
    stds = scaling[sel].repeat(N, 1)
    if samples is None:
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds)
    rots = build_rotation(s["p_rotation"][sel]).repeat(N, 1, 1)
    new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + s["p_xyz"][sel].repeat(N, 1),
           "scaling": torch.log(scaling[sel].repeat(N, 1) / (0.8 * N)),
           "rotation": s["p_rotation"][sel].repeat(N, 1),
           "f_dc": s["p_f_dc"][sel].repeat(N, 1, 1), "f_rest": s["p_f_rest"][sel].repeat(N, 1, 1),
           "opacity": s["p_opacity"][sel].repeat(N, 1)}
    s2 = postfix(s, new)
    prune = torch.cat((sel, torch.zeros(N * int(sel.sum()), device=dev, dtype=torch.bool)))
    return prune_points(s2, prune)


def densify_and_prune(s, max_grad, min_opacity, extent, max_screen_size, max_world_size, percent_dense, samples=None):
    """:405-418."""
    grads = s["xyz_gradient_accum"] / s["denom"]
    grads[grads.isnan()] = 0.0
    s = densify_and_clone(s, grads, max_grad, extent, percent_dense)
    s = densify_and_split(s, grads, max_grad, extent, percent_dense, samples)
    prune_mask = (torch.sigmoid(s["p_opacity"]) < min_opacity).squeeze()
    if max_screen_size:
        big_vs = s["max_radii2D"] > max_screen_size
        big_ws = torch.exp(s["p_scaling"]).max(dim=1).values > max_world_size
        prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
    return prune_points(s, prune_mask)


def prune_only(s, min_opacity=0.05, max_world_size=0.01):
    """:423-428."""
    prune_mask = (torch.sigmoid(s["p_opacity"]) < min_opacity).squeeze()
    big_ws = torch.exp(s["p_scaling"]).max(dim=1).values > max_world_size
    return prune_points(s, torch.logical_or(prune_mask, big_ws))


def reset_opacity(s):
    """:216-219 with replace_tensor_to_optimizer :266-279."""
    op = torch.sigmoid(s["p_opacity"])
    x = torch.min(op, torch.ones_like(op) * 0.01)
    out = dict(s)
    out["p_opacity"] = torch.log(x / (1 - x))
    out["m_opacity"] = torch.zeros_like(out["p_opacity"])
    out["v_opacity"] = torch.zeros_like(out["p_opacity"])
    return out


def state_from_inputs(inp, device="cpu"):
    """inp: mapping with xyz.., m_xyz.., v_xyz.., statistics (tests/golden layout) -> oracle state."""
    s = {}
    for n in NAMES:
        s["p_" + n] = torch.as_tensor(inp[n]).clone().to(device)
        s["m_" + n] = torch.as_tensor(inp["m_" + n]).clone().to(device)
        s["v_" + n] = torch.as_tensor(inp["v_" + n]).clone().to(device)
    for k in STATS:
        s[k] = torch.as_tensor(inp[k]).clone().to(device)
    return s


def run_case(op, s, samples=None):
    """The four operations the golden file records, with the arguments it used."""
    if op == "densify_and_prune":
        return densify_and_prune(s, 2e-4, 0.05, 1.2, 20, 0.1, 0.01, samples)
    if op == "densify_and_prune_no_screen":
        return densify_and_prune(s, 2e-4, 0.05, 1.2, None, 0.1, 0.01, samples)
    if op == "prune_only":
        return prune_only(s, min_opacity=0.05, max_world_size=0.05)
    if op == "reset_opacity":
        return reset_opacity(s)
    raise ValueError(op)
