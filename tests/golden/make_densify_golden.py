"""Writes tests/golden/ref_densify.npz — outputs of the REFERENCE's own GaussianModel code
(gaussiansplatting/scene/gaussian_model.py, imported from /root/reference and run on the CPU: its
device="cuda" requests are mapped to the CPU, `plyfile` is stubbed because only the tensor methods are
exercised, and torch.normal is replaced by a recorded draw so the GPU product can be fed the same samples).
RUN IN THE AUTHORING CONTAINER ONLY.   python tests/golden/make_densify_golden.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _cpu_only():
    for fn_name in ("zeros", "ones", "tensor", "empty", "normal"):
        orig = getattr(torch, fn_name)

        def wrapped(*a, __orig=orig, **k):
            if "device" in k:
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, fn_name, wrapped)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None


def build_state(P, sh_degree, seed):
    """Seeded raw tensors of a mid-training model: parameters, Adam moments, statistics."""
    g = torch.Generator().manual_seed(seed)
    K = (sh_degree + 1) ** 2
    r = lambda *s: torch.randn(*s, generator=g)
    st = {
        "xyz": r(P, 3) * 0.4, "f_dc": r(P, 1, 3), "f_rest": r(P, K - 1, 3) * 0.1, "opacity": r(P, 1) * 2.0,
        "scaling": r(P, 3) * 0.6 - 3.5, "rotation": torch.nn.functional.normalize(r(P, 4)) * (1 + 0.2 * r(P, 1)),
    }
    for k in list(st):
        st["m_" + k] = r(*st[k].shape) * 1e-3
        st["v_" + k] = (r(*st[k].shape) * 1e-3) ** 2
    st["xyz_gradient_accum"] = (torch.rand(P, 1, generator=g) * 6e-3) * (torch.rand(P, 1, generator=g) > 0.3)
    st["denom"] = torch.randint(0, 12, (P, 1), generator=g).float()        # zeros -> NaN grads, as in training
    st["max_radii2D"] = torch.rand(P, generator=g) * 40.0
    return st


def run_reference(st, sh_degree, op, split_samples_holder):
    for pkg in ("gaussiansplatting", "gaussiansplatting.utils", "gaussiansplatting.scene"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = object
    sys.modules["plyfile"] = ply
    for name in ("general_utils", "system_utils", "sh_utils", "graphics_utils"):
        if f"gaussiansplatting.utils.{name}" not in sys.modules or not hasattr(sys.modules[f"gaussiansplatting.utils.{name}"], "__file__"):
            _load(f"{REF}/gaussiansplatting/utils/{name}.py", f"gaussiansplatting.utils.{name}")
    gm = _load(f"{REF}/gaussiansplatting/scene/gaussian_model.py", "ref_gaussian_model")
    m = gm.GaussianModel(sh_degree)
    P = st["xyz"].shape[0]
    from torch import nn
    m._xyz = nn.Parameter(st["xyz"].clone().requires_grad_(True))
    m._features_dc = nn.Parameter(st["f_dc"].clone().requires_grad_(True))
    m._features_rest = nn.Parameter(st["f_rest"].clone().requires_grad_(True))
    m._opacity = nn.Parameter(st["opacity"].clone().requires_grad_(True))
    m._scaling = nn.Parameter(st["scaling"].clone().requires_grad_(True))
    m._rotation = nn.Parameter(st["rotation"].clone().requires_grad_(True))
    m.max_radii2D = st["max_radii2D"].clone()
    m.percent_dense = 0.01
    groups = [{"params": [m._xyz], "lr": 1e-4, "name": "xyz"}, {"params": [m._features_dc], "lr": 1e-3, "name": "f_dc"},
              {"params": [m._features_rest], "lr": 1e-4, "name": "f_rest"}, {"params": [m._opacity], "lr": 1e-2, "name": "opacity"},
              {"params": [m._scaling], "lr": 5e-3, "name": "scaling"}, {"params": [m._rotation], "lr": 1e-3, "name": "rotation"}]
    m.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for gdict in groups:
        p = gdict["params"][0]
        m.optimizer.state[p] = {"step": torch.tensor(7.0), "exp_avg": st["m_" + gdict["name"]].clone(),
                                "exp_avg_sq": st["v_" + gdict["name"]].clone()}
    m.xyz_gradient_accum = st["xyz_gradient_accum"].clone()
    m.denom = st["denom"].clone()

    orig_normal = torch.normal

    def recorded_normal(mean=None, std=None, **k):
        gen = torch.Generator().manual_seed(99)
        s = orig_normal(mean=mean, std=std, generator=gen)
        split_samples_holder.append(s.detach().clone())
        return s
    torch.normal = recorded_normal
    try:
        if op == "densify_and_prune":
            m.densify_and_prune(2e-4, 0.05, 1.2, 20, 0.1)
        elif op == "densify_and_prune_no_screen":
            m.densify_and_prune(2e-4, 0.05, 1.2, None, 0.1)
        elif op == "prune_only":
            m.prune_only(min_opacity=0.05, max_world_size=0.05)
        elif op == "reset_opacity":
            m.reset_opacity()
        else:
            raise ValueError(op)
    finally:
        torch.normal = orig_normal
    out = {}
    for gdict in m.optimizer.param_groups:
        p = gdict["params"][0]
        s = m.optimizer.state[p]
        out["p_" + gdict["name"]] = p.detach().numpy()
        out["m_" + gdict["name"]] = s["exp_avg"].numpy()
        out["v_" + gdict["name"]] = s["exp_avg_sq"].numpy()
    out["xyz_gradient_accum"] = m.xyz_gradient_accum.numpy()
    out["denom"] = m.denom.numpy()
    out["max_radii2D"] = m.max_radii2D.numpy()
    for k, attr in (("xyz", "_xyz"), ("opacity", "_opacity")):
        assert getattr(m, attr) is [g for g in m.optimizer.param_groups if g["name"] == k][0]["params"][0]
    return out


def main():
    _cpu_only()
    blob = {}
    cases = [("a", 1200, 1, 11, "densify_and_prune"), ("b", 800, 0, 12, "densify_and_prune_no_screen"),
             ("c", 600, 2, 13, "prune_only"), ("d", 300, 0, 14, "reset_opacity")]
    for tag, P, deg, seed, op in cases:
        st = build_state(P, deg, seed)
        holder = []
        out = run_reference(st, deg, op, holder)
        blob[f"{tag}_op"] = np.array(op)
        blob[f"{tag}_sh_degree"] = np.array(deg)
        for k, v in st.items():
            blob[f"{tag}_in_{k}"] = v.numpy()
        for k, v in out.items():
            blob[f"{tag}_out_{k}"] = v
        if holder:
            blob[f"{tag}_samples"] = holder[0].numpy()
        print(tag, op, "P", P, "->", out["p_xyz"].shape[0], "samples", holder[0].shape if holder else None)
    np.savez_compressed(os.path.join(HERE, "ref_densify.npz"), **blob)


if __name__ == "__main__":
    main()
