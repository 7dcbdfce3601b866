"""Fused Adam + densification statistics (additive).

Drop-in for the optimizer GaussianModel.training_setup builds
(gaussiansplatting/scene/gaussian_model.py:138-159: ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` over six
parameter groups named xyz / f_dc / f_rest / opacity / scaling / rotation, each with its own lr) with the
surface the reference touches:

* ``param_groups`` — list of dicts with ``params`` (one tensor), ``lr``, ``name``
  (``update_learning_rate`` rewrites ``lr`` per step, gaussian_model.py:161-167);
* ``state[param]`` — dict with ``exp_avg`` / ``exp_avg_sq`` (read and replaced by the densify / prune
  helpers, gaussian_model.py:266-330) and ``step``;
* ``step()``, ``zero_grad(set_to_none=True)``, ``state_dict()`` / ``load_state_dict()``.

``step()`` is ONE kernel launch for all groups; passing ``densify=(xyz_gradient_accum, denom,
max_radii2D, viewspace_grad, radii)`` also updates the densification statistics
(threestudio/systems/GaussianIP.py:456-457, gaussian_model.py:420-422) in the same launch.
CUDA only (no CPU fallback); element-wise math follows torch.optim.Adam's rounding order.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Iterable, Optional, Sequence

import torch

from . import _lib


class FusedGaussianAdam:
    def __init__(self, params: Iterable[dict], lr: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-15):
        self.param_groups = []
        for g in params:
            g = dict(g)
            ps = list(g["params"]) if not torch.is_tensor(g["params"]) else [g["params"]]
            if len(ps) != 1:
                raise ValueError("each group must hold exactly one tensor (as GaussianModel.training_setup does)")
            g["params"] = ps
            g.setdefault("lr", lr)
            g.setdefault("betas", betas)
            g.setdefault("eps", eps)
            self.param_groups.append(g)
        if len(self.param_groups) > 8:
            raise ValueError("at most 8 parameter groups")
        self.defaults = {"lr": lr, "betas": betas, "eps": eps}
        self.state = defaultdict(dict)

    # ---- torch.optim.Optimizer surface -----------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            p = g["params"][0]
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    # Keys torch.optim.Adam keeps in every param group: emitted so that a state_dict saved here loads into the
    # reference's torch.optim.Adam (GaussianModel.capture()/restore(), gaussian_model.py:62-82) and vice versa.
    _TORCH_GROUP_DEFAULTS = {"weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None,
                             "capturable": False, "differentiable": False, "fused": None,
                             "decoupled_weight_decay": False}

    def state_dict(self) -> dict:
        """torch.optim.Optimizer.state_dict() layout: ``state`` keyed by parameter index with step / exp_avg /
        exp_avg_sq, ``param_groups`` with ``params`` as index lists."""
        state, groups = {}, []
        for i, g in enumerate(self.param_groups):
            meta = dict(self._TORCH_GROUP_DEFAULTS)
            meta.update({k: v for k, v in g.items() if k != "params"})
            meta["params"] = [i]
            groups.append(meta)
            st = self.state.get(g["params"][0])
            if st and "exp_avg" in st:
                state[i] = {"step": torch.tensor(float(st.get("step", 0))), "exp_avg": st["exp_avg"].clone(),
                            "exp_avg_sq": st["exp_avg_sq"].clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts what torch.optim.Adam.state_dict() (or this class) produced for the same groups."""
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g, meta in zip(self.param_groups, groups):
            idx = list(meta["params"])
            if len(idx) != 1:
                raise ValueError("each group must hold exactly one tensor")
            g.update({k: v for k, v in meta.items() if k != "params"})
            p = g["params"][0]
            st = sd["state"].get(idx[0], sd["state"].get(str(idx[0])))
            if st is None:
                self.state.pop(p, None)
                continue
            step = st.get("step", 0)
            step = int(step.item()) if torch.is_tensor(step) else int(step)
            self.state[p] = {"step": step,
                             "exp_avg": st["exp_avg"].detach().to(device=p.device, dtype=p.dtype).clone(),
                             "exp_avg_sq": st["exp_avg_sq"].detach().to(device=p.device, dtype=p.dtype).clone()}

    def _state_of(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, densify: Optional[Sequence[torch.Tensor]] = None, grad_scale: float = 1.0) -> None:
        lib = _lib.load()
        active = [g for g in self.param_groups if g["params"][0].grad is not None and g["params"][0].numel() > 0]
        if not active and densify is None:
            return
        dev = (active[0]["params"][0] if active else densify[0]).device
        if dev.type != "cuda":
            raise ValueError("FusedGaussianAdam runs on CUDA tensors only (no CPU fallback)")
        n = len(active)
        ptr = lambda ts: (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in ts])
        ps, gs, ms, vs, keep = [], [], [], [], []
        for g in active:
            p = g["params"][0]
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"group {g.get('name')}: parameters must be contiguous fp32")
            grad = p.grad
            if grad.dtype != torch.float32 or not grad.is_contiguous():
                grad = grad.float().contiguous()
                keep.append(grad)
            st = self._state_of(p)
            if st["exp_avg"].shape != p.shape:        # densify/prune replaced the parameter: caller must reset state
                raise ValueError(f"group {g.get('name')}: optimizer state shape {tuple(st['exp_avg'].shape)} "
                                 f"does not match parameter {tuple(p.shape)}")
            st["step"] = st.get("step", 0) + 1
            ps.append(p); gs.append(grad); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
        b1, b2 = (active[0].get("betas", self.defaults["betas"]) if active else self.defaults["betas"])
        eps = active[0].get("eps", self.defaults["eps"]) if active else self.defaults["eps"]
        # bias correction per group from ITS step count (a group whose grad was None on earlier steps lags, as in torch)
        steps = (C.c_longlong * max(n, 1))(*[int(self.state[g["params"][0]]["step"]) for g in active])
        counts = (C.c_longlong * max(n, 1))(*[p.numel() for p in ps])
        lrs = (C.c_float * max(n, 1))(*[float(g["lr"]) for g in active])
        stats_n, sp = 0, [None] * 5
        if densify is not None:
            acc, den, mr, vgrad, radii = densify
            if radii.dtype != torch.int32:
                radii = radii.to(torch.int32)
            vgrad = vgrad.float().contiguous()
            keep += [vgrad, radii]
            stats_n = radii.numel()
            for t in (acc, den, mr):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != stats_n:
                    raise ValueError("densification statistics must be contiguous fp32 with one entry per Gaussian")
            sp = [vgrad.data_ptr(), radii.data_ptr(), acc.data_ptr(), den.data_ptr(), mr.data_ptr()]
        with torch.cuda.device(dev):
            rc = lib.gsb_adam_step_groups(n, ptr(ps), ptr(gs), ptr(ms), ptr(vs), counts, lrs, float(b1), float(b2),
                                          float(eps), steps, float(grad_scale), int(stats_n), sp[0], sp[1], sp[2], sp[3], sp[4],
                                   torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "gsb_adam_step_groups")
        del keep
