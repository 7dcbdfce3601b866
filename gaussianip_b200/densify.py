"""Densify / clone / split / prune with fused row movement (additive).

Function-for-method mirror of the adaptive-density control of ``GaussianModel``
(gaussiansplatting/scene/gaussian_model.py:243-418; same code in scene/gaussian_model.py and
avatar/gaussian_model.py).  Every function takes the model as first argument and reads / writes the
reference's own attribute names (``_xyz, _features_dc, _features_rest, _opacity, _scaling,
_rotation, optimizer, xyz_gradient_accum, denom, max_radii2D, percent_dense``), so a maintainer
switches with e.g. ``GaussianModel.densify_and_prune = gaussianip_b200.densify.densify_and_prune``.
The optimizer may be ``torch.optim.Adam`` or ``FusedGaussianAdam``: only ``param_groups`` (one tensor
per group, ``name`` key) and ``state[param]["exp_avg" | "exp_avg_sq"]`` are touched, as in the reference.

What changes is the data movement.  The reference indexes every parameter, both Adam moments of
each and the three statistics tensors separately with a boolean mask (21 ``t[mask]`` calls per
prune, each a nonzero + gather + host sync) and builds the split result by ``torch.cat`` followed by
a prune (two full copies).  Here a mask becomes an index list once (``gsb_mask_to_index``), the one
row count the host needs is read once, and ONE ``gsb_gather_rows`` launch moves the rows of all
tensors (zero-filling the Adam moments of new points); a split writes the survivors and the children
straight into their final places.  Values, row order and random draws are those of the reference
(``torch.normal`` is called with the same shapes, so the CUDA generator is consumed identically).
CUDA only — no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import _lib

_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
         "scaling": "_scaling", "rotation": "_rotation"}


# ---- primitives over the C ABI ---------------------------------------------------------------

def mask_to_index(mask: torch.Tensor) -> torch.Tensor:
    """Ascending positions of the True entries (== torch.nonzero(mask).squeeze(1)); one host read of the count."""
    if mask.device.type != "cuda":
        raise RuntimeError("gaussianip_b200.densify runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    m = mask.reshape(-1)
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    elif m.dtype != torch.uint8:
        m = (m != 0).view(torch.uint8)
    m = m.contiguous()
    n = m.numel()
    dev = m.device
    index = torch.empty(n, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        tmp = torch.empty(max(int(lib.gsb_mask_index_tmp_bytes(n)), 4), dtype=torch.uint8, device=dev)
        _lib.check(lib.gsb_mask_to_index(n, m.data_ptr(), index.data_ptr(), count.data_ptr(), tmp.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "gsb_mask_to_index")
    return index[:int(count.item())]


def gather_rows(srcs: Sequence[Optional[torch.Tensor]], dsts: Sequence[torch.Tensor], n_rows: int,
                index: Optional[torch.Tensor] = None, dst_row0: int = 0) -> None:
    """dsts[t][dst_row0 + r] = srcs[t][index[r] if index is not None else r] (zeros where srcs[t] is None),
    for all tensors in one launch (chunks of GSB_GATHER_MAX_TENSORS)."""
    if n_rows <= 0 or not dsts:
        return
    if any(d.shape[0] < dst_row0 + n_rows for d in dsts):
        raise ValueError("gather_rows: destination has too few rows")
    lib = _lib.load()
    dev = dsts[0].device
    keep = []
    rows = []
    for s, d in zip(srcs, dsts):
        if d.dtype != torch.float32 or not d.is_contiguous() or d.device != dev:
            raise ValueError("gather_rows moves contiguous fp32 CUDA tensors")
        w = (d.numel() // d.shape[0] if d.shape[0] else 0) if d.dim() > 1 else 1
        if w == 0:
            continue                     # e.g. f_rest at SH degree 0 is [P, 0, 3]: nothing to move
        if s is not None:
            if s.dtype != torch.float32 or s.device != dev:
                raise ValueError("gather_rows: source/destination dtype or device mismatch")
            if not s.is_contiguous():
                s = s.contiguous()
                keep.append(s)
            if ((s.numel() // s.shape[0] if s.shape[0] else w) if s.dim() > 1 else 1) != w:
                raise ValueError("gather_rows: row widths differ")
        rows.append((s, d, w))
    if index is not None:
        if index.dtype != torch.int64 or not index.is_contiguous() or index.numel() < n_rows:
            raise ValueError("index must be a contiguous int64 tensor with at least n_rows entries")
    M = _lib.GATHER_MAX_TENSORS
    if not rows:
        return
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for c0 in range(0, len(rows), M):
            chunk = rows[c0:c0 + M]
            n = len(chunk)
            sp = (C.c_void_p * n)(*[None if s is None else s.data_ptr() for s, _, _ in chunk])
            dp = (C.c_void_p * n)(*[d.data_ptr() for _, d, _ in chunk])
            wp = (C.c_int * n)(*[w for _, _, w in chunk])
            _lib.check(lib.gsb_gather_rows(n, sp, dp, wp, int(n_rows), None if index is None else index.data_ptr(),
                                           int(dst_row0), st), "gsb_gather_rows")
    del keep


# ---- optimizer-aware row movement -------------------------------------------------------------

def _groups(model):
    out = []
    for group in model.optimizer.param_groups:
        assert len(group["params"]) == 1
        p = group["params"][0]
        out.append((group, p, model.optimizer.state.get(p, None)))
    return out


def _has_moments(st) -> bool:
    return st is not None and "exp_avg" in st


def _install(model, group, old_param, new_tensor, state, new_avg, new_sq) -> nn.Parameter:
    """gaussian_model.py:283-296 / 325-343: swap the parameter object, carry the state over."""
    new_param = nn.Parameter(new_tensor.requires_grad_(True))
    if state is not None:
        if _has_moments(state):
            state["exp_avg"], state["exp_avg_sq"] = new_avg, new_sq
        del model.optimizer.state[old_param]
        model.optimizer.state[new_param] = state
    group["params"][0] = new_param
    setattr(model, _ATTR[group["name"]], new_param)
    return new_param


def _rebuild(model, keep_index: Optional[torch.Tensor], n_keep: int, new_rows: Optional[Dict[str, torch.Tensor]],
             stats: str) -> None:
    """New parameter/moment tensors = [old rows keep_index (all rows if None)] ++ [new_rows, moments zero].
    stats: 'gather' (prune_points :309-312) or 'zeros' (densification_postfix :363-365)."""
    groups = _groups(model)
    dev = groups[0][1].device
    n_new = 0 if not new_rows else next(iter(new_rows.values())).shape[0]
    total = n_keep + n_new
    srcs, dsts, plan = [], [], []
    for group, p, st in groups:
        shape = (total,) + tuple(p.shape[1:])
        new_p = torch.empty(shape, dtype=torch.float32, device=dev)
        new_avg = new_sq = None
        srcs.append(p.detach()); dsts.append(new_p)
        if _has_moments(st):
            new_avg, new_sq = torch.empty_like(new_p), torch.empty_like(new_p)
            srcs += [st["exp_avg"], st["exp_avg_sq"]]; dsts += [new_avg, new_sq]
        plan.append((group, p, st, new_p, new_avg, new_sq))
    stat_names = ("xyz_gradient_accum", "denom", "max_radii2D")
    new_stats = {}
    if stats == "gather":
        for name in stat_names:
            t = getattr(model, name)
            new_stats[name] = torch.empty((total,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
            srcs.append(t); dsts.append(new_stats[name])
    else:
        new_stats = {"xyz_gradient_accum": torch.zeros((total, 1), device=dev),
                     "denom": torch.zeros((total, 1), device=dev), "max_radii2D": torch.zeros((total,), device=dev)}
    gather_rows(srcs, dsts, n_keep, keep_index, 0)
    if n_new:
        srcs2, dsts2 = [], []
        for group, p, st, new_p, new_avg, new_sq in plan:
            ext = new_rows[group["name"]]
            srcs2.append(ext.detach().to(torch.float32).reshape((n_new,) + tuple(p.shape[1:]))); dsts2.append(new_p)
            if new_avg is not None:
                srcs2 += [None, None]; dsts2 += [new_avg, new_sq]
        if stats == "gather":
            for name in stat_names:     # not a reference path (postfix always resets), kept well defined
                srcs2.append(None); dsts2.append(new_stats[name])
        gather_rows(srcs2, dsts2, n_new, None, n_keep)
    for group, p, st, new_p, new_avg, new_sq in plan:
        _install(model, group, p, new_p, st, new_avg, new_sq)
    for name, t in new_stats.items():
        setattr(model, name, t)


# ---- the reference's methods --------------------------------------------------------------------

def get_scaling(model) -> torch.Tensor:
    return torch.exp(model._scaling)


def get_opacity(model) -> torch.Tensor:
    return torch.sigmoid(model._opacity)


def prune_points(model, mask: torch.Tensor) -> None:
    """gaussian_model.py:299-312 (+ _prune_optimizer :281-297): drop the rows where mask is True."""
    with torch.no_grad():
        keep = mask_to_index(~mask.reshape(-1).bool())
        _rebuild(model, keep, keep.numel(), None, "gather")


def densification_postfix(model, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                          new_rotation) -> None:
    """gaussian_model.py:345-365 (+ cat_tensors_to_optimizer :314-343)."""
    with torch.no_grad():
        new = {"xyz": new_xyz, "f_dc": new_features_dc, "f_rest": new_features_rest, "opacity": new_opacities,
               "scaling": new_scaling, "rotation": new_rotation}
        _rebuild(model, None, model._xyz.shape[0], new, "zeros")


def _build_rotation(r: torch.Tensor) -> torch.Tensor:
    """gaussiansplatting/utils/general_utils.py:78-100 (same op order)."""
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


def _take(t: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    out = torch.empty((index.numel(),) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
    gather_rows([t.detach()], [out], index.numel(), index, 0)
    return out


def densify_and_clone(model, grads, grad_threshold, scene_extent) -> None:
    """gaussian_model.py:391-403."""
    with torch.no_grad():
        sel = torch.norm(grads, dim=-1) >= grad_threshold
        sel = torch.logical_and(sel, torch.max(get_scaling(model), dim=1).values <= model.percent_dense * scene_extent)
        idx = mask_to_index(sel)
        params = {g["name"]: p for g, p, _ in _groups(model)}
        new = {name: _take(p, idx) for name, p in params.items()}
        _rebuild(model, None, model._xyz.shape[0], new, "zeros")


def densify_and_split(model, grads, grad_threshold, scene_extent, N: int = 2, samples: Optional[torch.Tensor] = None) -> None:
    """gaussian_model.py:367-389: the selected points are replaced by N children each.  The reference appends
    (postfix) and then prunes; the survivors and the children are written straight to their final rows here.
    ``samples`` (shape [N * n_selected, 3]) overrides the torch.normal draw (tests feed recorded draws)."""
    with torch.no_grad():
        n_init = model._xyz.shape[0]
        dev = model._xyz.device
        padded = torch.zeros((n_init,), device=dev)
        padded[:grads.shape[0]] = grads.squeeze()
        sel = padded >= grad_threshold
        scaling = get_scaling(model)
        sel = torch.logical_and(sel, torch.max(scaling, dim=1).values > model.percent_dense * scene_extent)
        idx = mask_to_index(sel)
        n = idx.numel()
        sc = _take(scaling, idx)                                   # get_scaling[selected]
        stds = sc.repeat(N, 1)
        if samples is None:
            samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds)
        rot_sel = _take(model._rotation, idx)
        rots = _build_rotation(rot_sel).repeat(N, 1, 1)
        new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + _take(model._xyz, idx).repeat(N, 1)
        new = {"xyz": new_xyz,
               "f_dc": _take(model._features_dc, idx).repeat(N, 1, 1),
               "f_rest": _take(model._features_rest, idx).repeat(N, 1, 1),
               "opacity": _take(model._opacity, idx).repeat(N, 1),
               "scaling": torch.log(stds / (0.8 * N)),
               "rotation": rot_sel.repeat(N, 1)}
        keep = mask_to_index(~sel)
        assert keep.numel() == n_init - n
        _rebuild(model, keep, keep.numel(), new, "zeros")


def densify_and_prune(model, max_grad, min_opacity, extent, max_screen_size, max_world_size,
                      split_samples: Optional[torch.Tensor] = None) -> None:
    """gaussian_model.py:405-418."""
    with torch.no_grad():
        grads = model.xyz_gradient_accum / model.denom
        grads[grads.isnan()] = 0.0
        densify_and_clone(model, grads, max_grad, extent)
        densify_and_split(model, grads, max_grad, extent, samples=split_samples)
        prune_mask = (get_opacity(model) < min_opacity).squeeze()
        if max_screen_size:
            big_points_vs = model.max_radii2D > max_screen_size
            big_points_ws = get_scaling(model).max(dim=1).values > max_world_size
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
        prune_points(model, prune_mask)


def prune_only(model, min_opacity=0.05, max_world_size=0.01) -> None:
    """gaussian_model.py:423-428."""
    with torch.no_grad():
        prune_mask = (get_opacity(model) < min_opacity).squeeze()
        big_points_ws = get_scaling(model).max(dim=1).values > max_world_size
        prune_points(model, torch.logical_or(prune_mask, big_points_ws))


def reset_opacity(model) -> None:
    """gaussian_model.py:216-219 + replace_tensor_to_optimizer :266-279 (moments of the group zeroed)."""
    with torch.no_grad():
        op = get_opacity(model)
        x = torch.min(op, torch.ones_like(op) * 0.01)
        new = torch.log(x / (1 - x))                                # inverse_sigmoid, general_utils.py:21-22
        for group, p, st in _groups(model):
            if group["name"] == "opacity":
                avg = sq = None
                if _has_moments(st):
                    avg, sq = torch.zeros_like(new), torch.zeros_like(new)
                _install(model, group, p, new, st, avg, sq)
