"""View-sharded data parallelism over the B200s of one box.

The path shards by CAMERA VIEW: every rank holds a full replica of the Gaussian parameters,
view ``v`` of a step goes to rank ``v mod world`` (the reference simply loops over the views
on one GPU: threestudio/systems/GaussianIP.py:154-159, 305-307).  There is exactly one
exchange step per optimiser step:

1. ``all_reduce(SUM)`` of ONE flat fp32 bucket that holds every parameter gradient AND the
   summed screen-space gradient (the reference sums ``viewspace_points.grad`` over views
   BEFORE taking the norm, GaussianIP.py:450-457, so the sum is what must be reduced);
   autograd accumulates straight into views of that bucket, so nothing is packed or copied
   before NCCL;
2. ``all_reduce(MAX)`` of ``radii`` (-> visibility_filter, max_radii2D; GaussianIP.py:161-167, 456).

Afterwards the densification statistics (gaussian_model.py:420-422) are updated identically
on every rank, so replicas stay bit-identical without further traffic.

Nothing in this module touches CUDA directly: it works on whatever device the tensors live
on, which is how the world_size-2 ``gloo`` tests exercise it on CPU.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Interleaved assignment v -> v mod world (neighbouring azimuths differ little in cost,
    so interleaving balances front / side views across ranks)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def shard_views_balanced(costs: Sequence[float], rank: int, world: int) -> List[int]:
    """Cost-aware assignment for a step whose views are known to every rank (same sampler seed): views are
    sorted by predicted cost and dealt out in snake order (0..W-1, W-1..0, ...), so every rank gets the same
    NUMBER of views (+-1) and nearly the same total cost.  A synchronous step lasts as long as its slowest
    rank; with random orbit cameras (distance 1.3-1.7, fovy 40-70 deg: projected area varies ~6x between
    views) the interleaved assignment leaves 10-20 % of the step to that skew.  Deterministic, so all ranks
    compute the same partition without communicating."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    order = sorted(range(len(costs)), key=lambda v: (-float(costs[v]), v))
    mine = []
    for pos, v in enumerate(order):
        rnd, k = divmod(pos, world)
        owner = k if rnd % 2 == 0 else world - 1 - k
        if owner == rank:
            mine.append(v)
    return sorted(mine)


def view_cost_proxy(distance: float, fovy: float) -> float:
    """Relative blend cost of an orbit view of an object at the origin: projected area ~ 1 / (d tan(fovy/2))^2."""
    import math
    return 1.0 / (distance * math.tan(0.5 * fovy)) ** 2


class GradBucket:
    """One flat fp32 buffer; named views are installed as ``.grad`` of the leaf parameters."""

    VIEWSPACE = "__viewspace__"

    def __init__(self, params: Dict[str, torch.Tensor], n_points: int):
        if not params:
            raise ValueError("no parameters")
        first = next(iter(params.values()))
        self.device = first.device
        self.names = list(params) + [self.VIEWSPACE]
        shapes = {k: tuple(v.shape) for k, v in params.items()}
        shapes[self.VIEWSPACE] = (n_points, 3)
        self.shapes = shapes
        sizes = [int(torch.Size(shapes[k]).numel()) for k in self.names]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=self.device)
        self.views = {k: self.flat[self.offsets[i]:self.offsets[i + 1]].view(shapes[k])
                      for i, k in enumerate(self.names)}
        self.params = params
        # shared zero-valued carrier for the screen-space gradient of every local view
        self.viewspace_points = torch.zeros(n_points, 3, dtype=torch.float32, device=self.device,
                                            requires_grad=True)

    def attach(self) -> None:
        """Point every leaf's .grad at its bucket view (autograd then accumulates in place)."""
        for k, p in self.params.items():
            if p.dtype != torch.float32:
                raise ValueError(f"{k}: bucket holds fp32 gradients only")
            p.grad = self.views[k] if p.numel() else torch.zeros_like(p)
        self.viewspace_points.grad = self.views[self.VIEWSPACE]

    def zero_(self) -> None:
        self.flat.zero_()

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def all_reduce(self, group=None, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    def viewspace_grad(self) -> torch.Tensor:
        return self.views[self.VIEWSPACE]


def all_reduce_radii_max(radii: torch.Tensor, group=None) -> torch.Tensor:
    """radii: per-rank max over local views, int32 [P]; reduced in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group)
    return radii


def add_densification_stats(xyz_gradient_accum: torch.Tensor, denom: torch.Tensor, max_radii2D: torch.Tensor,
                            viewspace_grad: torch.Tensor, radii: torch.Tensor) -> torch.Tensor:
    """GaussianIP.py:456-457 + gaussian_model.py:420-422 on the REDUCED quantities.  Returns the
    visibility filter.  Mask-multiply form (no boolean indexing -> no device sync)."""
    vis = radii > 0
    m = vis.to(xyz_gradient_accum.dtype)
    xyz_gradient_accum.add_((torch.norm(viewspace_grad[:, :2], dim=-1, keepdim=True)) * m[:, None])
    denom.add_(m[:, None])
    torch.maximum(max_radii2D, torch.where(vis, radii.to(max_radii2D.dtype), max_radii2D), out=max_radii2D)
    return vis


class ViewParallel:
    """Runs one optimisation step's worth of views, sharded over the process group.

    render_fn(view_index, viewspace_points) -> dict with at least 'radii' and whatever loss_fn
    needs; loss_fn(view_index, render_dict) -> scalar loss of that view.
    """

    def __init__(self, params: Dict[str, torch.Tensor], n_points: int, group=None, fused_exchange: bool = False,
                 exchange_algorithm: str = "auto"):
        """fused_exchange: reduce the gradients inside the backward kernel over NVLS multicast memory
        (gaussianip_b200.exchange.GradExchange) instead of all-reducing the bucket with NCCL afterwards.
        Only step_batched uses it; render_views_fn then receives the exchange as third argument and must
        pass it on to render_views / rasterize_views."""
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket = GradBucket(params, n_points)
        self.n_points = n_points
        self.exchange = None
        if fused_exchange and self.world > 1:
            from .exchange import GradExchange
            # leaves accumulate into the bucket views, so nothing keeps the symmetric buffer's views
            self.exchange = GradExchange(group, clone_outputs=False, algorithm=exchange_algorithm)

    def _attempt(self, body: Callable):
        """Run one whole step speculatively (rasterizer.speculation): the forwards inside do not stall
        the host on the per-view instance count; if a view turns out not to have fitted its capacity the
        step's results are discarded and the step is run again with the raised capacity.  Local only —
        there is no collective inside, so ranks need not agree on the number of attempts."""
        from . import rasterizer
        if rasterizer._active_speculation() is not None:
            # the caller owns the step (e.g. graph.CapturedStep captures / replays it and validates the counts)
            return body()
        for _ in range(8):
            with rasterizer.speculation() as spec:
                result = body()
            if spec.validate():
                return result
        raise RuntimeError("instance capacity overflowed on 8 consecutive attempts")

    def step(self, n_views: int, render_fn: Callable, loss_fn: Callable, views: Optional[Sequence[int]] = None):
        b = self.bucket
        local = shard_views(n_views, self.rank, self.world) if views is None else list(views)

        def body():
            b.zero_()
            b.attach()
            radii = torch.zeros(self.n_points, dtype=torch.int32, device=b.device)
            total = torch.zeros((), dtype=torch.float32, device=b.device)
            for v in local:
                out = render_fn(v, b.viewspace_points)
                radii = torch.maximum(radii, out["radii"])
                loss = loss_fn(v, out)
                loss.backward()            # accumulates into the bucket views, view after view
                total = total + loss.detach()
            return radii, total

        radii, total = self._attempt(body)
        b.all_reduce(self.group)
        all_reduce_radii_max(radii, self.group)
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return {"loss": total, "radii": radii, "viewspace_grad": b.viewspace_grad(), "local_views": local}

    def step_batched(self, n_views: int, render_views_fn: Callable, loss_fn: Callable,
                     views: Optional[Sequence[int]] = None):
        """Same exchange, but the local views go through ONE batched render call:
        render_views_fn(local_view_indices, viewspace_points) -> dict with 'radii' (max over the
        local views) and stacked outputs; loss_fn(local_view_indices, dict) -> scalar."""
        b = self.bucket
        local = shard_views(n_views, self.rank, self.world) if views is None else list(views)
        # Nothing has to be summed on this rank after the single backward call when there is no NCCL bucket to
        # fill (one rank, or the exchange already happened inside the kernel): leave .grad empty so autograd
        # simply adopts the tensors the backward returns — no bucket zeroing, no accumulate kernels.  The
        # gradients are then separate tensors (with the fused exchange: views of its buffer, valid until the
        # next step's backward) instead of views of the flat bucket.
        direct = self.exchange is not None or self.world == 1

        def body():
            if direct:
                for p in b.params.values():
                    p.grad = None
                b.viewspace_points.grad = None
            else:
                b.zero_()
                b.attach()
            if not local:
                return (torch.zeros(self.n_points, dtype=torch.int32, device=b.device),
                        torch.zeros((), dtype=torch.float32, device=b.device))
            if self.exchange is not None:
                out = render_views_fn(local, b.viewspace_points, self.exchange)
            else:
                out = render_views_fn(local, b.viewspace_points)
            loss = loss_fn(local, out)
            if self.exchange is not None:
                # the step's loss is summed over ranks by the fused kernel too (no NCCL call in the step)
                self.exchange.pending_scalar = loss.detach().to(torch.float32).reshape(1)
            loss.backward()
            return out["radii"], loss.detach().to(torch.float32)

        if self.exchange is not None:
            # the backward contains cross-rank barriers, so it must run exactly once per step on every rank:
            # no speculative redo (the forward validates its instance counts before returning), and every
            # rank needs at least one local view
            if not local:
                raise ValueError("fused exchange needs at least one view per rank")
            radii, total = body()
            # radii maximum and loss sum came out of the fused kernel (exchange.set_aux): no NCCL in the step
            radii = self.exchange.local("radii").view(torch.int32).reshape(-1).clone()
            total = self.exchange.local("scalars")[0].clone()
        else:
            radii, total = self._attempt(body)
            b.all_reduce(self.group)
            all_reduce_radii_max(radii, self.group)
            if self.world > 1:
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        if direct:
            vg = b.viewspace_points.grad
            vg = vg if vg is not None else torch.zeros_like(b.viewspace_points)
        else:
            vg = b.viewspace_grad()
        # "grads": the tensors that hold this step's parameter gradients (what .grad points at right now); a caller
        # that replays the step from a CUDA graph reads them from here, because .grad follows the LAST capture
        return {"loss": total, "radii": radii, "viewspace_grad": vg, "local_views": local,
                "grads": {k: p.grad for k, p in b.params.items()}}
