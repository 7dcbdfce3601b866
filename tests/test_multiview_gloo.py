"""CPU, world_size 2, gloo: the view-sharded step (flat gradient bucket all-reduce + radii max +
densification statistics) equals the single-process loop over all views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianip_b200 import multiview

P, V = 257, 6


def _make_params():
    g = torch.Generator().manual_seed(0)
    return {"xyz": torch.randn(P, 3, generator=g), "features_dc": torch.randn(P, 1, 3, generator=g),
            "opacity": torch.randn(P, 1, generator=g), "scaling": torch.randn(P, 3, generator=g),
            "rotation": torch.randn(P, 4, generator=g)}


def _render(params, v, vsp):
    """Stand-in for the CUDA op with the same autograd structure: depends on every parameter and
    on the zero-valued screen-space carrier; per-view radii."""
    g = torch.Generator().manual_seed(100 + v)
    w = torch.randn(P, 3, generator=g)
    img = (params["xyz"] * w).sum(1) * torch.sigmoid(params["opacity"][:, 0]) \
        + (torch.exp(params["scaling"]) * w).sum(1) + params["features_dc"][:, 0, :].sum(1) * (v + 1) \
        + torch.nn.functional.normalize(params["rotation"])[:, 0] + (vsp * w * (v + 2)).sum(1)
    radii = (torch.rand(P, generator=g) * 30).to(torch.int32) * (torch.rand(P, generator=g) > 0.3)
    return {"render": img, "radii": radii.to(torch.int32)}


def _loss(v, out):
    return (out["render"] ** 2).sum() * (0.5 + v)


def _run_step(group=None):
    params = {k: t.clone().requires_grad_(True) for k, t in _make_params().items()}
    vp = multiview.ViewParallel(params, P, group)
    res = vp.step(V, lambda v, vsp: _render(params, v, vsp), _loss)
    acc, den, mr = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
    vis = multiview.add_densification_stats(acc, den, mr, res["viewspace_grad"], res["radii"])
    return vp, res, acc, den, mr, vis


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vp, res, acc, den, mr, vis = _run_step()
        assert res["local_views"] == list(range(rank, V, world))
        ret[rank] = {"flat": vp.bucket.flat.clone(), "radii": res["radii"].clone(), "loss": res["loss"].clone(),
                     "acc": acc, "den": den, "mr": mr, "vis": vis}
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_two_rank_step_equals_single_process():
    vp1, res1, acc1, den1, mr1, vis1 = _run_step()           # world = 1: all views locally
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        got = ret[r]
        torch.testing.assert_close(got["flat"], vp1.bucket.flat, rtol=1e-4, atol=1e-4)
        assert torch.equal(got["radii"], res1["radii"])
        torch.testing.assert_close(got["loss"], res1["loss"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(got["acc"], acc1, rtol=1e-5, atol=1e-5)
        assert torch.equal(got["den"], den1) and torch.equal(got["mr"], mr1) and torch.equal(got["vis"], vis1)
    assert torch.equal(ret[0]["flat"], ret[1]["flat"])       # replicas stay bit-identical


def test_bucket_views_alias_leaf_grads():
    params = {k: t.clone().requires_grad_(True) for k, t in _make_params().items()}
    b = multiview.GradBucket(params, P)
    b.attach()
    assert b.nbytes() == 4 * (P * (3 + 3 + 1 + 3 + 4) + P * 3)
    (params["xyz"].sum() * 2 + (b.viewspace_points * 3).sum()).backward()
    (params["xyz"].sum() * 5).backward()                      # accumulates in place into the bucket
    assert params["xyz"].grad.data_ptr() == b.views["xyz"].data_ptr()
    assert float(b.views["xyz"].min()) == 7.0 and float(b.viewspace_grad().max()) == 3.0
    b.zero_()
    assert float(params["xyz"].grad.abs().max()) == 0.0


def test_densification_stats_match_reference_formulas():
    """gaussian_model.py:420-422 and GaussianIP.py:456 on explicit tensors."""
    g = torch.Generator().manual_seed(3)
    grad = torch.randn(P, 3, generator=g)
    radii = ((torch.rand(P, generator=g) * 40).to(torch.int32)) * (torch.rand(P, generator=g) > 0.5)
    acc, den, mr = torch.rand(P, 1, generator=g), torch.ones(P, 1), torch.rand(P, generator=g) * 20
    acc0, den0, mr0 = acc.clone(), den.clone(), mr.clone()
    vis = multiview.add_densification_stats(acc, den, mr, grad, radii.to(torch.int32))
    f = radii > 0
    exp_acc, exp_den, exp_mr = acc0.clone(), den0.clone(), mr0.clone()
    exp_acc[f] += torch.norm(grad[f, :2], dim=-1, keepdim=True)
    exp_den[f] += 1
    exp_mr[f] = torch.max(mr0[f], radii[f].float())
    assert torch.equal(vis, f)
    torch.testing.assert_close(acc, exp_acc)
    assert torch.equal(den, exp_den) and torch.equal(mr, exp_mr)


def test_balanced_sharding_partitions_and_balances():
    import random
    from gaussianip_b200.multiview import shard_views_balanced, view_cost_proxy
    rnd = random.Random(3)
    for world, per in ((2, 4), (8, 4), (4, 3), (8, 1)):
        costs = [view_cost_proxy(rnd.uniform(1.3, 1.7), rnd.uniform(0.7, 1.22)) for _ in range(world * per)]
        parts = [shard_views_balanced(costs, r, world) for r in range(world)]
        assert sorted(v for p in parts for v in p) == list(range(world * per))       # a partition
        assert all(len(p) == per for p in parts)                                      # same number of views each
        loads = [sum(costs[v] for v in p) for p in parts]
        naive = [sum(costs[v] for v in range(r, world * per, world)) for r in range(world)]
        assert max(loads) <= max(naive) + 1e-12
        if per >= 3:
            assert max(loads) / (sum(loads) / world) < 1.15
    assert shard_views_balanced([1.0, 1.0, 1.0], 0, 2) == [0]                         # ties broken by index,
    assert shard_views_balanced([1.0, 1.0, 1.0], 1, 2) == [1, 2]                      # dealt in snake order
    assert shard_views_balanced([], 0, 2) == []


def test_exchange_plan_covers_every_row_once():
    """Host-side planning of the fused exchange (exchange.py): field offsets are 16-byte aligned and the
    ownership blocks of all ranks tile every field exactly, warp-aligned, for any rank count."""
    from gaussianip_b200.exchange import plan_layout, plan_ownership
    present = {"means3D": True, "means2D": True, "opacities": True, "shs": True, "colors": False, "scales": True,
               "rotations": True, "cov3D": False}
    for P, K in ((1_000_000, 1), (50_001, 4), (33, 16), (7, 1)):
        layout, total = plan_layout(P, K, present)
        assert set(layout) == {"means3D", "means2D", "opacities", "shs", "scales", "rotations"}
        spans = sorted((off, off + int(torch.Size(shape).numel())) for off, shape in layout.values())
        assert all(off % 64 == 0 for off, _ in spans) and spans[-1][1] <= total
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))                 # fields do not overlap
        for world in (1, 2, 3, 4, 8, 16):
            covered = {name: [] for name in layout}
            rprs = set()
            for rank in range(world):
                rpr, segs = plan_ownership(P, layout, rank, world)
                rprs.add(rpr)
                assert rpr % 32 == 0
                for off, cnt in segs:
                    assert off % 4 == 0, "blocks must start 16-byte aligned for the vector stores"
                    name = max((n for n in layout if layout[n][0] <= off), key=lambda n: layout[n][0])
                    covered[name].append((off - layout[name][0], cnt))
            assert len(rprs) == 1
            # the kernel's owner rule (csrc/preprocess_bwd.cu, MODE 3): owner = min(first_row_of_warp / rpr, world - 1)
            rpr = rprs.pop()
            w3 = 3                                                   # means3D rows are 3 floats wide
            blocks = [plan_ownership(P, {"means3D": layout["means3D"]}, r, world)[1] for r in range(world)]
            for first in range(0, P, 32):
                owner = min(first // rpr, world - 1)
                (off, cnt), = blocks[owner] or [(None, 0)]
                assert off is not None and off <= layout["means3D"][0] + first * w3 < off + cnt, (P, world, first)
            for name, (off, shape) in layout.items():
                n = int(torch.Size(shape).numel())
                pieces = sorted(covered[name])
                assert pieces[0][0] == 0 and sum(c for _, c in pieces) == n
                assert all(a[0] + a[1] == b[0] for a, b in zip(pieces, pieces[1:]))  # contiguous, no gaps / overlap


def _run_step_batched(group=None):
    params = {k: t.clone().requires_grad_(True) for k, t in _make_params().items()}
    vp = multiview.ViewParallel(params, P, group)

    def render_views_fn(views, vsp):
        outs = [_render(params, v, vsp) for v in views]
        return {"render": torch.stack([o["render"] for o in outs]),
                "radii": torch.stack([o["radii"] for o in outs]).max(dim=0).values}

    def loss_fn(views, out):
        return sum((out["render"][i] ** 2).sum() * (0.5 + v) for i, v in enumerate(views))

    res = vp.step_batched(V, render_views_fn, loss_fn)
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    return res, grads


def _worker_batched(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res, grads = _run_step_batched()
        ret[rank] = {"grads": grads, "vg": res["viewspace_grad"].clone(), "radii": res["radii"].clone(),
                     "loss": res["loss"].clone()}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_batched_step_direct_gradients_and_two_ranks_agree_with_the_per_view_loop():
    """step_batched on one rank adopts the backward's tensors (no bucket); on two ranks it fills and all-reduces
    the bucket: both must equal the per-view step()."""
    vp1, res1, *_ = _run_step()                               # per-view loop, bucket
    ref = {k: vp1.bucket.views[k].clone() for k in vp1.bucket.params}
    res_b, grads_b = _run_step_batched()                      # world 1: direct mode
    for k in ref:
        torch.testing.assert_close(grads_b[k], ref[k], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(res_b["viewspace_grad"], res1["viewspace_grad"], rtol=1e-4, atol=1e-4)
    assert torch.equal(res_b["radii"], res1["radii"])
    torch.testing.assert_close(res_b["loss"], res1["loss"], rtol=1e-5, atol=1e-5)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_batched, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        for k in ref:
            torch.testing.assert_close(ret[r]["grads"][k], ref[k], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(ret[r]["vg"], res1["viewspace_grad"], rtol=1e-4, atol=1e-4)
        assert torch.equal(ret[r]["radii"], res1["radii"])
        torch.testing.assert_close(ret[r]["loss"], res1["loss"], rtol=1e-5, atol=1e-5)


def test_exchange_plan_with_radii_scalars_and_tiny_clouds():
    """The extras that ride in the fused exchange: the per-Gaussian radii (one int per row, owned like every other
    row block) and the scalar slots (owned by rank 0 as a whole).  With P <= 32 (world - 1) some ranks own no rows at
    all: their segment list is empty — they must still take part in every barrier (GradExchange.end launches
    nothing for them but always meets the closing barrier)."""
    from gaussianip_b200.exchange import SCALAR_SLOTS, plan_layout, plan_ownership
    present = {"means3D": True, "means2D": True, "opacities": True, "shs": True, "scales": True, "rotations": True,
               "radii": True, "scalars": True}
    for P, world in ((5, 4), (40, 8), (1000, 2)):
        layout, total = plan_layout(P, 1, present)
        assert layout["radii"][1] == (P, 1) and layout["scalars"][1] == (SCALAR_SLOTS,)
        owners_of_scalars, empty_ranks = [], 0
        rows_seen = 0
        for rank in range(world):
            rpr, segs = plan_ownership(P, layout, rank, world)
            names = []
            for off, cnt in segs:
                name = max((n for n in layout if layout[n][0] <= off), key=lambda n: layout[n][0])
                names.append(name)
                assert off + cnt <= total
                if name == "radii":
                    rows_seen += cnt
            if "scalars" in names:
                owners_of_scalars.append(rank)
            if not [n for n in names if n != "scalars"]:
                empty_ranks += 1
        assert owners_of_scalars == [0]
        assert rows_seen == P                                   # every Gaussian's radius has exactly one owner
        if P <= 32 * (world - 1):
            assert empty_ranks >= 1
