#!/usr/bin/env python
"""Headline benchmark: fwd+bwd views/s of the Gaussian-splatting rasterizer at 1 M Gaussians,
4 x 1024^2 AHDS orbit views per step per GPU (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one optimisation step's worth of rendering: every rank renders `views` random
orbit views of the shared 1 M-Gaussian humanoid (colour + depth + alpha), back-propagates a
dense synthetic loss through the operator and the parameter activations, then (N > 1) the
flat gradient bucket and the radii are all-reduced with NCCL.  `value` = views of all ranks /
device time (CUDA events, max over ranks), inputs resident in HBM.  `e2e` = the same step
through the public render() API starting from pinned HOST parameter buffers, with the
host->device copies and the device->host loss read inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fwd+bwd views/s @1024^2, 1M Gaussians (4-view AHDS batch per GPU)"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--views", type=int, default=4, help="views per step per GPU")
    ap.add_argument("--sh-degree", type=int, default=0)
    ap.add_argument("--variant", default="native", choices=["native", "standin", "packed_bwd"],
                    help="standin = reference-STRUCTURE kernels of csrc/standin.cu + 64-bit key sort + per-view "
                         "Python loop, for context only (never the reference, never the product); packed_bwd = "
                         "EXPERIMENTAL backward with a packed shared-memory reduction (parity-checked on one scene, not benchmarked)")
    ap.add_argument("--view-sharding", default="interleaved", choices=["balanced", "interleaved"],
                    help="N>1: how the step's world x views cameras are dealt to the ranks")
    ap.add_argument("--exchange-algo", default="auto", choices=["auto", "push_all", "owner_push"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "fused"],
                    help="N>1: all-reduce the gradient bucket with NCCL, or reduce inside the backward kernel "
                         "over NVLink peer / NVLS multicast memory (gaussianip_b200/exchange.py); auto = fused "
                         "when every rank can map multicast memory, else NCCL")
    ap.add_argument("--torch-activations", action="store_true",
                    help="evaluate sigmoid/exp/normalize with torch around the operator (the reference's getters) "
                         "instead of inside the per-Gaussian kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    return ap.parse_args()


def workload_name(a):
    return (f"AHDS stage-1 shape: {a.points} Gaussians (capsule humanoid, seed 0), SH deg {a.sh_degree}, "
            f"{a.views} random orbit views/GPU/step at {a.res}x{a.res}, colour+depth+alpha fwd+bwd")


# ---- clocks ------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- reference arm: the oracle's CPU evaluation of the identical math -------------------------

def cpu_reference_run(a, steps, warmup, budget_s):
    """Times oracle/splat_torch.py (pure PyTorch fp32, all host threads) on a bounded sample of
    the bench workload: per step one view, full preprocess + binning of all Gaussians, blending
    + autograd backward of every `stride`-th tile row.  views/s = (fraction of the view's
    instances that were blended) / step time."""
    from oracle import splat_torch as O
    from gaussianip_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cl = synthetic.make_cloud(a.points, a.sh_degree, 0)
    cams = synthetic.ahds_cameras(a.views, a.res, a.res, seed=1, device="cpu")
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(3, a.res, a.res, generator=g), torch.randn(1, a.res, a.res, generator=g),
         torch.randn(1, a.res, a.res, generator=g))
    tile_rows_total = (a.res + 15) // 16

    def one(view, stride, phase):
        cam = cams[view % len(cams)]
        st = O.Settings(a.res, a.res, cam.tanfovx, cam.tanfovy, torch.zeros(3), 1.0, cam.world_view_transform,
                        cam.full_proj_transform, a.sh_degree, cam.camera_center)
        leaves = [t.clone().requires_grad_(True) for t in (cl.xyz, cl.features_dc, cl.features_rest, cl.scaling,
                                                            cl.rotation, cl.opacity)]
        xyz, fdc, frest, sc, rot, op = leaves
        m2d = torch.zeros_like(xyz, requires_grad=True)
        t0 = time.perf_counter()
        out = O.rasterize(st, xyz, m2d, torch.sigmoid(op), shs=torch.cat((fdc, frest), 1), scales=torch.exp(sc),
                          rotations=torch.nn.functional.normalize(rot), return_aux=True,
                          tile_rows=(phase, stride))
        color, radii, depth, alpha, geom, binning, img = out
        ((color * w[0]).sum() + (depth * w[1]).sum() + (alpha * w[2]).sum()).backward()
        dt = time.perf_counter() - t0
        rng = binning.ranges
        rows = torch.arange(rng.shape[0]) // binning.grid[0]
        sel = (rows % stride) == phase
        frac = float((rng[:, 1] - rng[:, 0])[sel.numpy()].sum()) / max(1, len(binning.keys))
        return dt, frac

    # Calibration: t_step = t_const + t_var * frac, where t_const is the per-view work that does not
    # shrink with the tile sample (projection, binning and their backward over all Gaussians) and
    # frac the blended share of the view's instances.  Two probes with different strides give both.
    sa, sb = max(2, tile_rows_total // 4), max(1, tile_rows_total // 8)
    ta, fa = one(0, sa, 0)
    tb, fb = one(0, sb, 0)
    t_var = max(1e-6, (tb - ta) / max(fb - fa, 1e-6))
    t_const = min(max(0.0, ta - t_var * fa), ta)
    est_full = t_const + t_var
    per_step_budget = budget_s / max(1, steps + warmup)
    stride = 1
    while stride < tile_rows_total and t_const + t_var / stride > per_step_budget:
        stride *= 2
    times, fulls, fracs = [], [], []
    for i in range(warmup + steps):
        dt, frac = one(i, stride, i % stride)
        if i >= warmup:
            times.append(dt); fracs.append(frac)
            fulls.append(t_const + max(0.0, dt - t_const) / max(frac, 1e-6) if stride > 1 else dt)
    vps = len(fulls) / sum(fulls)
    sample = (f"{steps} steps x 1 view: full preprocess+binning of {a.points} Gaussians, blend+backward of every "
              f"{stride}-th tile row ({100 * sum(fracs) / len(fracs):.1f}% of the view's instances per step), "
              f"{sum(times):.1f} s CPU wall; full-view time per step = t_const + (t_step - t_const) / blended fraction "
              f"with t_const = {t_const:.2f} s from a two-stride calibration (estimated full view {est_full:.1f} s); "
              f"views/s = steps / sum of full-view times")
    return vps, cores, sample, sum(fulls) / len(fulls) * 1e3


# ---- our arm -----------------------------------------------------------------------------------------------

def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        steps, warmup = max(1, a.steps), max(0, min(a.warmup, 2))
        vps, cores, sample, ms = cpu_reference_run(a, steps, warmup, a.cpu_budget_s)
        line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": a.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "points": a.points, "resolution": a.res,
                           "views_per_step_per_gpu": a.views, "sh_degree": a.sh_degree},
                "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "the reference's rasterizer is an external CUDA-only dependency absent from the tree; "
                        "this arm is the CPU oracle port of its published algorithm (oracle/splat_torch.py)"}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gaussianip_b200 has no CPU fallback")
    import torch.distributed as dist
    from gaussianip_b200 import _lib, multiview, rasterizer, renderer, synthetic
    from gaussianip_b200.cameras import Camera, cameras_from_c2w, look_at_c2w, orbit_position
    import numpy as np

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- workload: shared cloud (same on all ranks), per-rank cameras and loss weights ----
    cloud_host = synthetic.make_cloud(a.points, a.sh_degree, 0)
    host = {k: getattr(cloud_host, k).pin_memory() for k in
            ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")}
    params = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}

    class Model:            # GaussianModel getters (gaussian_model.py:84-107) over a dict of leaves
        active_sh_degree = a.sh_degree

        def __init__(self, p):
            self.p = p
        _opacity = property(lambda s: s.p["opacity"])       # the raw parameters, reference attribute names
        _scaling = property(lambda s: s.p["scaling"])
        _rotation = property(lambda s: s.p["rotation"])
        get_xyz = property(lambda s: s.p["xyz"])
        get_features = property(lambda s: torch.cat((s.p["features_dc"], s.p["features_rest"]), dim=1))
        get_opacity = property(lambda s: torch.sigmoid(s.p["opacity"]))
        get_scaling = property(lambda s: torch.exp(s.p["scaling"]))
        get_rotation = property(lambda s: torch.nn.functional.normalize(s.p["rotation"]))

    # The step's GLOBAL batch of views (world x views-per-GPU random orbit cameras) is sampled identically on
    # every rank (common seed); each rank then takes its share: cost-balanced (default) or interleaved.
    rng = np.random.default_rng(1000)
    n_global = a.views * world

    def sample_cameras_host():
        """camera_data.py:349-364 distributions; returns this rank's (c2w, fovy) per view, host side."""
        specs, costs = [], []
        for i in range(n_global):
            az = (rng.random() + i) / n_global * 360.0 - 180.0
            el, dist_, fovy = rng.uniform(-30, 30), rng.uniform(1.3, 1.7), float(np.radians(rng.uniform(40, 70)))
            specs.append((look_at_c2w(orbit_position(az, el, dist_)), fovy))
            costs.append(multiview.view_cost_proxy(dist_, fovy))
        if a.view_sharding == "balanced":
            mine = multiview.shard_views_balanced(costs, rank, world)
        else:
            mine = multiview.shard_views(n_global, rank, world)
        return [specs[v] for v in mine]

    total_steps = a.warmup + a.steps
    cam_specs = [sample_cameras_host() for _ in range(total_steps)]
    gen = torch.Generator().manual_seed(2 + rank)
    weights = [tuple(torch.randn(c, a.res, a.res, generator=gen).to(dev) for c in (3, 1, 1)) for _ in range(a.views)]
    bg = torch.zeros(3, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    fused = world > 1 and a.exchange in ("auto", "fused")
    if fused:
        from gaussianip_b200.exchange import GradExchange
        if not GradExchange.available(dev):
            if a.exchange == "fused":
                raise SystemExit("--exchange fused: NVLS multicast symmetric memory is not available on this box")
            fused = False
    vp = multiview.ViewParallel(params, a.points, fused_exchange=fused, exchange_algorithm=a.exchange_algo)
    model = Model(params)
    standin = a.variant == "standin"
    if standin:
        rasterizer.set_blend_variant("standin")
        rasterizer.set_binning_mode("flat64", dev)
    elif a.variant == "packed_bwd":
        rasterizer.set_blend_variant("packed_bwd")

    w_color = torch.stack([w[0] for w in weights])
    w_depth = torch.stack([w[1] for w in weights])
    w_alpha = torch.stack([w[2] for w in weights])

    wc_flat, wd_flat, wa_flat = w_color.reshape(-1), w_depth.reshape(-1), w_alpha.reshape(-1)

    def loss_of(views, out):
        # L = sum_views <w_c, colour> + <w_d, depth> + <w_a, alpha>  (SURVEY.md §8d): dense, non-trivial
        # dL/dcolour, dL/ddepth, dL/dalpha; written as dot products (one reduction kernel each)
        return torch.dot(out["render"].reshape(-1), wc_flat) + torch.dot(out["depth_3dgs"].reshape(-1), wd_flat) + \
            torch.dot(out["alpha_3dgs"].reshape(-1), wa_flat)

    def loss_one(v, out):
        return torch.dot(out["render"].reshape(-1), w_color[v].reshape(-1)) + \
            torch.dot(out["depth_3dgs"].reshape(-1), w_depth[v].reshape(-1)) + \
            torch.dot(out["alpha_3dgs"].reshape(-1), w_alpha[v].reshape(-1))

    def run_step(step_idx, cams):
        if standin:     # the reference's structure: one render() call per view from a Python loop
            return vp.step(a.views, lambda v, vsp: renderer.render(cams[v], model, None, bg, screenspace_points=vsp),
                           loss_one, views=range(a.views))
        # public API: all views of the step in one batched render call (same kernels per view as
        # the single-view operator; activations evaluated once per step, one autograd node)
        def render_views_fn(views, vsp, exchange=None):
            return renderer.render_views([cams[v] for v in views], model, None, bg, screenspace_points=vsp,
                                         exchange=exchange, fused_activations=not a.torch_activations)
        return vp.step_batched(a.views, render_views_fn, loss_of, views=range(a.views))

    def device_cams(step_idx):
        spec = cam_specs[step_idx]
        return cameras_from_c2w([c for c, _ in spec], [f for _, f in spec], a.res, a.res, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM ------------------------------------------------------
    cams_all = [device_cams(i) for i in range(total_steps)]
    for i in range(a.warmup):
        run_step(i, cams_all[i])
    barrier()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    st0 = rasterizer.stats()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t_wall0 = time.perf_counter()
    host_s = 0.0
    for k in range(a.steps):
        flush_buf.fill_(k & 0xFF)               # L2 flush between timed steps (outside the event pair)
        ev[k][0].record()
        t_h = time.perf_counter()
        run_step(a.warmup + k, cams_all[a.warmup + k])
        host_s += time.perf_counter() - t_h     # host time to ENQUEUE the step (includes any host-side waits)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clock_info = clocks.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    my_ms = sum(step_ms)
    launches = _lib.launch_count() - launches0
    prof_conc = _lib.profile_read()
    _lib.profile_enable(False)
    # Per-kernel durations for the roofline: in the timed region the views of a step run on
    # separate streams, so their kernels overlap and a per-launch CUDA-event duration there is not
    # the kernel's own speed.  Re-run a few steps of the SAME workload with the views back to back
    # on one stream and take the stage times from those (reported separately from `value`).
    rasterizer.set_multistream(False)
    n_serial = max(3, a.steps // 4)
    for i in range(2):
        run_step(i, cams_all[i])
    barrier()
    _lib.profile_enable(True)
    for i in range(n_serial):
        run_step(i, cams_all[i % total_steps])
    barrier()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    rasterizer.set_multistream(True)
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    views_total = a.views * world * a.steps
    value = views_total / (total_ms * 1e-3)

    # ---- leg 2: end to end from pinned host buffers -----------------------------------------
    # Every step's parameters come from pinned HOST memory: one flat pinned buffer, uploaded with ONE async
    # copy per step on a copy stream into one of two flat device buffers while the previous step computes.
    # The leaves are then pointed at that buffer's views (a host-side pointer swap, no device copy), so the
    # only device work the upload adds is the DMA itself.
    copy_stream = torch.cuda.Stream(device=dev)
    names = list(params)
    sizes = [params[k2].numel() for k2 in names]
    offs = [0]
    for n_ in sizes:
        offs.append(offs[-1] + (n_ + 63) // 64 * 64)
    flat_host = torch.empty(offs[-1], dtype=torch.float32).pin_memory()
    for k2, o_ in zip(names, offs):
        flat_host[o_:o_ + host[k2].numel()].copy_(host[k2].reshape(-1))
    slots = []
    for _ in range(2):
        buf = torch.empty(offs[-1], dtype=torch.float32, device=dev)
        slots.append({"buf": buf, "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                      "views": {k2: buf[o_:o_ + params[k2].numel()].view(params[k2].shape) for k2, o_ in zip(names, offs)}})
    for sl in slots:
        sl["free"].record()
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def prefetch(slot):
        sl = slots[slot]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])          # the step that last read this buffer has been enqueued
            sl["buf"].copy_(flat_host, non_blocking=True)   # H2D of every parameter of the step, one DMA
            sl["ready"].record()

    def e2e_step(step_idx, last):
        cur = torch.cuda.current_stream(dev)
        slot = step_idx & 1
        sl = slots[slot]
        cur.wait_event(sl["ready"])
        for k2 in names:
            params[k2].data = sl["views"][k2]           # pointer swap; the previous buffer is free from here on
        slots[1 - slot]["free"].record()
        if not last:
            prefetch(1 - slot)                          # next step's upload overlaps this step's kernels
        cams = device_cams(step_idx)                    # camera matrices built on host, one async upload
        out = run_step(step_idx, cams)
        loss_host.copy_(out["loss"].reshape(1), non_blocking=True)   # D2H read of the step's result
        return out

    n_e2e_warm = min(3, a.warmup)
    prefetch(0)
    for i in range(n_e2e_warm):
        e2e_step(i, False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        e2e_step(n_e2e_warm + k, k == a.steps - 1)     # consecutive indices: the two buffers strictly alternate
    e1.record()
    barrier()
    e2e_loss = float(loss_host.item())
    assert e2e_loss == e2e_loss, "e2e loss is NaN"
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = views_total / (float(t2.item()) * 1e-3)
    h2d = flat_host.numel() * 4 + a.views * (16 + 16 + 16 + 3) * 4
    d2h = 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    st1 = rasterizer.stats()
    D = (st1["num_rendered_sum"] - st0["num_rendered_sum"]) / max(1, st1["views"] - st0["views"])   # mean D per view
    HW, P, K = a.res * a.res, a.points, (a.sh_degree + 1) ** 2
    alg_bytes = {   # SURVEY.md §8(d) per-stage algorithmic bytes
        "preprocess_fwd": P * (44 + 12 * K + 48), "depth_sort": 4 * 16 * P, "scan_emit": 8 * P + 20 * P + 12 * D,
        "tile_sort": 2 * 16 * D, "ranges": 8 * D, "render_fwd": 44 * D + 28 * HW,
        "render_bwd": 44 * D + 28 * HW + 40 * P, "preprocess_bwd": P * (40 + 44 + 12 * K + 56 + 12 * K)}
    stage_ms = {k: (m / c if c else 0.0) for k, (m, c) in prof.items()}
    dom = max(stage_ms, key=lambda k: stage_ms[k] * prof[k][1])
    achieved = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    b_view = P * (300 + 36 * K) + 260 * D + 56 * HW
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "avg_launch_ms": stage_ms[dom],
                "whole_view": {"algorithmic_bytes": b_view, "achieved_gbs": b_view * value / world / 1e9,
                               "frac": b_view * value / world / 1e9 / peak},
                "secondary": {"bound": "instruction issue (not HBM): ncu 2.9-3.1 of 4 inst/cycle while active",
                              "pair_evals_upper_bound_per_s": 256 * D / (stage_ms[dom] * 1e-3)
                              if stage_ms[dom] > 0 else None},
                "sort": {"depth_sort_keys_per_s": (P * 4 / (stage_ms["depth_sort"] * 1e-3)) if stage_ms["depth_sort"] > 0 else None,
                         "tile_sort_keys_per_s": (D * 2 / (stage_ms["tile_sort"] * 1e-3)) if stage_ms["tile_sort"] > 0 else None,
                         "note": "keys x 8-bit passes per second: 32-bit depth keys of the P Gaussians (4 passes) and "
                                 "tile ids of the D instances (2 passes at 4096 tiles); same final order as one "
                                 "64-bit (tile|depth) sort of D keys"},
                "stage_us_per_view": {k: round(v * 1e3, 1) for k, v in stage_ms.items()},
                "stage_us_per_view_overlapped": {k: round(m / c * 1e3, 1) if c else 0.0 for k, (m, c) in prof_conc.items()},
                "note": "avg_launch_ms / stage_us_per_view: CUDA events inside bench.py on a serialised pass of the same "
                        "workload (views back to back on one stream); in the timed region the 4 views of a step run "
                        "on 4 streams and overlap (stage_us_per_view_overlapped), which is what `value` measures"}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        vps, cores, sample, _ = cpu_reference_run(a, 1, 0, 40.0)
        cpu = {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {"variant": "reference-STRUCTURE stand-in (csrc/standin.cu + flat 64-bit sort + per-view loop); NOT the "
                       "reference and NOT the product path"} if standin else (
        {"variant": "EXPERIMENTAL packed-reduction backward (gsb_set_blend_variant(2))"} if a.variant == "packed_bwd" else {})
    line.update({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "points": a.points, "resolution": a.res,
                       "views_per_step_per_gpu": a.views, "sh_degree": a.sh_degree, "num_rendered_D": D,
                       "l2": "256 MiB flush write between timed steps (outside the event pairs)",
                       "activations": "torch getters around the operator" if a.torch_activations else
                                      "fused into the per-Gaussian kernels (render_views(fused_activations=True))",
                       "view_sharding": "n/a" if world == 1 else a.view_sharding,
                       "parallelism": ("single GPU" if world == 1 else
                                       f"view-sharded dp{world}, gradients reduced INSIDE the backward kernel "
                                       f"({vp.exchange.algorithm} over NVLink peer / NVLS multicast memory, {vp.exchange.nbytes() >> 20} MiB) "
                                       f"+ NCCL radii max per step" if fused else
                                       f"view-sharded dp{world}, NCCL all-reduce of the flat gradient bucket "
                                       f"({vp.bucket.nbytes() >> 20} MiB) + radii max per step"),
                       "wall_s_timed_region": t_wall,
                       "host_enqueue_ms_per_step_rank0": host_s / a.steps * 1e3, "host_cores": os.cpu_count()},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu})
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
