#!/usr/bin/env python
"""Benchmarks of the Gaussian-splatting rasterizer hot path, one line of JSON per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config ahds|c3|vcr|playback]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Configurations (BASELINE.json `configs`):
  ahds      (default, the headline metric) 1 M Gaussians, SH 0, 4 random AHDS orbit views per GPU per step at
            1024^2, colour + depth + alpha forward + backward.  Weak scaling (every rank renders its own 4 views).
  c3        configs[2]: 3 M Gaussians, SH degree 3, otherwise as ahds.
  vcr       configs[3]: the 64 fixed VCR refinement views (GaussianIP.py:232-281) of one step, STRONG-sharded:
            rank r renders views r, r + N, ... and the gradients are summed over ranks once per step.
  playback  configs[4]: forward-only animation playback (animation.py:463-467): per frame new Gaussian positions,
            Renderer.render with a MiniCam, image to the host; frames/s.

A training "step" renders the rank's views through the public batched API (render_views), back-propagates a dense
synthetic loss and (N > 1) sums the gradients over ranks.  `value` = views of all ranks / device time (CUDA events,
max over ranks) with the inputs resident in HBM; `e2e` = the same step starting from pinned HOST buffers with the
host->device copies and the device->host result read inside the timed region.  By default the step is captured in a
CUDA graph (gaussianip_b200.graph) and replayed; `--graph off` enqueues it kernel by kernel.
In the default configuration every line additionally carries `vcr`: config 4 measured in the same process
(64 views / N per rank per step, strong scaling), so that a 1/2/4/8-GPU sweep also holds that series.
"""
from __future__ import annotations

import argparse
import json
import math
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

UNIT = "views/s"
CONFIGS = {
    "ahds": dict(points=1_000_000, sh_degree=0, views=4, cameras="ahds", scaling="weak",
                 metric="fwd+bwd views/s @1024^2, 1M Gaussians (4-view AHDS batch per GPU)"),
    "c3": dict(points=3_000_000, sh_degree=3, views=4, cameras="ahds", scaling="weak",
               metric="fwd+bwd views/s @1024^2, 3M Gaussians SH deg 3 (4-view batch per GPU)"),
    "vcr": dict(points=1_000_000, sh_degree=0, views=64, cameras="vcr", scaling="strong",
                metric="fwd+bwd views/s @1024^2, 1M Gaussians (64 VCR views per step, sharded over the GPUs)"),
    "playback": dict(points=1_000_000, sh_degree=0, views=1, cameras="playback", scaling="weak",
                     metric="forward-only playback frames/s @1024^2, 1M Gaussians (new positions per frame)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ahds", choices=sorted(CONFIGS))
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--views", type=int, default=None, help="ahds/c3: views per step per GPU; vcr: views per step in total")
    ap.add_argument("--sh-degree", type=int, default=None)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (auto: on, falling back to eager enqueue if capture fails)")
    ap.add_argument("--variant", default="native", choices=["native", "standin", "replay_bwd", "rescan_bwd", "rescan_packed_bwd"],
                    help="standin = reference-STRUCTURE kernels of csrc/standin.cu + 64-bit key sort + per-view "
                         "Python loop, for context only (never the reference, never the product)")
    ap.add_argument("--view-sharding", default="interleaved", choices=["balanced", "interleaved"],
                    help="N>1, ahds/c3: how the step's world x views cameras are dealt to the ranks")
    ap.add_argument("--exchange-algo", default="auto", choices=["auto", "push_all", "owner_push"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "fused"],
                    help="N>1: all-reduce the gradient bucket with NCCL, or reduce inside the backward kernel "
                         "over NVLink peer / NVLS multicast memory (gaussianip_b200/exchange.py); auto = fused "
                         "when every rank can map multicast memory, else NCCL")
    ap.add_argument("--torch-activations", action="store_true",
                    help="evaluate sigmoid/exp/normalize with torch around the operator (the reference's getters) "
                         "instead of inside the per-Gaussian kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vcr", action="store_true", help="default config: skip the additional config-4 measurement")
    ap.add_argument("--no-exchange-check", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    for k, v in (("points", a.points), ("views", a.views), ("sh_degree", a.sh_degree)):
        if v is not None:
            cfg[k] = v
    a.cfg = cfg
    return a


def workload_name(a, cfg=None, name=None):
    cfg = a.cfg if cfg is None else cfg
    name = a.config if name is None else name
    P, sh, V = cfg["points"], cfg["sh_degree"], cfg["views"]
    if name == "vcr":
        return (f"VCR stage-2 shape: {P} Gaussians (capsule humanoid, seed 0), SH deg {sh}, the {V} fixed refinement "
                f"views (elevation 17 deg, distance 1.5, fovy 70 deg) of one step at {a.res}x{a.res} dealt to the "
                f"ranks, colour+depth+alpha fwd+bwd, one gradient sum over ranks per step")
    if name == "playback":
        return (f"animation playback shape: {P} Gaussians, SH deg {sh}, forward only, one {a.res}x{a.res} MiniCam "
                f"frame per step with new positions per frame, image to the host")
    return (f"AHDS stage-1 shape: {P} Gaussians (capsule humanoid, seed 0), SH deg {sh}, "
            f"{V} random orbit views/GPU/step at {a.res}x{a.res}, colour+depth+alpha fwd+bwd")


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

def cpu_reference_run(a, cfg, steps, warmup, budget_s):
    """Times oracle/splat_torch.py (pure PyTorch fp32, all host threads) on a bounded sample of the workload.

    One step = one view: projection + binning of ALL Gaussians and their backward are run and timed in full; the
    per-tile blend and its backward run on every `stride`-th tile row (phase rotating with the step) and their time
    is scaled by instances-of-the-view / instances-blended.  The three phases are timed separately (the blend works
    on detached copies of the per-Gaussian records, whose gradients are then pushed through the projection), so no
    calibration fit is involved; stride 1 (a full view, nothing scaled) is used whenever the budget allows.
    forward-only configs skip the backward phases."""
    from oracle import splat_torch as O
    from gaussianip_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, sh, res = cfg["points"], cfg["sh_degree"], a.res
    fwd_only = cfg["cameras"] == "playback"
    cl = synthetic.make_cloud(P, sh, 0)
    if cfg["cameras"] == "vcr":
        cams = synthetic.vcr_cameras(cfg["views"], res, res, device="cpu")
    elif fwd_only:
        cams = synthetic.playback_cameras(136, res, res, device="cpu")
    else:
        cams = synthetic.ahds_cameras(max(4, cfg["views"]), res, res, seed=1, device="cpu")
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(3, res, res, generator=g), torch.randn(1, res, res, generator=g),
         torch.randn(1, res, res, generator=g))
    tile_rows_total = (res + 15) // 16
    geom_fields = ("xy", "conic", "opacity", "rgb", "depth")

    def one(view, stride, phase):
        cam = cams[(view * (7 if cfg["cameras"] == "vcr" else 1)) % len(cams)]
        st = O.Settings(res, res, cam.tanfovx, cam.tanfovy, torch.zeros(3), 1.0, cam.world_view_transform,
                        cam.full_proj_transform, sh, cam.camera_center)
        leaves = [t.clone().requires_grad_(not fwd_only) for t in (cl.xyz, cl.features_dc, cl.features_rest,
                                                                    cl.scaling, cl.rotation, cl.opacity)]
        xyz, fdc, frest, sc, rot, op = leaves
        if fwd_only:
            xyz = synthetic.playback_sway(xyz, view, 136)
        t0 = time.perf_counter()
        with torch.set_grad_enabled(not fwd_only):
            geom = O.preprocess(st, xyz, torch.sigmoid(op), shs=torch.cat((fdc, frest), 1), scales=torch.exp(sc),
                                rotations=torch.nn.functional.normalize(rot))
            binning = O.bin_and_sort(geom, res, res)
        t_pre = time.perf_counter() - t0
        # blend on detached copies of the per-Gaussian records (the interface between the two stages)
        import copy
        gd = copy.copy(geom)
        det = {}
        for f in geom_fields:
            det[f] = getattr(geom, f).detach().requires_grad_(not fwd_only)
            setattr(gd, f, det[f])
        t0 = time.perf_counter()
        with torch.set_grad_enabled(not fwd_only):
            img = O.blend_tiles(gd, binning, st.bg, res, res, tile_rows=(phase, stride) if stride > 1 else None)
            if not fwd_only:
                loss = (img.color * w[0]).sum() + (img.depth * w[1]).sum() + (img.alpha * w[2]).sum()
                if loss.grad_fn is not None:         # a sample whose tile rows hold no instance has nothing to blend
                    loss.backward()
        t_blend = time.perf_counter() - t0
        t_pre_bwd = 0.0
        if not fwd_only:
            t0 = time.perf_counter()
            outs = [getattr(geom, f) for f in geom_fields if det[f].grad is not None and getattr(geom, f).requires_grad]
            grads = [det[f].grad for f in geom_fields if det[f].grad is not None and getattr(geom, f).requires_grad]
            if outs:
                torch.autograd.backward(outs, grads)
            t_pre_bwd = time.perf_counter() - t0
        rng = binning.ranges
        rows = torch.arange(rng.shape[0]) // binning.grid[0]
        sel = ((rows % stride) == phase).numpy() if stride > 1 else slice(None)
        blended = float((rng[:, 1] - rng[:, 0])[sel].sum())
        return t_pre + t_pre_bwd, t_blend, blended, float(len(binning.keys))

    per_step_budget = budget_s / max(1, steps + warmup)
    # probe at a coarse stride to size the sample (this probe is the first warm-up step)
    probe_stride = max(1, tile_rows_total // 8)
    tc, tb, nb, nd = one(0, probe_stride, probe_stride // 2)
    rate = tb / max(nb, 1.0)                     # blend seconds per instance
    stride = 1
    while stride < tile_rows_total and tc + rate * nd / stride > per_step_budget:
        stride *= 2
    rec = []
    for i in range(max(0, warmup - 1) + steps):
        r = one(i, stride, i % stride)
        if i >= max(0, warmup - 1):
            rec.append(r)
    blend_s, blend_n = sum(r[1] for r in rec), sum(r[2] for r in rec)
    if blend_n <= 0:
        raise RuntimeError("the CPU sample blended no instance: lower the stride / raise --cpu-budget-s")
    rate = blend_s / blend_n
    fulls = [r[0] + rate * r[3] for r in rec] if stride > 1 else [r[0] + r[1] for r in rec]
    vps = len(fulls) / sum(fulls)
    wall = sum(r[0] + r[1] for r in rec)
    sample = (f"{steps} steps x 1 view: projection + binning (+ their backward) of all {P} Gaussians in full, blend"
              f"{'' if fwd_only else ' + backward'} of every {stride}-th tile row "
              f"({100 * blend_n / max(1.0, sum(r[3] for r in rec)):.1f}% of the instances), {wall:.1f} s CPU wall; "
              + ("stride 1 = complete views, nothing scaled" if stride == 1 else
                 "full-view time = per-Gaussian phases (measured in full) + blend time x instances-of-the-view / "
                 "instances-blended (phases timed separately, no fit)"))
    return vps, cores, sample, sum(fulls) / len(fulls) * 1e3


def reference_arm(a, rank):
    if rank != 0:
        return 0
    cfg = a.cfg
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    vps, cores, sample, ms = cpu_reference_run(a, cfg, steps, warmup, a.cpu_budget_s)
    unit = "frames/s" if a.config == "playback" else UNIT
    line = {"impl": "reference", "metric": cfg["metric"], "value": vps, "unit": unit, "n_gpus": a.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "points": cfg["points"], "resolution": a.res,
                       "views_per_step": cfg["views"], "sh_degree": cfg["sh_degree"]},
            "cpu_baseline": {"value": vps, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": vps, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's rasterizer is an external CUDA-only dependency absent from the tree; "
                    "this arm is the CPU oracle port of its published algorithm (oracle/splat_torch.py)"}
    print(json.dumps(line))
    return 0


# ---- our arm: training-shaped configurations -----------------------------------------------------------------

class TrainWorkload:
    """Everything one training-shaped configuration needs on one rank: parameter slots, camera block, loss weights,
    the eager step and its CUDA-graph captures."""

    NAMES = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")

    def __init__(self, a, cfg, name, dev, rank, world, fused, total_steps, n_slots=2):
        from gaussianip_b200 import multiview, synthetic
        from gaussianip_b200.cameras import CameraBlock, look_at_c2w, orbit_position
        import numpy as np
        self.a, self.cfg, self.name, self.dev, self.rank, self.world = a, cfg, name, dev, rank, world
        self.P, self.sh, self.res = cfg["points"], cfg["sh_degree"], a.res
        P, res = self.P, self.res
        cloud = synthetic.make_cloud(P, self.sh, 0)
        # one flat pinned host buffer (the e2e leg uploads it with ONE DMA per step) and n_slots flat device buffers
        # whose views are the leaf parameters: slot s is read by graph variant s
        shapes = {k: tuple(getattr(cloud, k).shape) for k in self.NAMES}
        offs = [0]
        for k in self.NAMES:
            offs.append(offs[-1] + (int(math.prod(shapes[k])) + 63) // 64 * 64)
        self.flat_host = torch.zeros(offs[-1], dtype=torch.float32).pin_memory()
        for k, o in zip(self.NAMES, offs):
            self.flat_host[o:o + int(math.prod(shapes[k]))].copy_(getattr(cloud, k).reshape(-1))
        self.slots = []
        for _ in range(n_slots):
            buf = self.flat_host.to(dev)
            leaves = {k: buf[o:o + int(math.prod(shapes[k]))].view(shapes[k]).requires_grad_(True)
                      for k, o in zip(self.NAMES, offs)}
            self.slots.append({"buf": buf, "leaves": leaves, "ready": torch.cuda.Event(), "free": torch.cuda.Event()})
        # cameras
        if cfg["cameras"] == "vcr":
            n_global = cfg["views"]
            self.local_ids = multiview.shard_views(n_global, rank, world)
            az = np.linspace(-180, 180, n_global + 1)[:n_global]
            fixed = [(look_at_c2w(orbit_position(float(z), 17.0, 1.5)), math.radians(70.0)) for z in az]
            self.specs = [[fixed[v] for v in self.local_ids]] * total_steps
            self.views_global = n_global
        else:
            rng = np.random.default_rng(1000)
            n_global = cfg["views"] * world
            self.specs = []
            for _ in range(total_steps):
                specs, costs = [], []
                for i in range(n_global):          # camera_data.py:349-364 distributions, identical on every rank
                    azd = (rng.random() + i) / n_global * 360.0 - 180.0
                    el, dist_, fovy = rng.uniform(-30, 30), rng.uniform(1.3, 1.7), float(np.radians(rng.uniform(40, 70)))
                    specs.append((look_at_c2w(orbit_position(azd, el, dist_)), fovy))
                    costs.append(multiview.view_cost_proxy(dist_, fovy))
                mine = (multiview.shard_views_balanced(costs, rank, world) if a.view_sharding == "balanced"
                        else multiview.shard_views(n_global, rank, world))
                self.specs.append([specs[v] for v in mine])
            self.views_global = n_global
        self.V = len(self.specs[0])
        if self.V == 0:
            raise SystemExit(f"config {name}: rank {rank} has no view (views {cfg['views']}, world {world})")
        self.cams = CameraBlock(self.V, res, res, device=dev, ring=4)
        gen = torch.Generator().manual_seed(2 + rank)
        self.w = [torch.randn(self.V, c, res, res, generator=gen).to(dev).reshape(-1) for c in (3, 1, 1)]
        self.bg = torch.zeros(3, device=dev)
        self.vp = multiview.ViewParallel(self.slots[0]["leaves"], P, fused_exchange=fused,
                                         exchange_algorithm=a.exchange_algo)
        self.captured = {}
        self.graph_error = None

    # -- model / step ---------------------------------------------------------------------------------------
    def model(self, slot):
        p = self.slots[slot]["leaves"]
        sh = self.sh

        class Model:            # GaussianModel getters (gaussian_model.py:84-107) over a dict of leaves
            active_sh_degree = sh
            _opacity = property(lambda s: p["opacity"])       # the raw parameters, reference attribute names
            _scaling = property(lambda s: p["scaling"])
            _rotation = property(lambda s: p["rotation"])
            get_xyz = property(lambda s: p["xyz"])
            get_features = property(lambda s: torch.cat((p["features_dc"], p["features_rest"]), dim=1))
            get_opacity = property(lambda s: torch.sigmoid(p["opacity"]))
            get_scaling = property(lambda s: torch.exp(p["scaling"]))
            get_rotation = property(lambda s: torch.nn.functional.normalize(p["rotation"]))
        return Model()

    def loss_of(self, views, out):
        # L = sum_views <w_c, colour> + <w_d, depth> + <w_a, alpha>  (SURVEY.md §8d): dense, non-trivial
        # dL/dcolour, dL/ddepth, dL/dalpha; written as dot products (one reduction kernel each)
        return torch.dot(out["render"].reshape(-1), self.w[0]) + torch.dot(out["depth_3dgs"].reshape(-1), self.w[1]) + \
            torch.dot(out["alpha_3dgs"].reshape(-1), self.w[2])

    def step(self, slot=0, vp=None):
        """One step on the cameras currently in the camera block, parameters of `slot`."""
        from gaussianip_b200 import renderer
        a, vp = self.a, (self.vp if vp is None else vp)
        model, cams, V = self.model(slot), self.cams.cameras, self.V
        vp.bucket.params = self.slots[slot]["leaves"]
        if a.variant == "standin":     # the reference's structure: one render() call per view from a Python loop
            def loss_one(v, out):
                n3, n1 = 3 * self.res * self.res, self.res * self.res
                return torch.dot(out["render"].reshape(-1), self.w[0][v * n3:(v + 1) * n3]) + \
                    torch.dot(out["depth_3dgs"].reshape(-1), self.w[1][v * n1:(v + 1) * n1]) + \
                    torch.dot(out["alpha_3dgs"].reshape(-1), self.w[2][v * n1:(v + 1) * n1])
            return vp.step(V, lambda v, vsp: renderer.render(cams[v], model, None, self.bg, screenspace_points=vsp),
                           loss_one, views=range(V))

        def render_views_fn(views, vsp, exchange=None):
            return renderer.render_views([cams[v] for v in views], model, None, self.bg, screenspace_points=vsp,
                                         exchange=exchange, fused_activations=not a.torch_activations)
        return vp.step_batched(V, render_views_fn, self.loss_of, views=range(V))

    def set_cameras(self, step_idx):
        spec = self.specs[step_idx % len(self.specs)]
        self.cams.update([c for c, _ in spec], [f for _, f in spec])

    # -- CUDA graph -----------------------------------------------------------------------------------------
    def capture(self, slot):
        from gaussianip_b200.graph import CapturedStep
        if slot in self.captured:
            return self.captured[slot]
        if self.vp.exchange is not None:
            self.vp.exchange.static = True         # graph-safe protocol (fixed half, zeroed in-stream)
        cs = CapturedStep(lambda: self.step(slot), device=self.dev, max_forwards=max(8, self.V))
        cs.capture()
        self.captured[slot] = cs
        return cs


def finish(world, rc=0):
    """End of a run.  N > 1: every rank meets at a barrier, flushes and leaves with os._exit — tearing the process
    group down while CUDA graphs that captured collective / symmetric-memory kernels are alive can hang."""
    if world > 1:
        import torch.distributed as dist
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(rc)
    return rc


def barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def reduce_max_ms(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_training(wl, steps, warmup, use_graph, flush_buf, clocks=None, profile=True, e2e=True):
    """The two timed legs of a training-shaped configuration.  Returns a dict of measurements (all ranks)."""
    from gaussianip_b200 import _lib, rasterizer
    a, dev, world = wl.a, wl.dev, wl.world
    res = {"graph": False}
    # ---- leg 1: inputs resident in HBM ------------------------------------------------------
    wl.set_cameras(0)
    for i in range(max(2, min(warmup, 3))):
        wl.set_cameras(i)
        wl.step(0)                                    # eager warm-up (sizes the workspaces)
    cs = None
    if use_graph:
        try:
            cs = wl.capture(0)
            res["graph"] = True
        except Exception as ex:                       # noqa: BLE001 -- report and fall back to eager enqueue
            wl.graph_error = f"{type(ex).__name__}: {str(ex)[:300]}"
            if a.graph == "on":
                raise
            cs = None
            torch.cuda.synchronize()
            if wl.vp.exchange is not None:
                wl.vp.exchange.static = False
                wl.vp.exchange.reset()

    def run(k):
        wl.set_cameras(k)
        if cs is not None:
            out = cs.replay()
            return out
        return wl.step(0)

    def settle(k_prev):
        if cs is not None and k_prev is not None:
            if not cs.validate(k_prev):
                raise RuntimeError("instance capacity overflowed inside the timed region (graph re-captured); rerun")

    tick = None
    for i in range(warmup):
        run(i)
        settle(tick)
        tick = cs.replays - 1 if cs is not None else None
    settle(tick)
    tick = None
    barrier(world)
    if profile:
        _lib.profile_enable(cs is None)               # stage events cannot be recorded inside a replayed graph
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    st0 = rasterizer.stats()
    if clocks is not None:
        clocks.start()
    barrier(world)
    t_wall0 = time.perf_counter()
    host_s = 0.0
    for k in range(steps):
        flush_buf.fill_(k & 0xFF)               # L2 flush between timed steps (outside the event pair)
        ev[k][0].record()
        t_h = time.perf_counter()
        out = run(warmup + k)
        host_s += time.perf_counter() - t_h     # host time to ENQUEUE the step
        ev[k][1].record()
        settle(tick)                            # validates the PREVIOUS step: the host stays one step ahead
        tick = cs.replays - 1 if cs is not None else None
    settle(tick)
    barrier(world)
    res["wall_s"] = time.perf_counter() - t_wall0
    res["clocks"] = clocks.stop() if clocks is not None else None
    my_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    res["host_enqueue_ms"] = host_s / steps * 1e3
    launches = _lib.launch_count() - launches0
    if cs is not None:
        # the graph replays the kernels its capture enqueued; count those (one capture = one step's launches)
        res["launches_per_step"] = cs.launches_per_capture
        res["gpu_launches"] = cs.launches_per_capture * steps
    else:
        res["launches_per_step"] = launches / steps
        res["gpu_launches"] = int(launches)
    res["prof_overlapped"] = _lib.profile_read() if (profile and cs is None) else None
    if profile:
        _lib.profile_enable(False)
    st1 = rasterizer.stats()
    res["D"] = (st1["num_rendered_sum"] - st0["num_rendered_sum"]) / max(1, st1["views"] - st0["views"])
    total_ms = reduce_max_ms(my_ms, dev, world)
    res["total_ms"] = total_ms
    res["views_total"] = wl.views_global * steps
    res["value"] = res["views_total"] / (total_ms * 1e-3)
    res["loss"] = float(out["loss"].item())

    # ---- serialised pass for per-kernel durations (roofline) --------------------------------
    if profile:
        rasterizer.set_multistream(False)
        n_serial = max(3, min(steps // 4, 8))
        for i in range(2):
            wl.set_cameras(i); wl.step(0)
        barrier(world)
        _lib.profile_enable(True)
        for i in range(n_serial):
            wl.set_cameras(i); wl.step(0)
        barrier(world)
        res["prof_serial"] = _lib.profile_read()
        _lib.profile_enable(False)
        rasterizer.set_multistream(True)

    if not e2e:
        return res
    # ---- leg 2: end to end from pinned host buffers -----------------------------------------
    # Every step's parameters come from pinned HOST memory: one flat pinned buffer, uploaded with ONE async copy per
    # step on a copy stream into one of two flat device buffers while the previous step computes.  The leaves of
    # slot s ARE views of buffer s, so nothing is copied on the device; with CUDA graphs there is one captured step
    # per slot.  Cameras: host algebra + one pinned upload per step.  Result: the loss, read back every step.
    copy_stream = torch.cuda.Stream(device=dev)
    graphs = None
    if cs is not None:
        try:
            graphs = [cs, wl.capture(1)]
        except Exception as ex:                       # noqa: BLE001
            wl.graph_error = f"{type(ex).__name__}: {str(ex)[:300]}"
            graphs = None
    for sl in wl.slots:
        sl["free"].record()
    loss_host = torch.zeros(8, dtype=torch.float32).pin_memory()

    def prefetch(slot):
        sl = wl.slots[slot]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])          # the step that last read this buffer has finished
            sl["buf"].copy_(wl.flat_host, non_blocking=True)   # H2D of every parameter of the step, one DMA
            sl["ready"].record()

    ticks = [None, None]

    def e2e_step(step_idx, last):
        cur = torch.cuda.current_stream(dev)
        slot = step_idx & 1
        sl = wl.slots[slot]
        cur.wait_event(sl["ready"])
        if not last:
            prefetch(1 - slot)                          # next step's upload overlaps this step's kernels
        wl.set_cameras(step_idx)                        # camera matrices built on host, one async upload
        if graphs is not None:
            out = graphs[slot].replay()
            ticks[slot] = graphs[slot].replays - 1
        else:
            out = wl.step(slot)
        sl["free"].record()
        loss_host[step_idx & 7:(step_idx & 7) + 1].copy_(out["loss"].reshape(1), non_blocking=True)   # D2H result
        if graphs is not None and ticks[1 - slot] is not None:
            if not graphs[1 - slot].validate(ticks[1 - slot]):      # previous step (other slot)
                raise RuntimeError("instance capacity overflowed inside the e2e region (graph re-captured); rerun")
        return out

    n_e2e_warm = min(4, max(2, warmup)) & ~1            # even, so slot parity continues into the timed steps
    prefetch(0)
    for i in range(n_e2e_warm):
        e2e_step(i, False)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_h = time.perf_counter()
    e0.record()
    for k in range(steps):
        e2e_step(n_e2e_warm + k, k == steps - 1)
    e1.record()
    res["e2e_host_enqueue_ms"] = (time.perf_counter() - t_h) / steps * 1e3
    barrier(world)
    e2e_loss = float(loss_host[(n_e2e_warm + steps - 1) & 7].item())
    assert e2e_loss == e2e_loss, "e2e loss is NaN"
    res["e2e_graph"] = graphs is not None
    res["e2e_value"] = res["views_total"] / (reduce_max_ms(e0.elapsed_time(e1), dev, world) * 1e-3)
    res["h2d"] = wl.flat_host.numel() * 4 + wl.cams.nbytes()
    res["d2h"] = 4
    return res


def exchange_check(wl):
    """N > 1: one step through the fused in-kernel exchange and one through the NCCL bucket all-reduce on the same
    views and parameters; every rank must end up with the same reduced gradients.  Returns the JSON object."""
    import torch.distributed as dist
    from gaussianip_b200 import multiview
    dev = wl.dev
    if wl.vp.exchange is None:
        return {"skipped": "NCCL exchange in use (nothing fused to check)"}
    wl.set_cameras(0)

    def flat_of(out, leaves):
        return torch.cat([out["grads"][k].reshape(-1).float() if out["grads"][k] is not None
                          else torch.zeros_like(leaves[k]).reshape(-1) for k in wl.NAMES]
                         + [out["viewspace_grad"].reshape(-1)]).clone()
    leaves = wl.slots[0]["leaves"]
    out_f = wl.step(0)
    flat_f, radii_f, loss_f = flat_of(out_f, leaves), out_f["radii"].clone(), float(out_f["loss"].item())
    vp_n = multiview.ViewParallel(leaves, wl.P, fused_exchange=False)
    out_n = wl.step(0, vp=vp_n)
    flat_n, radii_n, loss_n = flat_of(out_n, leaves), out_n["radii"].clone(), float(out_n["loss"].item())
    for p in leaves.values():
        p.grad = None
    scale = float(flat_n.abs().max().item())
    err = float((flat_f - flat_n).abs().max().item())
    lo, hi = flat_f.clone(), flat_f.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max().item())
    chk = {"rel": err / max(scale, 1e-30), "max_abs_err": err, "grad_scale": scale,
           "radii_equal": bool(torch.equal(radii_f, radii_n)), "replica_spread": spread,
           "loss_fused": loss_f, "loss_nccl": loss_n, "algorithm": wl.vp.exchange.algorithm,
           "views_per_rank": wl.V, "tolerance": 1e-5}
    flag = torch.tensor([int(chk["rel"] <= 1e-5 and chk["radii_equal"] and spread <= 1e-5 * max(scale, 1e-30))],
                        device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    chk["ok"] = bool(flag.item())
    del vp_n
    return chk


def roofline_of(res, cfg, a, world, peaks):
    P, K, HW, D = cfg["points"], (cfg["sh_degree"] + 1) ** 2, a.res * a.res, res["D"]
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_bytes = {   # SURVEY.md §8(d) per-stage algorithmic bytes
        "preprocess_fwd": P * (44 + 12 * K + 48), "depth_sort": 4 * 16 * P, "scan_emit": 8 * P + 20 * P + 12 * D,
        "tile_sort": 2 * 16 * D, "ranges": 8 * D, "render_fwd": 44 * D + 28 * HW,
        "render_bwd": 44 * D + 28 * HW + 40 * P, "preprocess_bwd": P * (40 + 44 + 12 * K + 56 + 12 * K)}
    prof = res["prof_serial"]
    stage_ms = {k: (m / c if c else 0.0) for k, (m, c) in prof.items()}
    dom = max(stage_ms, key=lambda k: stage_ms[k] * prof[k][1])
    achieved = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(dom)
        traffic_src = "static: ncu --set full capture committed under profiles/ (" + str(tj.get("_source", "traffic.json")) + ")"
    except Exception:
        pass
    b_view = P * (300 + 36 * K) + 260 * D + 56 * HW
    value = res["value"]
    r = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
         "algorithmic_bytes_per_launch": alg_bytes[dom], "avg_launch_ms": stage_ms[dom],
         "whole_view": {"algorithmic_bytes": b_view, "achieved_gbs": b_view * value / world / 1e9,
                        "frac": b_view * value / world / 1e9 / peak},
         "sort": {"depth_sort_keys_per_s": (P * 4 / (stage_ms["depth_sort"] * 1e-3)) if stage_ms["depth_sort"] > 0 else None,
                  "tile_sort_keys_per_s": (D * 2 / (stage_ms["tile_sort"] * 1e-3)) if stage_ms["tile_sort"] > 0 else None,
                  "note": "keys x 8-bit passes per second: 32-bit depth keys of the P Gaussians (4 passes) and "
                          "tile ids of the D instances (2 passes at 4096 tiles); same final order as one "
                          "64-bit (tile|depth) sort of D keys"},
         "stage_us_per_view": {k: round(v * 1e3, 1) for k, v in stage_ms.items()},
         "stage_hbm_frac": {k: round(alg_bytes[k] / (v * 1e-3) / 1e9 / peak, 3) if v > 0 else None
                            for k, v in stage_ms.items()},
         "note": "avg_launch_ms / stage_us_per_view: CUDA events recorded by the library at stage boundaries on a "
                 "serialised pass of the same workload inside bench.py (views back to back on one stream); in the "
                 "timed region the views of a step overlap on side streams (or graph branches), which is what "
                 "`value` measures"}
    if res.get("prof_overlapped"):
        r["stage_us_per_view_overlapped"] = {k: round(m / c * 1e3, 1) if c else 0.0
                                             for k, (m, c) in res["prof_overlapped"].items()}
    return r


def training_arm(a, rank, world, local_rank):
    import torch.distributed as dist
    from gaussianip_b200 import _lib, rasterizer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    cfg = a.cfg
    fused = world > 1 and a.exchange in ("auto", "fused")
    if fused:
        from gaussianip_b200.exchange import GradExchange
        if not GradExchange.available(dev):
            if a.exchange == "fused":
                raise SystemExit("--exchange fused: NVLS multicast symmetric memory is not available on this box")
            fused = False
    standin = a.variant == "standin"
    if standin:
        rasterizer.set_blend_variant("standin")
        rasterizer.set_binning_mode("flat64", dev)
    elif a.variant != "native":
        rasterizer.set_blend_variant(a.variant)
    use_graph = a.graph != "off" and not standin
    total_steps = a.warmup + a.steps + 8
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    wl = TrainWorkload(a, cfg, a.config, dev, rank, world, fused, total_steps)

    check = None
    if world > 1 and not a.no_exchange_check:
        check = exchange_check(wl)

    clocks = ClockSampler(local_rank) if rank == 0 else None
    res = measure_training(wl, a.steps, a.warmup, use_graph, flush_buf, clocks=clocks)

    # ---- config 4 in the same process (default configuration only) -----------------------------
    vcr = None
    if a.config == "ahds" and not a.no_vcr and not standin and a.points is None and a.views is None:
        try:
            del wl.captured
            wl.captured = {}
            torch.cuda.synchronize()
            vcfg = dict(CONFIGS["vcr"])
            if vcfg["views"] % world == 0:
                steps_v = max(3, min(a.steps, 10))
                wv = TrainWorkload(a, vcfg, "vcr", dev, rank, world, fused, steps_v + 6, n_slots=2)
                rv = measure_training(wv, steps_v, 2, use_graph, flush_buf, profile=False)
                vcr = {"metric": vcfg["metric"], "value": rv["value"], "unit": UNIT, "scaling": "strong",
                       "steps": steps_v, "warmup": 2, "views_per_step_total": vcfg["views"],
                       "views_per_step_per_gpu": wv.V, "ms_per_step": rv["total_ms"] / steps_v,
                       "e2e": {"value": rv["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": rv["h2d"],
                               "d2h_bytes_per_step": rv["d2h"]},
                       "graph": rv["graph"], "graph_error": wv.graph_error, "num_rendered_D": rv["D"],
                       "workload": workload_name(a, vcfg, "vcr")}
                del wv
        except Exception as ex:                       # noqa: BLE001 -- the headline line must still be printed
            vcr = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    rc = 0 if (check is None or check.get("ok", True)) else 3
    if rank != 0:
        return finish(world, rc)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = roofline_of(res, cfg, a, world, peaks)
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        vps, cores, sample, _ = cpu_reference_run(a, cfg, 1, 0, 40.0)
        cpu = {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {"variant": "reference-STRUCTURE stand-in (csrc/standin.cu + flat 64-bit sort + per-view loop); NOT the "
                       "reference and NOT the product path"} if standin else (
        {"variant": f"backward blend variant {a.variant} (gsb_set_blend_variant)"} if a.variant != "native" else {})
    if fused:
        par = (f"view-sharded dp{world}, gradients reduced INSIDE the backward kernel ({wl.vp.exchange.algorithm} over "
               f"NVLink peer / NVLS multicast memory, {wl.vp.exchange.nbytes() >> 20} MiB) + NCCL radii max per step")
    elif world > 1:
        par = (f"view-sharded dp{world}, NCCL all-reduce of the flat gradient bucket ({wl.vp.bucket.nbytes() >> 20} MiB) "
               f"+ radii max per step")
    else:
        par = "single GPU"
    line.update({
        "metric": cfg["metric"], "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": res["total_ms"] / a.steps, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "name": a.config, "points": cfg["points"], "resolution": a.res,
                   "views_per_step_per_gpu": wl.V, "views_per_step_total": wl.views_global,
                   "sh_degree": cfg["sh_degree"], "num_rendered_D": res["D"],
                   "l2": "256 MiB flush write between timed steps (outside the event pairs)",
                   "activations": "torch getters around the operator" if a.torch_activations else
                                  "fused into the per-Gaussian kernels (render_views(fused_activations=True))",
                   "launch": ("one CUDA-graph replay per step (gaussianip_b200.graph.CapturedStep; cameras and "
                              "intrinsics read from a device block refreshed per step)" if res["graph"] else
                              "eager: every kernel enqueued through ctypes / torch per step"),
                   "graph_error": wl.graph_error,
                   "view_sharding": "n/a" if world == 1 else ("views r, r+N, ..." if a.config == "vcr" else a.view_sharding),
                   "parallelism": par, "wall_s_timed_region": res["wall_s"],
                   "host_enqueue_ms_per_step_rank0": res["host_enqueue_ms"],
                   "e2e_host_ms_per_step_rank0": res.get("e2e_host_enqueue_ms"), "e2e_graph": res.get("e2e_graph"),
                   "host_cores": os.cpu_count(),
                   "parity_notes": "gradients within 1e-4 (per element, see tests/test_gpu_parity.py); n_contrib exact "
                                   "except pixels the oracle flags marginal (ex2.approx vs CPU exp); fused activations "
                                   "may move a radius by one pixel for < 1e-4 of the points"},
        "clocks": res["clocks"],
        "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"]},
        "gpu_launches": int(res["gpu_launches"]), "gpu_launches_per_step": res["launches_per_step"],
        "roofline": roofline, "cpu_baseline": cpu})
    if check is not None:
        line["exchange_check"] = check
    if vcr is not None:
        line["vcr"] = vcr
    print(json.dumps(line))
    return finish(world, rc)


# ---- our arm: playback (forward only) ---------------------------------------------------------------------------

def playback_arm(a, rank, world, local_rank):
    """animation.py:463-467 per frame: new Gaussian positions (the reference poses them on the host and uploads
    them), gs.render(MiniCam) -> image.permute(1, 2, 0).contiguous().detach().cpu().numpy().  Frames are independent,
    so N GPUs play N disjoint frame ranges (replicas, no collective)."""
    import torch.distributed as dist
    from gaussianip_b200 import _lib, renderer, synthetic
    from gaussianip_b200.graph import CapturedStep
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    cfg = a.cfg
    P, sh, res = cfg["points"], cfg["sh_degree"], a.res
    n_frames, n_pose = 136, 34
    cloud = synthetic.make_cloud(P, sh, 0)
    poses_host = [synthetic.playback_sway(cloud.xyz, (i * 4 + rank) % n_frames, n_frames).pin_memory()
                  for i in range(n_pose)]
    poses_dev = [p.to(dev) for p in poses_host]
    cl = cloud.to(dev)
    cams_all = synthetic.playback_cameras(n_frames, res, res, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    class Model:
        active_sh_degree = sh
        max_sh_degree = sh

        def __init__(self, xyz):
            self.xyz = xyz
        get_xyz = property(lambda s: s.xyz)
        get_features = property(lambda s: cl.get_features())
        get_opacity = property(lambda s: cl.get_opacity())
        get_scaling = property(lambda s: cl.get_scaling())
        get_rotation = property(lambda s: cl.get_rotation())

    slots = []
    for s in range(2):
        xyz = poses_dev[0].clone()
        cam_rows = torch.zeros(4, 16, device=dev)     # view, proj(unused), full, centre of the frame's MiniCam
        slots.append({"xyz": xyz, "r": renderer.Renderer(sh, False, gaussians=Model(xyz), device=dev),
                      "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                      "img_host": torch.empty(res, res, 3, dtype=torch.float32).pin_memory(),
                      "img_done": torch.cuda.Event()})
    cam0 = cams_all[0]

    class StaticCam:                                  # MiniCam whose matrices live at fixed addresses
        def __init__(self):
            self.image_width, self.image_height, self.FoVx, self.FoVy = cam0.image_width, cam0.image_height, cam0.FoVx, cam0.FoVy
            self.znear, self.zfar = cam0.znear, cam0.zfar
            self.world_view_transform = cam0.world_view_transform.contiguous().clone()
            self.projection_matrix = cam0.projection_matrix.contiguous().clone()
            self.full_proj_transform = cam0.full_proj_transform.contiguous().clone()
            self.camera_center = cam0.camera_center.contiguous().clone()
    scam = StaticCam()
    cam_stage = torch.stack([torch.cat((c.world_view_transform.contiguous().reshape(-1),
                                        c.full_proj_transform.contiguous().reshape(-1),
                                        c.camera_center.reshape(-1))) for c in cams_all]).contiguous()   # [136, 35] device

    def set_cam(f):
        row = cam_stage[f % n_frames]
        scam.world_view_transform.copy_(row[0:16].view(4, 4))
        scam.full_proj_transform.copy_(row[16:32].view(4, 4))
        scam.camera_center.copy_(row[32:35])

    def frame(slot):
        with torch.no_grad():
            out = slots[slot]["r"].render(scam)
            return out["image"].permute(1, 2, 0).contiguous()     # [H, W, 3] in [0, 1], as render_gs builds it

    use_graph = a.graph != "off"
    graphs, graph_error = None, None
    for s in range(2):
        frame(s)
    torch.cuda.synchronize()
    if use_graph:
        try:
            graphs = [CapturedStep(lambda s=s: frame(s), device=dev, max_forwards=4).capture() for s in range(2)]
        except Exception as ex:                       # noqa: BLE001
            graph_error = f"{type(ex).__name__}: {str(ex)[:300]}"
            graphs = None
            if a.graph == "on":
                raise

    def render_slot(slot):
        if graphs is not None:
            img = graphs[slot].replay()
            return img
        return frame(slot)

    # ---- leg 1: poses resident in HBM, no image read-back --------------------------------------
    def value_frame(f):
        slots[0]["xyz"].copy_(poses_dev[f % n_pose])
        set_cam(f)
        return render_slot(0)
    for f in range(a.warmup):
        value_frame(f)
    if graphs is not None:
        graphs[0].validate()
    barrier(world)
    launches0 = _lib.launch_count()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier(world)
    for k in range(a.steps):
        flush_buf.fill_(k & 0xFF)
        ev[k][0].record()
        value_frame(a.warmup + k)
        ev[k][1].record()
    if graphs is not None and not graphs[0].validate():
        raise RuntimeError("instance capacity overflowed inside the timed region; rerun")
    barrier(world)
    clock_info = clocks.stop() if clocks is not None else None
    launches = _lib.launch_count() - launches0
    if graphs is not None:
        launches = graphs[0].launches_per_capture * a.steps
    total_ms = reduce_max_ms(sum(e0.elapsed_time(e1) for e0, e1 in ev), dev, world)
    value = a.steps * world / (total_ms * 1e-3)

    # per-stage device times of the same frames (eager enqueue, events recorded by the library at stage boundaries)
    from gaussianip_b200 import rasterizer
    st0 = rasterizer.stats()
    for f in range(2):
        slots[0]["xyz"].copy_(poses_dev[f % n_pose]); set_cam(f); frame(0)
    barrier(world)
    _lib.profile_enable(True)
    n_prof = 12
    for f in range(n_prof):
        slots[0]["xyz"].copy_(poses_dev[f % n_pose]); set_cam(f); frame(0)
    barrier(world)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    st1 = rasterizer.stats()
    D_mean = (st1["num_rendered_sum"] - st0["num_rendered_sum"]) / max(1, st1["views"] - st0["views"])

    # ---- leg 2: positions from pinned host memory, image to the host, every frame -----------------
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    for sl in slots:
        sl["free"].record(); sl["img_done"].record()

    def prefetch(f, slot):
        sl = slots[slot]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])
            sl["xyz"].copy_(poses_host[f % n_pose], non_blocking=True)      # H2D 12 B x P
            sl["ready"].record()

    def e2e_frame(f, last):
        cur = torch.cuda.current_stream(dev)
        slot = f & 1
        sl = slots[slot]
        cur.wait_event(sl["ready"])
        if not last:
            prefetch(f + 1, 1 - slot)
        set_cam(f)
        sl["img_done"].synchronize()                     # the host has consumed this slot's previous image
        img = render_slot(slot)
        sl["free"].record()
        rendered = torch.cuda.Event()
        rendered.record()
        img.record_stream(d2h_stream)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(rendered)
            sl["img_host"].copy_(img, non_blocking=True)                    # D2H 12 B x H x W
            sl["img_done"].record()
        return sl

    prefetch(0, 0)
    for f in range(4):
        e2e_frame(f, False)
    for sl in slots:
        sl["img_done"].synchronize()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        e2e_frame(4 + k, k == a.steps - 1)
    for sl in slots:
        sl["img_done"].synchronize()
    e1.record()
    barrier(world)
    if graphs is not None:
        for g_ in graphs:
            if not g_.validate():
                raise RuntimeError("instance capacity overflowed inside the e2e region; rerun")
    e2e_value = a.steps * world / (reduce_max_ms(e0.elapsed_time(e1), dev, world) * 1e-3)
    mean_img = float(slots[(4 + a.steps - 1) & 1]["img_host"].mean())
    assert 0.0 < mean_img < 1.0, "playback image is empty"
    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            vps, cores, sample, _ = cpu_reference_run(a, cfg, 1, 0, 40.0)
            cpu = {"value": vps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        K_ = (sh + 1) ** 2
        alg = {"preprocess_fwd": P * (44 + 12 * K_ + 48), "depth_sort": 4 * 16 * P, "scan_emit": 28 * P + 12 * D_mean,
               "tile_sort": 2 * 16 * D_mean, "ranges": 8 * D_mean, "render_fwd": 44 * D_mean + 28 * res * res}
        stage_ms = {k: (m / c if c else 0.0) for k, (m, c) in prof.items() if k in alg}
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        ach = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        b_frame = sum(alg.values())
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": stage_ms[dom],
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else
                                   "fallback (B200_PROFILING.md 6.65 TB/s)",
                    "whole_frame": {"algorithmic_bytes": b_frame, "achieved_gbs": b_frame * value / world / 1e9,
                                    "frac": b_frame * value / world / 1e9 / peak},
                    "stage_us_per_frame": {k: round(v * 1e3, 1) for k, v in stage_ms.items()},
                    "num_rendered_D": D_mean,
                    "note": "stage times: CUDA events recorded by the library at stage boundaries on an eager pass "
                            "of the same frames"}
        line = {"metric": cfg["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "name": "playback", "points": P, "resolution": res,
                           "sh_degree": sh, "frames_in_motion": n_frames, "distinct_poses": n_pose,
                           "l2": "256 MiB flush write between timed frames (outside the event pairs)",
                           "launch": "one CUDA-graph replay per frame" if graphs is not None else "eager",
                           "graph_error": graph_error,
                           "parallelism": "single GPU" if world == 1 else f"replicas: {world} GPUs play disjoint frames, no collective"},
                "clocks": clock_info,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": P * 12,
                        "d2h_bytes_per_step": res * res * 12},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    return finish(world, 0)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return reference_arm(a, rank)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gaussianip_b200 has no CPU fallback")
    if a.config == "playback":
        return playback_arm(a, rank, world, local_rank)
    return training_arm(a, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
