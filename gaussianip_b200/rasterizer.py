"""Drop-in for the ``diff_gaussian_rasterization`` Python surface GaussianIP uses.

Same names, argument meaning, return tuple and error behaviour as the operator the
reference constructs at gaussiansplatting/gaussian_renderer/__init__.py:36-51 and calls at
:85-93 (also :124-139/175-183, :213-228/240-248 and gs_renderer.py:943-958/992-1001):

    settings   = GaussianRasterizationSettings(image_height=..., ..., debug=False)
    rasterizer = GaussianRasterizer(raster_settings=settings)
    color, radii, depth, alpha = rasterizer(means3D=..., means2D=..., shs=..., colors_precomp=...,
                                            opacities=..., scales=..., rotations=..., cov3D_precomp=...)

Host code is Python/PyTorch (device memory, streams, autograd plumbing); all arithmetic
runs in the hand-written sm_100a kernels of libgsb.so through the C ABI in include/gsb.h.
There is no CPU path: a missing library or a non-CUDA tensor raises.

Differences from the external operator, by design (B200-first):
* No device->host synchronisation to learn ``num_rendered``: instance buffers are sized by a
  capacity that follows the scene, D is read back asynchronously right after the scan, and
  the host only waits on that tiny copy (normally long finished).  If D exceeded the
  capacity the forward is re-enqueued with a larger workspace before anything else can
  observe the outputs.
* The library never allocates: this module owns a per-(device, stream) scratch block and a
  per-call ``saved`` block kept for backward.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ---- workspace --------------------------------------------------------------------------------

class _Workspace:
    """Transient scratch shared by all calls on one (device, stream)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.scratch: Optional[torch.Tensor] = None
        self.d_cap = 0
        self.host_counts = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(device))   # materialise the CUDA handle
        self.binning_mode = _lib.BIN_TWO_LEVEL
        self.last_num_rendered = 0
        self.retries = 0
        self.pending = None            # speculative forward whose counts were not read yet

    def capacity_for(self, P: int) -> int:
        if self.d_cap == 0:
            self.d_cap = max(1 << 16, 4 * P)
        return self.d_cap

    def ensure_scratch(self, nbytes: int) -> torch.Tensor:
        if self.scratch is None or self.scratch.numel() < nbytes:
            self.scratch = None
            self.scratch = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=self.device)
        return self.scratch


_workspaces = {}                # (device index, stream handle) -> _Workspace, most recently used last
MAX_WORKSPACES = 64             # streams come and go; a dead stream's scratch block must not live forever


def _workspace(device: torch.device) -> _Workspace:
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.pop(key, None)
    if ws is None:
        ws = _Workspace(device)
        while len(_workspaces) >= MAX_WORKSPACES:           # drop the least recently used (dicts keep insertion order)
            old = _workspaces.pop(next(iter(_workspaces)))
            if old.pending is not None:
                old.pending.settle()
    _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    """Free every per-stream scratch block (they are re-created on demand)."""
    for ws in list(_workspaces.values()):
        if ws.pending is not None:
            ws.pending.settle()
    _workspaces.clear()


def set_binning_mode(mode: str, device=None) -> None:
    """'two_level' (default) or 'flat64' (reference-structure 64-bit key sort); identical results."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _workspace(dev).binning_mode = {"two_level": _lib.BIN_TWO_LEVEL, "flat64": _lib.BIN_FLAT64}[mode]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check(t: torch.Tensor, name: str, shape_tail, device) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device != device:
        raise ValueError(f"{name} must be on {device}, got {t.device}")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != len(shape_tail) + 1 or tuple(t.shape[1:]) != tuple(shape_tail):
        raise ValueError(f"{name} must have shape [P, {', '.join(map(str, shape_tail))}], got {tuple(t.shape)}")
    return t.contiguous()


def _small(t: torch.Tensor, n: int, name: str, device) -> torch.Tensor:
    if not torch.is_tensor(t) or t.numel() != n:
        raise ValueError(f"{name} must be a tensor with {n} elements")
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


def _make_settings(rs: GaussianRasterizationSettings, device, raw: int = 0, tanfov_dev: Optional[torch.Tensor] = None,
                   forward_only: bool = False):
    bg = _small(rs.bg, 3, "bg", device)
    view = _small(rs.viewmatrix, 16, "viewmatrix", device)
    proj = _small(rs.projmatrix, 16, "projmatrix", device)
    campos = _small(rs.campos, 3, "campos", device)
    if not (0 <= int(rs.sh_degree) <= 3):
        raise ValueError("sh_degree must be in 0..3")
    s = _lib.GsbSettings(int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
                         float(rs.scale_modifier), int(rs.sh_degree), int(bool(rs.prefiltered)),
                         int(bool(rs.debug)), int(raw), int(bool(forward_only)), bg.data_ptr(), view.data_ptr(), proj.data_ptr(),
                         campos.data_ptr(), None if tanfov_dev is None else _small(tanfov_dev, 2, "tanfov_dev", device).data_ptr())
    return s, (bg, view, proj, campos, tanfov_dev)      # keep the tensors alive next to the struct


class _Saved:
    """Per-call state kept for backward (and exposed to the parity tests)."""
    __slots__ = ("block", "layout", "d_cap", "P", "K", "H", "W", "num_rendered", "scratch", "settings_keep",
                 "forward_only")

    def view(self, off: int, nbytes: int, dtype) -> torch.Tensor:
        return self.block[off:off + nbytes].view(dtype)

    # named views -------------------------------------------------------------------------
    def geom(self):
        return self.view(self.layout.off_geom, self.P * 48, torch.float32).view(self.P, 12)

    def point_list(self):
        return self.view(self.layout.off_point_list, self.num_rendered * 4, torch.int32)

    def ranges(self):
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        return self.view(self.layout.off_ranges, T * 8, torch.int32).view(T, 2)

    def n_contrib(self):
        return self.view(self.layout.off_n_contrib, self.H * self.W * 4, torch.int32).view(self.H, self.W)

    def final_T(self):
        return self.view(self.layout.off_final_T, self.H * self.W * 4, torch.float32).view(self.H, self.W)

    def sorted_keys(self) -> torch.Tensor:
        """Materialise the sorted 64-bit (tile<<32 | depth bits) keys (valid until the next call
        on the same stream reuses the scratch block)."""
        out = torch.empty(max(self.num_rendered, 1), dtype=torch.int64, device=self.block.device)
        st = torch.cuda.current_stream(self.block.device).cuda_stream
        _lib.check(_lib.load().gsb_debug_sorted_keys(self.P, self.H, self.W, self.block.data_ptr(),
                                                     self.scratch.data_ptr(), self.d_cap, out.data_ptr(), st),
                   "gsb_debug_sorted_keys")
        return out[:self.num_rendered]


class _ForwardCall:
    """One view's forward: enqueue() queues every kernel on the CURRENT stream; finish() waits for
    the 32-byte counts copy and, if D exceeded the instance capacity, re-enqueues with a larger
    workspace (nothing can have observed the outputs yet)."""

    def __init__(self, rs, means3D, shs, colors, opacities, scales, rotations, cov3D, out=None, raw: int = 0,
                 tanfov_dev=None, forward_only: bool = False):
        self.raw = raw
        self.tanfov_dev = tanfov_dev
        self.forward_only = bool(forward_only)    # no backward will follow: no hit records, shorter saved block
        self.args = (means3D, shs, colors, opacities, scales, rotations, cov3D)
        device = means3D.device
        if device.type != "cuda":
            raise ValueError("gaussianip_b200 runs on CUDA tensors only (no CPU fallback)")
        self.device = device
        self.P = means3D.shape[0]
        self.H, self.W = int(rs.image_height), int(rs.image_width)
        self.K = shs.shape[1] if shs is not None else 0
        self.rs = rs
        self.out = out
        self.stream = None
        self.counts = None
        self.mark_binned = False      # record self.binned between binning and blend (staggered multi-view forward)
        self.binned = None

    def enqueue(self):
        lib = _lib.load()
        means3D, shs, colors, opacities, scales, rotations, cov3D = self.args
        device, P, H, W, K = self.device, self.P, self.H, self.W, self.K
        with torch.cuda.device(device):
            if self.stream is None:
                self.stream = torch.cuda.current_stream(device)
                self.ws = _workspace(device)
                self.s, self.keep = _make_settings(self.rs, device, self.raw, self.tanfov_dev, self.forward_only)
                if self.out is not None:
                    self.color, self.radii, self.depth, self.alpha = self.out   # caller-owned contiguous slices
                else:
                    self.color = torch.empty(3, H, W, dtype=torch.float32, device=device)
                    self.depth = torch.empty(1, H, W, dtype=torch.float32, device=device)
                    self.alpha = torch.empty(1, H, W, dtype=torch.float32, device=device)
                    self.radii = torch.empty(P, dtype=torch.int32, device=device)
            ws = self.ws
            spec = _active_speculation()
            if spec is not None and spec.capture:
                # CUDA-graph capture (gaussianip_b200.graph): the counts copy becomes a graph node writing this
                # call's OWN pinned block on every replay; no event (a captured event cannot be waited on)
                if self.counts is None:
                    self.counts = spec.take_counts()
                counts_ptr, event = self.counts.data_ptr(), None
            else:
                if ws.pending is not None and ws.pending is not self:
                    ws.pending.settle()         # its counts buffer is about to be reused
                self.counts = None
                counts_ptr, event = ws.host_counts.data_ptr(), ws.event.cuda_event
            self.d_cap = ws.capacity_for(P)
            self.L = _lib.layout(P, H, W, self.d_cap)
            self.scratch = ws.ensure_scratch(self.L.scratch_bytes)
            self.block = torch.empty(self.L.saved_bytes_forward_only if self.forward_only else self.L.saved_bytes,
                                     dtype=torch.uint8, device=device)
            if self.mark_binned:
                # the three stages one by one, with an event between binning and blend: the NEXT view's pipeline
                # is started behind it (rasterize_views, staggered forward)
                st = self.stream.cuda_stream
                rc = lib.gsb_preprocess_fwd(C.byref(self.s), P, K, _ptr(means3D), _ptr(scales), _ptr(rotations),
                                            _ptr(opacities), _ptr(shs), _ptr(colors), _ptr(cov3D),
                                            self.radii.data_ptr(), self.block.data_ptr(), self.scratch.data_ptr(),
                                            self.d_cap, st)
                _lib.check(rc, "gsb_preprocess_fwd")
                rc = lib.gsb_bin_sort(C.byref(self.s), P, self.block.data_ptr(), self.scratch.data_ptr(), self.d_cap,
                                      ws.binning_mode, counts_ptr, event, st)
                _lib.check(rc, "gsb_bin_sort")
                self.binned = torch.cuda.Event()
                self.binned.record(self.stream)
                rc = lib.gsb_render_fwd(C.byref(self.s), P, self.block.data_ptr(), self.d_cap, self.color.data_ptr(),
                                        self.depth.data_ptr(), self.alpha.data_ptr(), st)
                _lib.check(rc, "gsb_render_fwd")
                return self
            rc = lib.gsb_forward(C.byref(self.s), P, K, _ptr(means3D), _ptr(scales), _ptr(rotations),
                                 _ptr(opacities), _ptr(shs), _ptr(colors), _ptr(cov3D), self.radii.data_ptr(),
                                 self.color.data_ptr(), self.depth.data_ptr(), self.alpha.data_ptr(),
                                 self.block.data_ptr(), self.scratch.data_ptr(), self.d_cap, ws.binning_mode,
                                 counts_ptr, event, self.stream.cuda_stream)
            _lib.check(rc, "gsb_forward")
        return self

    PREFILTER_MSG = "Point is filtered although prefiltered is set. This shouldn't happen!"

    def _account(self, D: int) -> None:
        ws, P = self.ws, self.P
        ws.last_num_rendered = D
        _stats["num_rendered"] = D
        _stats["views"] += 1
        _stats["num_rendered_sum"] += D
        # follow the scene downwards slowly so one huge view does not pin memory forever
        if D * 4 < ws.d_cap and ws.d_cap > max(1 << 16, 4 * P):
            ws.d_cap = max(1 << 16, 4 * P, 2 * D)

    def _saved(self, D) -> "_Saved":
        sv = _Saved()
        sv.block, sv.layout, sv.d_cap, sv.P, sv.K, sv.H, sv.W = self.block, self.L, self.d_cap, self.P, self.K, self.H, self.W
        sv.num_rendered, sv.scratch, sv.settings_keep = D, self.scratch, self.keep
        sv.forward_only = self.forward_only
        return sv

    def finish(self):
        ws, P = self.ws, self.P
        spec = _active_speculation()
        if spec is not None:
            # whole-step speculation (see `speculation`): do not wait for the counts now; the
            # owner of the step validates them after everything is enqueued and redoes the step
            # if this view did not fit
            self.sv = self._saved(None)
            if not spec.capture:
                ws.pending = self
            spec.calls.append(self)
            return self.color, self.radii, self.depth, self.alpha, self.sv
        while True:
            ws.event.synchronize()          # waits for the 32-byte counts copy only
            D = int(ws.host_counts[0].item()) & 0xFFFFFFFF if P > 0 else 0
            if P > 0 and int(ws.host_counts[5].item()):
                raise RuntimeError(self.PREFILTER_MSG)      # the external operator traps the device here
            if D <= self.d_cap:
                break
            ws.d_cap = int(D * 1.25) + 4096   # outputs were not observable yet: enqueue again, larger
            ws.retries += 1
            with torch.cuda.stream(self.stream):
                self.enqueue()
        self._account(D)
        return self.color, self.radii, self.depth, self.alpha, self._saved(D)

    def settle(self) -> bool:
        """Speculative call: read the counts (they were produced long ago on the device timeline).
        Returns False if the view overflowed its instance capacity; the capacity is then raised
        for the next attempt."""
        ws, P = self.ws, self.P
        if self.counts is not None:
            # captured call: the caller has waited for the replay (graph.CapturedStep.validate)
            D = int(self.counts[0].item()) & 0xFFFFFFFF if P > 0 else 0
            if P > 0 and int(self.counts[5].item()):
                raise RuntimeError(self.PREFILTER_MSG)
            self.fits = D <= self.d_cap
            if self.fits:
                self.sv.num_rendered = D
                _stats["num_rendered"] = D
                _stats["views"] += 1
                _stats["num_rendered_sum"] += D
            else:
                ws.d_cap = max(ws.d_cap, int(D * 1.25) + 4096)
                ws.retries += 1
            return self.fits
        if ws.pending is not self:
            return self.fits
        ws.event.synchronize()
        D = int(ws.host_counts[0].item()) & 0xFFFFFFFF if P > 0 else 0
        ws.pending = None
        if P > 0 and int(ws.host_counts[5].item()):
            raise RuntimeError(self.PREFILTER_MSG)
        self.fits = D <= self.d_cap
        if self.fits:
            self.sv.num_rendered = D
            self._account(D)
        else:
            ws.d_cap = int(D * 1.25) + 4096
            ws.retries += 1
        return self.fits


class speculation:
    """Context for callers that own a WHOLE step (forward of all views, loss, backward) and can
    redo it: inside, forwards do not block the host on the 32-byte instance count; afterwards
    ``validate()`` reads the counts and reports whether every view fitted its capacity.  On False
    the capacities have been raised and the caller must discard the step's results and run it
    again (gaussianip_b200.multiview.ViewParallel does).  Removes the one host stall per step
    that otherwise lets the GPU run dry between the forward and the backward."""

    def __init__(self, capture: bool = False, counts_pool=None):
        """capture=True (used by gaussianip_b200.graph.CapturedStep while a CUDA graph is being captured): every
        forward writes its instance counts to its own block of ``counts_pool`` (pinned int32 [n, 8], allocated
        BEFORE the capture) and records no event; ``validate(keep=True)`` may then be called after every replay."""
        self.calls = []
        self.capture = bool(capture)
        self.counts_pool = counts_pool
        self.counts_used = 0

    def take_counts(self) -> torch.Tensor:
        if self.counts_pool is None or self.counts_used >= self.counts_pool.shape[0]:
            raise RuntimeError("capture needs a pinned counts pool with one row per forward call")
        row = self.counts_pool[self.counts_used]
        self.counts_used += 1
        return row

    def __enter__(self):
        if getattr(_tls, "spec", None) is not None:
            raise RuntimeError("speculation contexts do not nest")
        _tls.spec = self
        return self

    def __exit__(self, *exc):
        _tls.spec = None
        return False

    def validate(self, keep: bool = False) -> bool:
        ok = True
        for call in self.calls:
            ok = call.settle() and ok
        if not keep:
            self.calls = []
        return ok


_tls = threading.local()


def _active_speculation():
    return getattr(_tls, "spec", None)


def _forward_impl(rs, means3D, shs, colors, opacities, scales, rotations, cov3D, out=None, raw: int = 0,
                  tanfov_dev=None, forward_only: bool = False):
    return _ForwardCall(rs, means3D, shs, colors, opacities, scales, rotations, cov3D, out, raw,
                        tanfov_dev, forward_only).enqueue().finish()


# Side streams for the batched multi-view entry: the binning stages of a 1-2 M instance view are
# latency-bound (one wave of blocks per radix pass), so the views of a step run on separate
# streams and one view's sort overlaps another view's blend.
_side_streams = {}
_multistream = True
_stats = {"num_rendered": 0, "views": 0, "num_rendered_sum": 0}


def stats() -> dict:
    """Host-side counters: D (num_rendered) of the last view, number of views, sum of D."""
    return dict(_stats)


BLEND_VARIANTS = {"native": 0, "standin": 1, "replay_bwd": 2, "rescan_bwd": 3, "rescan_packed_bwd": 4,
                  "fwd_per_hit": 10, "fwd_transposed": 11, "fwd_gather4": 12, "fwd_precull": 13}


def set_blend_variant(name: str) -> None:
    """'native' (default, the product kernels: the forward blend records per-warp hit lists, the backward replays
    them in two transposed phases, lane = pixel then lane = record), 'standin' (reference-STRUCTURE blend kernels of
    csrc/standin.cu, for measurement context and GPU cross-checks only), 'replay_bwd' (replays the records one hit per
    half-warp at a time with a shuffle butterfly), 'rescan_bwd' (the record-free round-1 backward that re-walks the
    tile lists) or 'rescan_packed_bwd'; the last three are cross-checks of the default.  'fwd_per_hit' (default) /
    'fwd_transposed' / 'fwd_gather4' (TMA row gathers) / 'fwd_precull' (CTA-wide pre-cull) select the FORWARD blend
    kernel only: measured alternatives that produce identical images, contributor counts and hit records."""
    _lib.check(_lib.load().gsb_set_blend_variant(BLEND_VARIANTS[name]), "gsb_set_blend_variant")


_pbwd_group = int(__import__("os").environ.get("GSB_PBWD_GROUP", "8"))


_fwd_stagger = int(__import__("os").environ.get("GSB_FWD_STAGGER", "0"))


def set_forward_stagger(k: int) -> None:
    """rasterize_views forward on side streams: 0 = all views start together; k > 0 = view v starts its
    per-Gaussian + binning stages when view v - k has finished binning (so they overlap that view's blend)."""
    global _fwd_stagger
    _fwd_stagger = max(0, int(k))


def set_backward_grouping(views_per_launch: int) -> None:
    """How many views one launch of the fused per-Gaussian backward covers in rasterize_views (single GPU): 8 = one
    launch per step (every gradient written once), smaller = more launches that overlap the blend backwards of the
    remaining views (the later ones accumulate)."""
    global _pbwd_group
    _pbwd_group = max(1, int(views_per_launch))


def set_multistream(enabled: bool) -> None:
    """Run the views of rasterize_views on separate CUDA streams (default) or back to back."""
    global _multistream
    _multistream = bool(enabled)


MAX_SIDE_STREAMS = 8


def _streams_for(device: torch.device, n: int):
    """One side stream per view, at most MAX_SIDE_STREAMS: with more views (the 64 VCR views of one step) view v
    shares stream v mod 8 — and with it that stream's scratch workspace — with views v + 8, v + 16, ..."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    pool = _side_streams.setdefault(key, [])
    while len(pool) < min(n, MAX_SIDE_STREAMS):
        pool.append(torch.cuda.Stream(device=device))
    return [pool[i % MAX_SIDE_STREAMS] for i in range(n)]


def _backward_impl(rs, sv: _Saved, means3D, shs, colors, opacities, scales, rotations, cov3D, radii,
                   g_color, g_depth, g_alpha, out=None, accumulate=False, raw: int = 0, tanfov_dev=None):
    lib = _lib.load()
    device = means3D.device
    P, K, H, W = sv.P, sv.K, sv.H, sv.W
    with torch.cuda.device(device):
        ws = _workspace(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        s, keep = _make_settings(rs, device, raw, tanfov_dev)
        scratch = ws.ensure_scratch(sv.layout.scratch_bytes)

        def grad_in(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=device)
            return g.to(dtype=torch.float32).contiguous()

        g_color = grad_in(g_color, (3, H, W))
        g_depth = grad_in(g_depth, (1, H, W))
        g_alpha = grad_in(g_alpha, (1, H, W))
        if out is None:
            e = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=device)
            out = {"means3D": e(P, 3), "means2D": e(P, 3), "opacities": e(P, 1),
                   "shs": e(P, K, 3) if shs is not None else None,
                   "colors": e(P, 3) if colors is not None else None,
                   "scales": e(P, 3) if scales is not None else None,
                   "rotations": e(P, 4) if rotations is not None else None,
                   "cov3D": e(P, 6) if cov3D is not None else None}
        rc = lib.gsb_backward(C.byref(s), P, K, _ptr(means3D), _ptr(scales), _ptr(rotations), _ptr(opacities),
                              _ptr(shs), _ptr(colors), _ptr(cov3D), radii.data_ptr(), sv.block.data_ptr(),
                              scratch.data_ptr(), sv.d_cap, g_color.data_ptr(), g_depth.data_ptr(),
                              g_alpha.data_ptr(), _ptr(out["means3D"]), _ptr(out["means2D"]), _ptr(out["shs"]),
                              _ptr(out["colors"]), _ptr(out["opacities"]), _ptr(out["scales"]),
                              _ptr(out["rotations"]), _ptr(out["cov3D"]), int(bool(accumulate)), stream)
        _lib.check(rc, "gsb_backward")
        del keep
    return out


def _backward_views(settings_list, svs, means3D, shs, colors, opacities, scales, rotations, cov3D, radii,
                    g_color, g_depth, g_alpha, exchange=None, raw: int = 0, tanfov_dev=None):
    """Backward of V views: the blend backward of view v runs on side stream v (they overlap);
    the per-Gaussian stage runs view after view on the calling stream because it accumulates
    (beta = 1) into one set of gradient tensors."""
    lib = _lib.load()
    device = means3D.device
    V = len(settings_list)
    P, K, H, W = svs[0].P, svs[0].K, svs[0].H, svs[0].W
    with torch.cuda.device(device):
        main = torch.cuda.current_stream(device)
        e = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=device)
        if exchange is not None:
            # fused exchange (gaussianip_b200/exchange.py): the per-Gaussian kernel adds into NVLS multicast
            # addresses; `out` are this rank's views of the symmetric buffer, `optr` what the kernel is given
            exchange.ensure(P, K, {"means3D": True, "means2D": True, "opacities": True, "shs": shs is not None,
                                   "colors": colors is not None, "scales": scales is not None,
                                   "rotations": rotations is not None, "cov3D": cov3D is not None,
                                   "radii": True, "scalars": True}, device)
            exchange.begin()                 # zero + barrier, overlaps the blend backward on the side streams
            out = {k: exchange.local(k) for k in ("means3D", "means2D", "opacities", "shs", "colors", "scales",
                                                  "rotations", "cov3D")}     # ("radii" / "scalars": exchange.local)
            optr = {k: exchange.output_ptr(k) for k in out}
        else:
            out = {"means3D": e(P, 3), "means2D": e(P, 3), "opacities": e(P, 1),
                   "shs": e(P, K, 3) if shs is not None else None, "colors": e(P, 3) if colors is not None else None,
                   "scales": e(P, 3) if scales is not None else None,
                   "rotations": e(P, 4) if rotations is not None else None,
                   "cov3D": e(P, 6) if cov3D is not None else None}
            optr = {k: _ptr(t) for k, t in out.items()}

        def grad_in(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=device)
            return g.to(dtype=torch.float32).contiguous()

        g_color, g_depth, g_alpha = grad_in(g_color, (V, 3, H, W)), grad_in(g_depth, (V, 1, H, W)), \
            grad_in(g_alpha, (V, 1, H, W))
        # set_multistream(False) (measurement passes): everything on the calling stream, one view at a time — the views
        # then share one scratch block, so each view's per-Gaussian launch follows its blend backward directly
        streams = _streams_for(device, V) if _multistream else [main] * V
        # Groups of <= MAX_VIEWS views: their blend backwards run on the side streams (they overlap), then ONE fused
        # per-Gaussian kernel sums the group's contributions in registers and writes (first group) or accumulates
        # (later groups) every gradient tensor.  A side stream's scratch block holds the GGrad records of one view
        # at a time, so the next group's blend backward waits for this group's per-Gaussian kernel.
        G = min(_lib.MAX_VIEWS, MAX_SIDE_STREAMS) if _multistream else 1
        for v0 in range(0, V, G):
            n = min(G, V - v0)
            fork = torch.cuda.Event()
            fork.record(main)                # after the previous group's per-Gaussian kernel (and the exchange zeroing)
            staged = []
            for j in range(n):
                v = v0 + j
                rs, sv, st = settings_list[v], svs[v], streams[v]
                if st is not main:
                    st.wait_event(fork)
                with torch.cuda.stream(st):
                    ws = _workspace(device)
                    s, keep = _make_settings(rs, device, raw, None if tanfov_dev is None else tanfov_dev[v])
                    scratch = ws.ensure_scratch(sv.layout.scratch_bytes)
                    rc = lib.gsb_render_bwd(C.byref(s), P, sv.block.data_ptr(), scratch.data_ptr(), sv.d_cap,
                                            g_color[v].data_ptr(), g_depth[v].data_ptr(), g_alpha[v].data_ptr(),
                                            st.cuda_stream)
                    _lib.check(rc, "gsb_render_bwd")
                    done = torch.cuda.Event()
                    done.record(st)
                staged.append((s, keep, scratch, done))
            # Per-Gaussian stage in sub-groups of `sub` views: a sub-group's launch waits only for ITS blend backwards,
            # so it runs (memory-bound) underneath the blend backwards (issue-bound) of the views that follow; the
            # first launch writes the gradient tensors, later ones accumulate.  With the fused exchange every launch
            # adds into remote memory, so the whole group stays one launch there.
            sub = n if exchange is not None else max(1, min(n, _pbwd_group))
            for j0 in range(0, n, sub):
                m = min(sub, n - j0)
                for j in range(j0, j0 + m):
                    main.wait_event(staged[j][3])
                    svs[v0 + j].block.record_stream(main)
                sp = (C.POINTER(_lib.GsbSettings) * m)(*[C.pointer(staged[j][0]) for j in range(j0, j0 + m)])
                rp = (C.c_void_p * m)(*[radii[v0 + j].data_ptr() for j in range(j0, j0 + m)])
                svp = (C.c_void_p * m)(*[svs[v0 + j].block.data_ptr() for j in range(j0, j0 + m)])
                scp = (C.c_void_p * m)(*[staged[j][2].data_ptr() for j in range(j0, j0 + m)])
                dcp = (C.c_longlong * m)(*[svs[v0 + j].d_cap for j in range(j0, j0 + m)])
                if exchange is not None:
                    # the radii maximum (and, once per backward, the step's scalar) ride in the fused launch
                    exchange.set_aux(getattr(exchange, "pending_scalar", None), first_launch=v0 == 0)
                rc = lib.gsb_preprocess_bwd_views(m, sp, P, K, _ptr(means3D), _ptr(scales), _ptr(rotations),
                                                  _ptr(opacities), _ptr(shs), _ptr(colors), _ptr(cov3D), rp, svp, scp,
                                                  dcp, optr["means3D"], optr["means2D"], optr["shs"],
                                                  optr["colors"], optr["opacities"], optr["scales"],
                                                  optr["rotations"], optr["cov3D"],
                                                  exchange.mode if exchange is not None else int(v0 + j0 > 0),
                                                  main.cuda_stream)
                _lib.check(rc, "gsb_preprocess_bwd_views")
        if exchange is not None:
            exchange.end()                   # every rank's reds have landed: the local copy is the global sum
            if exchange.clone_outputs:
                out = {k: (None if t is None else t.clone()) for k, t in out.items()}
        # the side streams' scratch blocks are read by the calling stream above; they are persistent
        # per-(device, stream) workspaces, so no allocator hand-over is involved
        back = torch.cuda.Event()
        back.record(main)
        for st in set(streams):
            if st is not main:
                st.wait_event(back)          # next use of a side workspace is ordered after these reads
    return out


def _prepare_inputs(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp):
    none_if_empty = lambda t: None if (t is None or t.numel() == 0) else t
    sh, colors_precomp = none_if_empty(sh), none_if_empty(colors_precomp)
    scales, rotations, cov3Ds_precomp = none_if_empty(scales), none_if_empty(rotations), none_if_empty(cov3Ds_precomp)
    device = means3D.device
    means3D = _check(means3D, "means3D", (3,), device)
    P = means3D.shape[0]
    opacities = opacities.reshape(P, 1) if torch.is_tensor(opacities) and opacities.numel() == P else opacities
    opacities = _check(opacities, "opacities", (1,), device)
    if sh is not None:
        if sh.dim() != 3 or sh.shape[2] != 3:
            raise ValueError(f"shs must have shape [P, K, 3], got {tuple(sh.shape)}")
        sh = _check(sh, "shs", tuple(sh.shape[1:]), device)
    if colors_precomp is not None:
        colors_precomp = _check(colors_precomp, "colors_precomp", (3,), device)
    if scales is not None:
        scales = _check(scales, "scales", (3,), device)
    if rotations is not None:
        rotations = _check(rotations, "rotations", (4,), device)
    if cov3Ds_precomp is not None:
        cov3Ds_precomp = _check(cov3Ds_precomp, "cov3D_precomp", (6,), device)
    for name, t in (("shs", sh), ("colors_precomp", colors_precomp), ("scales", scales),
                    ("rotations", rotations), ("cov3D_precomp", cov3Ds_precomp)):
        if t is not None and t.shape[0] != P:
            raise ValueError(f"{name} has {t.shape[0]} rows, means3D has {P}")
    return means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp = _prepare_inputs(
            means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        color, radii, depth, alpha, sv = _forward_impl(raster_settings, means3D, sh, colors_precomp, opacities,
                                                       scales, rotations, cov3Ds_precomp,
                                                       forward_only=not any(ctx.needs_input_grad))
        ctx.raster_settings = raster_settings
        ctx.sv = sv
        ctx.present = (sh is not None, colors_precomp is not None, scales is not None, rotations is not None,
                       cov3Ds_precomp is not None)
        e = means3D.new_empty(0)
        ctx.save_for_backward(means3D, sh if sh is not None else e,
                              colors_precomp if colors_precomp is not None else e, opacities,
                              scales if scales is not None else e, rotations if rotations is not None else e,
                              cov3Ds_precomp if cov3Ds_precomp is not None else e, radii)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        means3D, sh, colors, opacities, scales, rotations, cov3D, radii = ctx.saved_tensors
        has_sh, has_col, has_sc, has_rot, has_cov = ctx.present
        out = _backward_impl(ctx.raster_settings, ctx.sv, means3D, sh if has_sh else None,
                             colors if has_col else None, opacities, scales if has_sc else None,
                             rotations if has_rot else None, cov3D if has_cov else None, radii,
                             grad_color, grad_depth, grad_alpha)
        return (out["means3D"], out["means2D"], out["shs"], out["colors"], out["opacities"], out["scales"],
                out["rotations"], out["cov3D"], None)


class _RasterizeViews(torch.autograd.Function):
    """V views of the same Gaussians in one autograd node (additive API).

    Every view runs exactly the kernels of the single-view operator; what changes is the host
    side: one node instead of V, outputs written into slices of stacked tensors, and the
    backward of view v > 0 ACCUMULATES (beta = 1) into the gradients view 0 wrote, so the sum
    over views — which is what the optimiser and the densification statistics consume
    (threestudio/systems/GaussianIP.py:450-457) — is formed inside preprocess_bwd instead of by
    V autograd AccumulateGrad passes per tensor."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                settings_list, exchange=None, raw=0, tanfov_dev=None):
        means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp = _prepare_inputs(
            means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        ctx.exchange = exchange
        ctx.tanfov_dev = tanfov_dev
        tf = (lambda v: None) if tanfov_dev is None else (lambda v: tanfov_dev[v])
        ctx.raw = raw = int(raw)
        if cov3Ds_precomp is not None:
            raw &= ~(_lib.RAW_SCALE | _lib.RAW_ROTATION)
            ctx.raw = raw
        V = len(settings_list)
        fo = not any(ctx.needs_input_grad)          # inference: no hit records, shorter saved blocks
        if V == 0:
            raise ValueError("no views")
        H, W = int(settings_list[0].image_height), int(settings_list[0].image_width)
        for rs in settings_list:
            if (int(rs.image_height), int(rs.image_width)) != (H, W):
                raise ValueError("all views of one call must have the same resolution")
        dev, P = means3D.device, means3D.shape[0]
        color = torch.empty(V, 3, H, W, dtype=torch.float32, device=dev)
        depth = torch.empty(V, 1, H, W, dtype=torch.float32, device=dev)
        alpha = torch.empty(V, 1, H, W, dtype=torch.float32, device=dev)
        radii = torch.empty(V, P, dtype=torch.int32, device=dev)
        svs = []
        ctx.multistream = _multistream and V > 1
        if ctx.multistream:
            main = torch.cuda.current_stream(dev)
            streams = _streams_for(dev, V)
            fork = torch.cuda.Event()
            fork.record(main)
            for st in set(streams):
                st.wait_event(fork)
            # groups of MAX_SIDE_STREAMS views: a stream's counts block / event belongs to one pending view at a time
            for v0 in range(0, V, MAX_SIDE_STREAMS):
                calls = []
                for v in range(v0, min(V, v0 + MAX_SIDE_STREAMS)):
                    with torch.cuda.stream(streams[v]):
                        call = _ForwardCall(settings_list[v], means3D, sh, colors_precomp, opacities, scales,
                                            rotations, cov3Ds_precomp,
                                            out=(color[v], radii[v], depth[v], alpha[v]), raw=raw,
                                            tanfov_dev=tf(v), forward_only=fo)
                        if _fwd_stagger > 0:
                            # Staggered start: view v's per-Gaussian + binning stages (latency-bound, they leave most
                            # of the machine idle) begin when view v - k's binning is done, i.e. they run UNDER that
                            # view's blend (issue-bound) instead of in lockstep with the other views' binning.
                            call.mark_binned = True
                            j = len(calls) - _fwd_stagger
                            if j >= 0 and calls[j].binned is not None:
                                streams[v].wait_event(calls[j].binned)
                        calls.append(call.enqueue())
                for call in calls:
                    svs.append(call.finish()[4])
            for st in set(streams):
                join = torch.cuda.Event()
                join.record(st)
                main.wait_event(join)
        else:
            for v, rs in enumerate(settings_list):
                _, _, _, _, sv = _forward_impl(rs, means3D, sh, colors_precomp, opacities, scales, rotations,
                                               cov3Ds_precomp, out=(color[v], radii[v], depth[v], alpha[v]), raw=raw,
                                               tanfov_dev=tf(v), forward_only=fo)
                svs.append(sv)
        ctx.settings_list, ctx.svs = list(settings_list), svs
        ctx.present = (sh is not None, colors_precomp is not None, scales is not None, rotations is not None,
                       cov3Ds_precomp is not None)
        e = means3D.new_empty(0)
        ctx.save_for_backward(means3D, sh if sh is not None else e,
                              colors_precomp if colors_precomp is not None else e, opacities,
                              scales if scales is not None else e, rotations if rotations is not None else e,
                              cov3Ds_precomp if cov3Ds_precomp is not None else e, radii)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        means3D, sh, colors, opacities, scales, rotations, cov3D, radii = ctx.saved_tensors
        has_sh, has_col, has_sc, has_rot, has_cov = ctx.present
        if ctx.multistream or ctx.exchange is not None:
            out = _backward_views(ctx.settings_list, ctx.svs, means3D, sh if has_sh else None,
                                  colors if has_col else None, opacities, scales if has_sc else None,
                                  rotations if has_rot else None, cov3D if has_cov else None, radii,
                                  grad_color, grad_depth, grad_alpha, exchange=ctx.exchange, raw=ctx.raw,
                                  tanfov_dev=ctx.tanfov_dev)
            return (out["means3D"], out["means2D"], out["shs"], out["colors"], out["opacities"], out["scales"],
                    out["rotations"], out["cov3D"], None, None, None, None)
        out = None
        for v, (rs, sv) in enumerate(zip(ctx.settings_list, ctx.svs)):
            out = _backward_impl(rs, sv, means3D, sh if has_sh else None, colors if has_col else None, opacities,
                                 scales if has_sc else None, rotations if has_rot else None,
                                 cov3D if has_cov else None, radii[v],
                                 None if grad_color is None else grad_color[v],
                                 None if grad_depth is None else grad_depth[v],
                                 None if grad_alpha is None else grad_alpha[v], out=out, accumulate=v > 0,
                                 raw=ctx.raw, tanfov_dev=None if ctx.tanfov_dev is None else ctx.tanfov_dev[v])
        return (out["means3D"], out["means2D"], out["shs"], out["colors"], out["opacities"], out["scales"],
                out["rotations"], out["cov3D"], None, None, None, None)


def rasterize_views(settings_list, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                    rotations=None, cov3D_precomp=None, exchange=None, raw_inputs: int = 0, tanfov_dev=None):
    """Batched form of GaussianRasterizer(...)(...): returns stacked (color [V,3,H,W], radii [V,P],
    depth [V,1,H,W], alpha [V,1,H,W]); means2D.grad receives the SUM over views.  With
    ``exchange`` (gaussianip_b200.exchange.GradExchange) the input gradients returned by the backward
    are already summed over the ranks of the exchange's group (reduction fused into the kernel).
    ``raw_inputs`` (bits _lib.RAW_OPACITY | RAW_SCALE | RAW_ROTATION): the flagged inputs are the model's raw
    parameters (logits / log-scales / unnormalised quaternions); the kernels apply sigmoid / exp / normalize
    and the returned gradients are with respect to the raw tensors (include/gsb.h GSB_RAW_*).
    ``tanfov_dev`` (optional, one device tensor [2] = (tanfovx, tanfovy) per view): the kernels read the intrinsics
    from device memory instead of ``settings.tanfovx / tanfovy`` — needed when the call is captured in a CUDA
    graph and replayed with other cameras (gaussianip_b200.graph, cameras.CameraBlock)."""
    if tanfov_dev is not None and len(tanfov_dev) != len(settings_list):
        raise ValueError("tanfov_dev needs one entry per view")
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    return _RasterizeViews.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                 cov3D_precomp, tuple(settings_list), exchange, int(raw_inputs),
                                 None if tanfov_dev is None else tuple(tanfov_dev))


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Boolean mask of points in front of the camera (view z > 0.2)."""
        with torch.no_grad():
            rs = self.raster_settings
            positions = _check(positions, "positions", (3,), positions.device)
            if positions.device.type != "cuda":
                raise ValueError("markVisible needs CUDA tensors (no CPU fallback)")
            P = positions.shape[0]
            view = _small(rs.viewmatrix, 16, "viewmatrix", positions.device)
            proj = _small(rs.projmatrix, 16, "projmatrix", positions.device)
            present = torch.empty(P, dtype=torch.uint8, device=positions.device)
            with torch.cuda.device(positions.device):
                st = torch.cuda.current_stream(positions.device).cuda_stream
                _lib.check(_lib.load().gsb_mark_visible(P, positions.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                        present.data_ptr(), st), "gsb_mark_visible")
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs)
