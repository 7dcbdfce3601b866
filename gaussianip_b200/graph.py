"""CUDA-graph capture of a whole rendering step.

The reference drives every view of an optimisation step from Python (threestudio/systems/GaussianIP.py:154-159,
296-308): one ``render()`` per view, each a dozen kernel launches.  Here a step's work on the device — the forward of
all local views on their side streams, the loss, the blend backward per view, the fused per-Gaussian backward, the
gradient exchange — is a fixed sequence of ~80 kernels / copies whose SHAPES do not depend on the data (instance
buffers have a capacity, D stays on the device), so it is captured once into a ``torch.cuda.CUDAGraph`` and replayed:
one ``cudaGraphLaunch`` per step instead of ~80 ctypes / torch launches (host enqueue 2.9 ms -> < 0.2 ms per
4-view step).

What is fixed at capture time: every tensor ADDRESS the step touches (parameters, camera block, loss weights,
outputs, gradients, saved blocks — allocated from the graph's private pool) and the instance capacity ``D_cap``.  What
may change between replays: the CONTENTS of those tensors (new parameters, new cameras written into the camera
block).  Each forward copies its 32-byte counts block to its own pinned slot inside the graph; ``validate()`` reads
them after a replay and, if a view's D exceeded the capacity, raises the capacity and re-captures — the step is then
replayed again (its results were not yet consumed), exactly the redo-on-overflow rule of ``rasterizer.speculation``.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import rasterizer


class CapturedStep:
    """``body()`` enqueues one whole step on the CURRENT stream (and the rasterizer's side streams) and returns a
    dict / tuple / tensor of results; all its inputs must live at fixed addresses.  ``replay()`` re-runs it."""

    def __init__(self, body: Callable[[], object], device: Optional[torch.device] = None, max_forwards: int = 64,
                 warmup: int = 2):
        self.body = body
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)     # capture stream (also used for the eager warm-up)
        # pinned counts rows, allocated BEFORE any capture (cudaHostAlloc is not allowed while capturing)
        self.counts_pool = torch.zeros(max_forwards, 8, dtype=torch.int32).pin_memory()
        self.warmup = max(1, int(warmup))
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.spec = None
        self.outputs = None
        self._events = [torch.cuda.Event() for _ in range(4)]      # one per replay in flight
        self.captures = 0
        self.replays = 0
        self.launches_per_capture = 0       # libgsb kernels enqueued by one pass of the body (= per replay)

    # ---- capture -------------------------------------------------------------------------------------------
    def _eager_warmup(self) -> None:
        """Run the body eagerly on the capture stream until every view fits its capacity: creates the per-stream
        workspaces (pinned memory, events), configures the kernels' launch attributes and sizes the scratch."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            done = 0
            for _ in range(8 + self.warmup):
                with rasterizer.speculation() as spec:
                    out = self.body()
                fits = spec.validate()
                del out
                done = done + 1 if fits else 0
                if done >= self.warmup:
                    break
            else:
                raise RuntimeError("instance capacity kept overflowing during the warm-up")
        cur.wait_stream(self.stream)
        torch.cuda.synchronize(self.device)

    def capture(self) -> "CapturedStep":
        self._eager_warmup()
        self.graph = None
        self.outputs = None
        g = torch.cuda.CUDAGraph()
        spec = rasterizer.speculation(capture=True, counts_pool=self.counts_pool)
        from . import _lib
        n0 = _lib.launch_count()
        with torch.cuda.graph(g, stream=self.stream):
            with spec:
                self.outputs = self.body()
        self.launches_per_capture = _lib.launch_count() - n0
        self.graph, self.spec = g, spec
        self.captures += 1
        return self

    # ---- replay --------------------------------------------------------------------------------------------
    def replay(self):
        """Launch the captured step on the current stream; returns the (static) result tensors of ``body``.
        Asynchronous.  ``validate()`` must pass before the results are consumed; the host may run one replay
        ahead (``validate(i)`` for replay i after replay i + 1 has been launched)."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        self._events[self.replays % len(self._events)].record(torch.cuda.current_stream(self.device))
        self.replays += 1
        return self.outputs

    def validate(self, replay_index: Optional[int] = None) -> bool:
        """Wait for replay ``replay_index`` (default: the last one) and check that every view's instance count
        fitted the captured capacity.  On False the capacity has been raised and the graph re-captured: the step's
        results are invalid, replay it again."""
        if self.replays == 0:
            return True
        idx = self.replays - 1 if replay_index is None else int(replay_index)
        if idx < self.replays - len(self._events):
            raise ValueError("that replay's completion event has been reused")
        self._events[idx % len(self._events)].synchronize()
        if self.spec.validate(keep=True):
            return True
        torch.cuda.synchronize(self.device)
        self.capture()
        return False

    def run(self):
        """replay() until validate() passes (normally once)."""
        for _ in range(8):
            out = self.replay()
            if self.validate():
                return out
        raise RuntimeError("instance capacity overflowed on 8 consecutive replays")
