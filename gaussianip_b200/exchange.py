"""Gradient exchange fused into the backward kernel (NVLink 5 / NVSwitch).

With views sharded over the GPUs of one box, the only exchange of a step is the SUM over ranks of the
rasterizer's input gradients.  Instead of letting the per-Gaussian backward kernel write local
gradients and calling NCCL afterwards, the kernel's epilogue issues ``multimem.red.add`` to the NVLS
MULTICAST address of a symmetric buffer that is mapped on every GPU: the NVSwitch adds each
contribution into every rank's copy while the kernel is still computing other Gaussians, so when all
ranks' kernels have finished every copy already holds the reduced gradient (csrc/preprocess_bwd.cu,
accumulate mode 2).  Zero-valued rows (Gaussians no local view sees) are not sent at all.

The sum is taken over the gradients with respect to the rasterizer's INPUTS (activated scales,
opacities, normalised rotations ...).  The activation backward that follows is linear in the incoming
gradient and identical on all ranks (parameters are replicated), so applying it to the reduced
gradient yields exactly the reduced leaf gradients — no further exchange is needed.

Two algorithms (csrc/preprocess_bwd.cu, `accumulate` 2 and 3):
  "push_all"    the epilogue multicasts every row into all copies.  One phase, but every GPU receives N
                gradients, so it only pays for 2 ranks (measured: +3.5 % views/s at N=2, -3.5 % at N=8).
  "owner_push"  rows are owned by ranks in contiguous blocks; the epilogue adds each row into the OWNER's
                copy (plain red.global to peer memory), then every owner multicasts its reduced block to all
                copies with a small second kernel.  Every GPU receives ~2 gradients whatever N is.
Protocol per backward (all on the calling stream, no host synchronisation); the buffer has two halves used in turn:
  zero the OTHER half (next step's) -> [blend backward on side streams] -> fused kernel into this half ->
  barrier(all ranks' kernels done) [-> gather kernel -> barrier] -> consumers read the local copy.

Buffers come from torch.distributed._symmetric_memory (plumbing: allocation, handle exchange, signal-pad
barriers); the reduction itself is our kernel.  Requires NVLS multicast support; raises otherwise.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

_FIELDS = (("means3D", 3), ("means2D", 3), ("opacities", 1), ("shs", None), ("colors", 3), ("scales", 3),
           ("rotations", 4), ("cov3D", 6), ("radii", 1), ("scalars", None))
SCALAR_SLOTS = 64        # the "scalars" field: a few floats summed over ranks (slot 0: the step's loss)


def plan_layout(P: int, K: int, present: Dict[str, bool]):
    """Field offsets (in floats, every field 256 B aligned) of one half of the symmetric buffer.
    Returns ({name: (offset, shape)}, total_floats)."""
    off, layout = 0, {}
    for name, w in _FIELDS:
        if not present.get(name, False):
            continue
        shape = (P, K, 3) if name == "shs" else ((SCALAR_SLOTS,) if name == "scalars" else (P, w))
        n = 1
        for d in shape:
            n *= d
        layout[name] = (off, shape)
        off += (n + 63) // 64 * 64
    return layout, off


def plan_ownership(P: int, layout, rank: int, world: int):
    """owner_push: rank r owns rows [r * rows_per_rank, (r+1) * rows_per_rank) (the last rank also the tail);
    rows_per_rank is a multiple of 32 so the 32 rows of a warp have one owner.  Returns (rows_per_rank,
    [(offset_floats, count_floats)] of this rank's block in every field)."""
    rows_per_rank = max((-(-P // world) + 31) // 32 * 32, 32)
    r0 = min(rank * rows_per_rank, P)
    r1 = P if rank == world - 1 else min(r0 + rows_per_rank, P)
    segments = []
    for name, (foff, shape) in layout.items():
        if name == "scalars":                 # owned by rank 0 as a whole
            if rank == 0:
                segments.append((foff, shape[0]))
            continue
        w = 1
        for d in shape[1:]:
            w *= d
        if r1 > r0 and w > 0:
            segments.append((foff + r0 * w, (r1 - r0) * w))
    return rows_per_rank, segments


class GradExchange:
    def __init__(self, group=None, clone_outputs: bool = True, algorithm: str = "auto"):
        """clone_outputs=False hands the backward's consumers views of the symmetric buffer itself (valid until
        the next backward zeroes it).  Safe when every leaf already has a preallocated ``.grad`` that autograd
        accumulates INTO (ViewParallel's bucket), because then nothing keeps a reference to the views."""
        self.clone_outputs = clone_outputs
        self.algorithm = algorithm
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("GradExchange needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.algorithm == "auto":
            self.algorithm = "push_all" if self.world <= 2 else "owner_push"
        if self.algorithm not in ("push_all", "owner_push"):
            raise ValueError("algorithm must be 'auto', 'push_all' or 'owner_push'")
        self.mode = 2 if self.algorithm == "push_all" else 3      # `accumulate` value of gsb_preprocess_bwd_views
        self.key = None
        self.buf = None
        self.hdl = None
        self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self.steps = 0
        # static = True: graph-safe protocol.  A captured step always runs the same instructions, so the halves
        # cannot alternate: half 0 is used every time, zeroed in-stream at the start of the backward, with one more
        # barrier so that nobody adds into a copy that is still being zeroed.
        self.static = False

    @staticmethod
    def available(device: torch.device, group=None) -> bool:
        """COLLECTIVE probe: can every rank of the group map a symmetric buffer with NVLS multicast?  All ranks
        get the same answer (the local results are combined with a MIN all-reduce)."""
        ok = 0
        if torch.cuda.is_available() and dist.is_available() and dist.is_initialized():
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(1024, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
                ok = int(bool(getattr(hdl, "has_multicast_support", False)) and bool(hdl.multicast_ptr))
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(flag.item())

    def ensure(self, P: int, K: int, present: Dict[str, bool], device: torch.device) -> None:
        """(Re)allocate when the shapes change.  COLLECTIVE: every rank must call it with the same shapes."""
        key = (P, K, tuple(sorted(k for k, v in present.items() if v)), device.index)
        if key == self.key:
            return
        import torch.distributed._symmetric_memory as symm_mem
        layout, off = plan_layout(P, K, present)
        # two halves used alternately: the half of step k+1 is zeroed during step k and step k's closing barrier
        # tells every rank so, which saves the "everyone has zeroed" barrier in front of each kernel
        self.half = max(off, 64)
        self.buf = symm_mem.empty(2 * self.half, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        if not getattr(self.hdl, "has_multicast_support", False) or not self.hdl.multicast_ptr:
            raise RuntimeError("NVLS multicast is not available on this system; use the NCCL exchange")
        self.offsets, self.key = layout, key
        self.rows_per_rank, self.segments = plan_ownership(P, layout, self.rank, self.world)
        if self.mode == 3:
            import ctypes as C
            from . import _lib
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            arr = (C.c_void_p * self.world)(*ptrs)
            _lib.check(_lib.load().gsb_exchange_config(self.world, self.rank, self.rows_per_rank, arr),
                       "gsb_exchange_config")
        self.buf.zero_()
        self.hdl.barrier(channel=0)          # both halves are zero everywhere before the first kernel
        self.cur = 0

    def begin(self) -> None:
        """Pick this step's half (already zero on every rank) and zero the other one for the next step; the
        zeroing is ordered before this step's closing barrier, so nobody can add into it too early."""
        if self.static:
            self.cur = 0
            self.buf[:self.half].zero_()
            self.hdl.barrier(channel=0)
            return
        self.cur = self.steps & 1
        other = 1 - self.cur
        self.buf[other * self.half:(other + 1) * self.half].zero_()

    def end(self) -> None:
        """All ranks' kernels (and therefore all their reds) are complete; owner_push then redistributes."""
        self.hdl.barrier(channel=1)
        if self.mode == 3:
            # every rank takes part in the closing barrier, also one that owns no rows (P <= 32 (world - 1)):
            # gsb_exchange_gather with zero segments launches nothing
            import ctypes as C
            from . import _lib
            n = len(self.segments)
            off = (C.c_longlong * max(n, 1))(*[o + self.cur * self.half for o, _ in self.segments])
            cnt = (C.c_longlong * max(n, 1))(*[c for _, c in self.segments])
            dev = self.buf.device
            with torch.cuda.device(dev):
                _lib.check(_lib.load().gsb_exchange_gather(self.buf.data_ptr(), int(self.hdl.multicast_ptr), n, off, cnt,
                                                           torch.cuda.current_stream(dev).cuda_stream),
                           "gsb_exchange_gather")
            self.hdl.barrier(channel=2)
        self.steps += 1

    def reset(self) -> None:
        """COLLECTIVE: zero both halves everywhere (after switching `static`, or to resynchronise the alternation)."""
        if self.buf is None:
            return
        self.hdl.barrier(channel=0)
        self.buf.zero_()
        self.hdl.barrier(channel=0)
        self.steps = 0

    def set_aux(self, loss: Optional[torch.Tensor], first_launch: bool = True) -> None:
        """Arm the extras of the next fused launch: the radii maximum always (when the buffer has the field), the
        scalar `loss` only on the first launch of a backward (it must be added once)."""
        import ctypes as C
        from . import _lib
        radii = self.output_ptr("radii")
        sc = self.output_ptr("scalars") if (loss is not None and first_launch) else None
        _lib.check(_lib.load().gsb_exchange_set_aux(radii, loss.data_ptr() if sc is not None else None, sc),
                   "gsb_exchange_set_aux")

    def output_ptr(self, name: str) -> Optional[int]:
        """What the kernel is given for this field: the multicast address (push_all) or the local copy (owner_push)."""
        if name not in self.offsets:
            return None
        base = int(self.hdl.multicast_ptr) if self.mode == 2 else self.buf.data_ptr()
        return base + 4 * (self.offsets[name][0] + self.cur * self.half)

    def local(self, name: str) -> Optional[torch.Tensor]:
        if name not in self.offsets:
            return None
        off, shape = self.offsets[name]
        off += self.cur * self.half
        n = 1
        for d in shape:
            n *= d
        return self.buf[off:off + n].view(shape)

    def nbytes(self) -> int:
        return 0 if self.buf is None else self.half * 4
