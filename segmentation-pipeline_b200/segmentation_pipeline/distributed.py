"""Multi-GPU sharding of the sliding-window path (one process per GPU, ``torch.distributed``).

The reference has no distributed code at all; the path shards naturally (section 8e):

* **cohort mode** -- subjects are independent (reference prediction.py:131): subject i goes to rank i mod N; the
  only collective is the all-reduce of the tiny int64 confusion matrices.
* **z-slab mode** -- one volume, the sorted patch grid split along the first spatial axis.  Rank r evaluates the
  patches whose start plane falls into its range and accumulates them into a local slab; a patch reaches up to
  ``patch - 1`` planes past the range, and those partial sums ("halo") are sent forward to the rank that owns the
  planes.  Counts need no exchange (they are analytic).  Each rank finalises the planes it owns and the uint8
  label slabs are all-gathered.

The exchange logic is independent of where the arithmetic runs: ``SlabOps`` abstracts the four device steps so
that the CPU tests (gloo, world_size 2) can drive the same plan / send / recv / gather code with oracle
arithmetic, while production uses the CUDA kernels (``CudaSlabOps``).  Summation order: the halo is added to the
owner's accumulator AFTER its own patches, i.e. (sum of own patches) + (partial sum received), which differs
from the single-GPU order by fp32 re-association only ('determinism').
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .grid import PatchGrid


# ------------------------------------------------------------------------------------------------- cohort mode
def shard_subjects(n_subjects: int, rank: int, world: int) -> List[int]:
    """Indices of the subjects rank ``rank`` processes (round-robin)."""
    return list(range(rank, n_subjects, world))


def all_reduce_confusion(cm: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank int64 confusion matrices (bit-exact: integer addition)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group)
    return cm


# ------------------------------------------------------------------------------------------------- z-slab plan
@dataclass
class SlabPlan:
    """Partition of a PatchGrid along axis 0 (padded coordinates)."""
    grid: PatchGrid
    world: int
    start_ranges: List[Tuple[int, int]]   # per rank: [first, last) index into grid.axis_starts[0]
    own: List[Tuple[int, int]]            # per rank: owned padded planes [lo, hi)
    local: List[Tuple[int, int]]          # per rank: planes its local accumulator covers [lo, hi)

    def patches_of(self, rank: int) -> List[Tuple[int, ...]]:
        first, last = self.start_ranges[rank]
        starts = set(self.grid.axis_starts[0][first:last])
        return [loc for loc in self.grid.locations if loc[0] in starts]

    def sends(self, rank: int) -> List[Tuple[int, int, int]]:
        """(destination rank, plane lo, plane hi) of every halo slab rank ``rank`` must send forward."""
        lo, hi = self.local[rank]
        out = []
        for dst in range(rank + 1, self.world):
            a, b = max(self.own[dst][0], lo), min(self.own[dst][1], hi)
            if a < b:
                out.append((dst, a, b))
        return out

    def recvs(self, rank: int) -> List[Tuple[int, int, int]]:
        """(source rank, plane lo, plane hi) of every halo slab rank ``rank`` receives."""
        out = []
        for src in range(rank):
            for dst, a, b in self.sends(src):
                if dst == rank:
                    out.append((src, a, b))
        return out

    def owned_output(self, rank: int) -> Tuple[int, int]:
        """Unpadded output planes [lo, hi) rank ``rank`` finalises (may be empty)."""
        b = self.grid.border[0]
        w = self.grid.spatial_shape[0]
        lo, hi = self.own[rank]
        return min(max(lo - b, 0), w), min(max(hi - b, 0), w)


def make_slab_plan(grid: PatchGrid, world: int) -> SlabPlan:
    starts = grid.axis_starts[0]
    n = len(starts)
    if world < 1:
        raise ValueError("world must be >= 1")
    # contiguous, balanced split of the start planes (ranks beyond the number of start planes get nothing)
    ranges = []
    for r in range(world):
        a = (r * n) // world
        b = ((r + 1) * n) // world
        ranges.append((a, b))
    p0 = grid.patch_size[0]
    total = grid.padded_shape[0]
    own, local = [], []
    for r, (a, b) in enumerate(ranges):
        if a == b:
            own.append((0, 0))
            local.append((0, 0))
            continue
        lo = starts[a] if a > 0 else 0
        # the next non-empty rank's first start plane bounds the owned range
        nxt = next((starts[ra] for ra, rb in ranges[r + 1:] if ra < rb), None)
        hi = total if nxt is None else nxt
        own.append((lo, hi))
        local.append((starts[a], starts[b - 1] + p0))
    return SlabPlan(grid, world, ranges, own, local)


# ------------------------------------------------------------------------------------------------- device steps
class SlabOps:
    """The arithmetic of one rank, on whatever device the tensors live on."""

    def forward_accumulate(self, volume: torch.Tensor, grid: PatchGrid, patches, plane0: int, n_planes: int
                           ) -> torch.Tensor:
        """Extract + network forward + overlap-add of ``patches`` into a fresh fp32 accumulator
        (C_out, n_planes, PH, PD) whose first plane is padded plane ``plane0``."""
        raise NotImplementedError

    def add_slab(self, acc: torch.Tensor, slab: torch.Tensor, plane_offset: int) -> None:
        """acc[:, plane_offset : plane_offset + slab.shape[1]] += slab"""
        raise NotImplementedError

    def finalize(self, acc: torch.Tensor, grid: PatchGrid, plane0: int, out_lo: int, out_hi: int
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(probs (C, out_hi-out_lo, H, D) fp32, labels (out_hi-out_lo, H, D) uint8) of unpadded planes
        [out_lo, out_hi)."""
        raise NotImplementedError


class CudaSlabOps(SlabOps):
    """Production ops: libb200seg kernels + the native network plan."""

    def __init__(self, model, patch_batch_size: int = 16):
        self.model = model
        self.patch_batch_size = patch_batch_size

    def forward_accumulate(self, volume, grid, patches, plane0, n_planes):
        import b200seg
        from .models import _engine
        device = volume.device
        precision = _engine._resolve_precision(self.model, volume)
        compiled = _engine.compiled_for(self.model, precision, device)
        p0, p1, p2 = grid.patch_size
        acc = None
        for start in range(0, len(patches), self.patch_batch_size):
            locs = patches[start:start + self.patch_batch_size]
            buf = compiled.input_buffer(len(locs), p0, p1, p2)
            b200seg.grid_extract(volume, locs, grid.border, grid.pad_mode_code, grid.pad_value,
                                 buf.view(volume.shape[0]))
            y = compiled.run_blocked(len(locs), p0, p1, p2)
            if acc is None:
                acc = torch.zeros((y.shape[1], n_planes, *grid.padded_shape[1:]), dtype=torch.float32, device=device)
            shifted = [(l[0] - plane0, l[1], l[2], l[3] - plane0, l[4], l[5]) for l in locs]
            b200seg.overlap_add(acc, y, shifted)
        return acc

    def add_slab(self, acc, slab, plane_offset):
        import b200seg
        n = slab.shape[1]
        loc = (plane_offset, 0, 0, plane_offset + n, slab.shape[2], slab.shape[3])
        b200seg.overlap_add(acc, slab[None].contiguous(), [loc])

    def finalize(self, acc, grid, plane0, out_lo, out_hi):
        import b200seg
        device = acc.device
        counts = [torch.tensor(c, dtype=torch.int32, device=device) for c in grid.axis_counts()]
        b = grid.border
        n = out_hi - out_lo
        h, d = grid.spatial_shape[1:]
        probs = torch.empty((acc.shape[0], n, h, d), dtype=torch.float32, device=device)
        labels = torch.empty((n, h, d), dtype=torch.uint8, device=device)
        offset = (out_lo + b[0] - plane0, b[1], b[2])
        b200seg.finalize_region(acc, counts, offset, (n, h, d), probs, None, labels, count_offset=plane0)
        return probs, labels


# ------------------------------------------------------------------------------------------------- z-slab driver
def slab_predict(volume: torch.Tensor, grid: PatchGrid, ops: SlabOps, group=None, gather_probs: bool = False):
    """Runs this rank's share of the sliding window and returns (labels uint8 (W, H, D) on every rank,
    probs fp32 (C, W, H, D) or None).  ``volume`` is the full (C, W, H, D) volume on this rank's device."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = make_slab_plan(grid, world)
    patches = plan.patches_of(rank)
    lo, hi = plan.local[rank]
    acc = ops.forward_accumulate(volume, grid, patches, lo, hi - lo) if patches else None

    # ---- halo exchange: partial sums flow forward only (a patch never reaches planes before its start)
    pending = []
    send_bufs = []
    if world > 1:
        for dst, a, b in plan.sends(rank):
            slab = acc[:, a - lo:b - lo].contiguous()
            send_bufs.append(slab)
            pending.append(dist.isend(slab, dst, group=group))
        recv_bufs = []
        for src, a, b in plan.recvs(rank):
            buf = torch.empty((acc.shape[0], b - a, *acc.shape[2:]), dtype=acc.dtype, device=acc.device)
            recv_bufs.append((a, buf))
            pending.append(dist.irecv(buf, src, group=group))
        for req in pending:
            req.wait()
        for a, buf in recv_bufs:
            ops.add_slab(acc, buf, a - lo)

    # ---- finalise the owned planes, gather the label slabs
    out_lo, out_hi = plan.owned_output(rank)
    w, h, d = grid.spatial_shape
    device = volume.device
    if out_hi > out_lo:
        probs, labels = ops.finalize(acc, grid, lo, out_lo, out_hi)
    else:
        c_out = 0 if acc is None else acc.shape[0]
        probs = torch.empty((c_out, 0, h, d), dtype=torch.float32, device=device)
        labels = torch.empty((0, h, d), dtype=torch.uint8, device=device)
    if world == 1:
        return labels, (probs if gather_probs else None)
    # equal-size all_gather: pad every slab to the largest one
    spans = [plan.owned_output(r) for r in range(world)]
    longest = max(b - a for a, b in spans)
    padded = torch.zeros((longest, h, d), dtype=torch.uint8, device=device)
    padded[:labels.shape[0]] = labels
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    full = torch.empty((w, h, d), dtype=torch.uint8, device=device)
    for (a, b), part in zip(spans, gathered):
        full[a:b] = part[:b - a]
    full_probs = None
    if gather_probs:
        c_out = torch.tensor([probs.shape[0]], device=device)
        dist.all_reduce(c_out, op=dist.ReduceOp.MAX, group=group)
        c = int(c_out.item())
        pp = torch.zeros((c, longest, h, d), dtype=torch.float32, device=device)
        pp[:, :probs.shape[1]] = probs if probs.shape[0] == c else 0
        parts = [torch.empty_like(pp) for _ in range(world)]
        dist.all_gather(parts, pp, group=group)
        full_probs = torch.empty((c, w, h, d), dtype=torch.float32, device=device)
        for (a, b), part in zip(spans, parts):
            full_probs[:, a:b] = part[:, :b - a]
    return full, full_probs
