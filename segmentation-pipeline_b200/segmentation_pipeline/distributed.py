"""Multi-GPU sharding of the sliding-window path (one process per GPU, ``torch.distributed``).

The reference has no distributed code at all; the path shards naturally (section 8e):

* **cohort mode** -- subjects are independent (reference prediction.py:131): subject i goes to rank i mod N; the
  only collective is the all-reduce of the tiny int64 confusion matrices.
* **z-slab mode** -- one volume over N GPUs.  Two decoupled partitions (``SlabPlan``):

    - *compute*: the sorted patch list (torchio's order) is cut into N contiguous runs of equal length, so every rank
      evaluates floor/ceil(n / N) patches whatever the shape of the grid (config 3: 125 patches -> 15 or 16 each at
      8 ranks; splitting by start PLANES, as round 1 did, left ranks idle as soon as N exceeded the 5 start planes);
    - *ownership*: the padded output volume is cut into N slabs of planes along the first spatial axis; the owner of
      a slab accumulates, divides, crops and arg-maxes it.

  Every plane of every patch output belongs to exactly one owner, so after its forwards a rank stages, per patch and
  owner, the dense block of planes that falls into that owner's slab (``b200seg_copy_planes``) and ONE
  ``all_to_all_single`` over NCCL / NVLink moves all blocks (about 0.55 x the rank's outputs, < 1 ms at 8 GPUs for
  config 3).  The owner then adds the blocks it holds IN SORTED PATCH ORDER -- exactly the per-voxel order of the
  single-GPU accumulation and of torchio's ``add_batch`` -- so probabilities and labels are bit-identical to one GPU
  (raw contributions are exchanged, not partial sums: partial sums would re-associate the fp32 additions).  Counts
  need no exchange (they are analytic).  The uint8 label slabs are all-gathered.

The exchange logic is independent of where the arithmetic runs: ``SlabOps`` abstracts the device steps so that the CPU
tests (gloo, world_size 2 / 3) drive the same plan / staging / all-to-all / gather code with oracle arithmetic, while
production uses the CUDA kernels (``CudaSlabOps``).
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
class Block:
    """Planes [lo, hi) (padded coordinates) of patch ``patch`` -- computed on ``src``, accumulated by ``dst``."""
    patch: int
    src: int
    dst: int
    lo: int
    hi: int


@dataclass
class SlabPlan:
    grid: PatchGrid
    world: int
    runs: List[Tuple[int, int]]     # per rank: [first, last) indices into grid.locations (sorted order)
    own: List[Tuple[int, int]]      # per rank: owned padded planes [lo, hi) along axis 0

    def patches_of(self, rank: int) -> List[Tuple[int, ...]]:
        a, b = self.runs[rank]
        return self.grid.locations[a:b]

    def rank_of_patch(self, index: int) -> int:
        for r, (a, b) in enumerate(self.runs):
            if a <= index < b:
                return r
        raise IndexError(index)

    def blocks(self) -> List[Block]:
        """Every (patch, owner) intersection, ordered by (src, dst, patch): the order of the send buffers."""
        out = []
        for src, (a, b) in enumerate(self.runs):
            for dst, (lo, hi) in enumerate(self.own):
                for p in range(a, b):
                    i0, i1 = self.grid.locations[p][0], self.grid.locations[p][3]
                    x, y = max(i0, lo), min(i1, hi)
                    if x < y:
                        out.append(Block(p, src, dst, x, y))
        return out

    def plane_elems(self, channels: int) -> int:
        return channels * self.grid.patch_size[1] * self.grid.patch_size[2]

    def split_sizes(self, rank: int, channels: int):
        """(elements this rank sends to each rank, elements it receives from each rank)."""
        per = self.plane_elems(channels)
        send, recv = [0] * self.world, [0] * self.world
        for blk in self.blocks():
            n = (blk.hi - blk.lo) * per
            if blk.src == rank:
                send[blk.dst] += n
            if blk.dst == rank:
                recv[blk.src] += n
        return send, recv

    def owned_output(self, rank: int) -> Tuple[int, int]:
        """Unpadded output planes [lo, hi) rank ``rank`` finalises (may be empty)."""
        b = self.grid.border[0]
        w = self.grid.spatial_shape[0]
        lo, hi = self.own[rank]
        return min(max(lo - b, 0), w), min(max(hi - b, 0), w)


def make_slab_plan(grid: PatchGrid, world: int) -> SlabPlan:
    if world < 1:
        raise ValueError("world must be >= 1")
    n = len(grid.locations)
    planes = grid.padded_shape[0]
    runs = [((r * n) // world, ((r + 1) * n) // world) for r in range(world)]
    own = [((r * planes) // world, ((r + 1) * planes) // world) for r in range(world)]
    return SlabPlan(grid, world, runs, own)


# ------------------------------------------------------------------------------------------------- device steps
class SlabOps:
    """The arithmetic of one rank, on whatever device the tensors live on."""

    def forward(self, volume: torch.Tensor, grid: PatchGrid, patches) -> torch.Tensor:
        """Extract + network forward of ``patches`` -> fp32 (len(patches), C_out, p0, p1, p2)."""
        raise NotImplementedError

    def out_channels(self, volume: torch.Tensor) -> int:
        raise NotImplementedError

    def stage(self, patch_out: torch.Tensor, lo: int, hi: int, dst: torch.Tensor) -> None:
        """dst[: C * (hi - lo) * p1 * p2] = patch_out[:, lo:hi] (dense)."""
        raise NotImplementedError

    def accumulate(self, acc: torch.Tensor, block: torch.Tensor, loc) -> None:
        """acc[:, loc[0]:loc[3], loc[1]:loc[4], loc[2]:loc[5]] += block  (block: (C, q, p1, p2))."""
        raise NotImplementedError

    def finalize(self, acc: torch.Tensor, grid: PatchGrid, plane0: int, out_lo: int, out_hi: int
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(probs (C, out_hi-out_lo, H, D) fp32, labels (out_hi-out_lo, H, D) uint8) of unpadded planes
        [out_lo, out_hi) from the accumulator whose first plane is padded plane ``plane0``."""
        raise NotImplementedError


class CudaSlabOps(SlabOps):
    """Production ops: libb200seg kernels + the native network plan."""

    def __init__(self, model, patch_batch_size: int = 16):
        self.model = model
        self.patch_batch_size = patch_batch_size

    def _compiled(self, volume):
        from .models import _engine
        precision = _engine._resolve_precision(self.model, volume)
        return _engine.compiled_for(self.model, precision, volume.device)

    def out_channels(self, volume):
        return self._compiled(volume).plan.out_channels

    def forward(self, volume, grid, patches):
        import b200seg
        compiled = self._compiled(volume)
        p0, p1, p2 = grid.patch_size
        out = torch.empty((len(patches), compiled.plan.out_channels, p0, p1, p2), dtype=torch.float32,
                          device=volume.device)
        for start in range(0, len(patches), self.patch_batch_size):
            locs = patches[start:start + self.patch_batch_size]
            buf = compiled.input_buffer(len(locs), p0, p1, p2)
            b200seg.grid_extract(volume, locs, grid.border, grid.pad_mode_code, grid.pad_value,
                                 buf.view(volume.shape[0]))
            compiled.run_blocked(len(locs), p0, p1, p2, out=out[start:start + len(locs)])
        return out

    def stage(self, patch_out, lo, hi, dst):
        import b200seg
        b200seg.copy_planes(patch_out, lo, hi, dst)

    def accumulate(self, acc, block, loc):
        import b200seg
        b200seg.overlap_add(acc, block[None], [loc])

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
def slab_predict(volume: torch.Tensor, grid: PatchGrid, ops: SlabOps, group=None, gather_probs: bool = False,
                 timings: Optional[dict] = None):
    """Runs this rank's share of the sliding window and returns (labels uint8 (W, H, D) on every rank,
    probs fp32 (C, W, H, D) or None).  ``volume`` is the full (C, W, H, D) volume on this rank's device.
    Results are bit-identical to the single-device ``PatchPredict.predict_volume``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = make_slab_plan(grid, world)
    device = volume.device
    channels = ops.out_channels(volume)
    per = plan.plane_elems(channels)
    p1, p2 = grid.patch_size[1], grid.patch_size[2]

    # ---- forwards of this rank's run of patches
    first, last = plan.runs[rank]
    patches = plan.patches_of(rank)
    y = ops.forward(volume, grid, patches) if patches else None

    # ---- stage every plane of every patch for its owner; one all-to-all moves the blocks
    blocks = plan.blocks()
    send_sizes, recv_sizes = plan.split_sizes(rank, channels)
    send = torch.empty(sum(send_sizes), dtype=torch.float32, device=device)
    recv = torch.empty(sum(recv_sizes), dtype=torch.float32, device=device)
    off = 0
    for blk in blocks:                       # ordered by (src, dst, patch): contiguous per destination
        if blk.src != rank:
            continue
        i0 = grid.locations[blk.patch][0]
        n = (blk.hi - blk.lo) * per
        ops.stage(y[blk.patch - first], blk.lo - i0, blk.hi - i0, send[off:off + n])
        off += n
    if world > 1:
        dist.all_to_all_single(recv, send, recv_sizes, send_sizes, group=group)
    else:
        recv = send
    y = None

    # ---- owner: add the blocks in sorted patch order (== the single-device per-voxel order)
    lo, hi = plan.own[rank]
    acc = torch.zeros((channels, hi - lo, *grid.padded_shape[1:]), dtype=torch.float32, device=device)
    mine, off = [], 0
    for blk in blocks:                       # (src, dst, patch) order == layout of ``recv`` (grouped by src)
        if blk.dst != rank:
            continue
        n = (blk.hi - blk.lo) * per
        mine.append((blk.patch, off, blk))
        off += n
    for _, o, blk in sorted(mine, key=lambda t: t[0]):
        loc = grid.locations[blk.patch]
        q = blk.hi - blk.lo
        block = recv[o:o + q * per].view(channels, q, p1, p2)
        ops.accumulate(acc, block, (blk.lo - lo, loc[1], loc[2], blk.hi - lo, loc[4], loc[5]))

    # ---- finalise the owned planes, gather the label slabs
    out_lo, out_hi = plan.owned_output(rank)
    w, h, d = grid.spatial_shape
    if out_hi > out_lo:
        probs, labels = ops.finalize(acc, grid, lo, out_lo, out_hi)
    else:
        probs = torch.empty((channels, 0, h, d), dtype=torch.float32, device=device)
        labels = torch.empty((0, h, d), dtype=torch.uint8, device=device)
    if world == 1:
        return labels, (probs if gather_probs else None)
    # equal-size all_gather: pad every slab to the largest one
    spans = [plan.owned_output(r) for r in range(world)]
    longest = max(b - a for a, b in spans)
    padded = torch.zeros((longest, h, d), dtype=torch.uint8, device=device)
    padded[:labels.shape[0]] = labels
    gathered = torch.empty((world * longest, h, d), dtype=torch.uint8, device=device)   # concatenation along dim 0
    dist.all_gather_into_tensor(gathered, padded, group=group)
    gathered = gathered.view(world, longest, h, d)
    full = torch.empty((w, h, d), dtype=torch.uint8, device=device)
    for r, (a, b) in enumerate(spans):
        full[a:b] = gathered[r, :b - a]
    full_probs = None
    if gather_probs:
        pp = torch.zeros((channels, longest, h, d), dtype=torch.float32, device=device)
        pp[:, :probs.shape[1]] = probs
        parts = torch.empty((world * channels, longest, h, d), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(parts, pp, group=group)
        parts = parts.view(world, channels, longest, h, d)
        full_probs = torch.empty((channels, w, h, d), dtype=torch.float32, device=device)
        for r, (a, b) in enumerate(spans):
            full_probs[:, a:b] = parts[r, :, :b - a]
    return full, full_probs
