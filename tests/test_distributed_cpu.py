"""World-size-2 (and 3) gloo tests of the multi-GPU host logic on the CPU: the z-slab plan, the forward-only halo
exchange, the label all-gather and the cohort confusion all-reduce.  The per-rank arithmetic is injected through
``SlabOps``; here it is the CPU oracle (numpy), in production it is libb200seg (tests/test_gpu_models.py covers that
path with one rank on a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import evalstats, grid as ogrid  # noqa: E402
from segmentation_pipeline.distributed import (SlabOps, all_reduce_confusion, make_slab_plan, shard_subjects,  # noqa: E402
                                               slab_predict)
from segmentation_pipeline.grid import PatchGrid  # noqa: E402


def toy_model(patches: np.ndarray) -> np.ndarray:
    """A deterministic 'network': 3 class scores per voxel that depend on the patch content only."""
    x = patches.sum(axis=1, keepdims=True)
    logits = np.concatenate([x, -x, 0.3 * np.ones_like(x)], axis=1)
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


class OracleSlabOps(SlabOps):
    def out_channels(self, volume):
        return 3

    def forward(self, volume, grid, patches):
        padded = ogrid.pad_volume(volume.numpy(), grid.patch_overlap, grid.padding_mode)
        return torch.from_numpy(toy_model(ogrid.extract_patches(padded, np.array(patches))))

    def stage(self, patch_out, lo, hi, dst):
        block = patch_out[:, lo:hi].contiguous()
        dst[:block.numel()] = block.reshape(-1)

    def accumulate(self, acc, block, loc):
        acc[:, loc[0]:loc[3], loc[1]:loc[4], loc[2]:loc[5]] += block

    def finalize(self, acc, grid, plane0, out_lo, out_hi):
        cw, ch, cd = (np.array(c, np.float32) for c in grid.axis_counts())
        b = grid.border
        lo = out_lo + b[0] - plane0
        n = out_hi - out_lo
        h, d = grid.spatial_shape[1:]
        sub = acc.numpy()[:, lo:lo + n, b[1]:b[1] + h, b[2]:b[2] + d]
        cnt = cw[out_lo + b[0]:out_hi + b[0], None, None] * ch[None, b[1]:b[1] + h, None] * cd[None, None, b[2]:b[2] + d]
        probs = sub / cnt
        labels = evalstats.argmax_labels(probs)[0].astype(np.uint8)
        return torch.from_numpy(probs), torch.from_numpy(labels)


def _single_process_reference(vol, patch, overlap, padding):
    return ogrid.sliding_window(vol, toy_model, patch, overlap, padding, "average", patch_batch_size=4)


def _worker(rank, world, port, vol, patch, overlap, padding, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = PatchGrid(vol.shape[1:], patch, overlap, padding)
        labels, probs = slab_predict(torch.from_numpy(vol), grid, OracleSlabOps(), gather_probs=True)
        # cohort mode: every rank evaluates its subjects, confusion matrices are summed
        cm = torch.zeros((3, 3), dtype=torch.int64)
        for i in shard_subjects(5, rank, world):
            rng = np.random.default_rng(100 + i)
            p, t = rng.integers(0, 3, 500), rng.integers(0, 3, 500)
            cm += torch.from_numpy(evalstats.confusion_matrix(p, t, 3))
        all_reduce_confusion(cm)
        ret[rank] = (labels.numpy(), probs.numpy(), cm.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,patch,overlap,padding", [
    (2, (2, 24, 14, 12), (8, 8, 6), (4, 2, 2), None),
    (2, (1, 20, 10, 10), (8, 8, 8), (4, 4, 4), "edge"),
    (3, (1, 30, 9, 9), (8, 8, 8), (2, 2, 2), None),      # slabs thinner than a patch: a patch feeds three owners
    (3, (1, 12, 20, 20), (8, 8, 8), (4, 4, 4), None),    # 2 start planes < 3 ranks: the run split still balances
])
def test_slab_predict_matches_single_process(world, shape, patch, overlap, padding):
    rng = np.random.default_rng(7)
    vol = rng.standard_normal(shape).astype(np.float32)
    ref_probs = _single_process_reference(vol, patch, overlap, padding)
    ref_labels = evalstats.argmax_labels(ref_probs)[0].astype(np.uint8)
    manager = mp.Manager()
    ret = manager.dict()
    port = 29600 + (os.getpid() % 200) + world
    mp.spawn(_worker, args=(world, port, vol, patch, overlap, padding, ret), nprocs=world, join=True)
    cm_ref = np.zeros((3, 3), np.int64)
    for i in range(5):
        r = np.random.default_rng(100 + i)
        p, t = r.integers(0, 3, 500), r.integers(0, 3, 500)
        cm_ref += evalstats.confusion_matrix(p, t, 3)
    for rank in range(world):
        labels, probs, cm = ret[rank]
        assert labels.shape == ref_labels.shape
        np.testing.assert_array_equal(probs, ref_probs)     # raw blocks added in sorted patch order: BIT-exact
        np.testing.assert_array_equal(labels, ref_labels)
        np.testing.assert_array_equal(cm, cm_ref)                           # integer counts: bit-exact


def test_slab_plan_covers_everything_once():
    for shape, n_patches in (((256, 256, 192), 144), ((224, 224, 224), 125)):
        grid = PatchGrid(shape, 96, 48, "edge")
        assert len(grid.locations) == n_patches
        for world in (1, 2, 3, 4, 8):
            plan = make_slab_plan(grid, world)
            seen = []
            for r in range(world):
                seen += plan.patches_of(r)
            assert seen == grid.locations                                     # contiguous runs of the sorted list
            sizes = [b - a for a, b in plan.runs]
            assert max(sizes) - min(sizes) <= 1                               # every rank works (config 3 at 8: 15 / 16)
            assert plan.own[0][0] == 0 and plan.own[-1][1] == grid.padded_shape[0]
            for (a0, a1), (b0, b1) in zip(plan.own, plan.own[1:]):
                assert a1 == b0
            assert sum(b - a for a, b in (plan.owned_output(r) for r in range(world))) == shape[0]
            # every plane of every patch goes to exactly one owner
            blocks = plan.blocks()
            per_patch = {}
            for blk in blocks:
                assert blk.src == plan.rank_of_patch(blk.patch)
                assert plan.own[blk.dst][0] <= blk.lo < blk.hi <= plan.own[blk.dst][1]
                per_patch.setdefault(blk.patch, []).append((blk.lo, blk.hi))
            for pidx, spans in per_patch.items():
                spans.sort()
                assert spans[0][0] == grid.locations[pidx][0] and spans[-1][1] == grid.locations[pidx][3]
                for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                    assert a1 == b0
            # the two sides of the all-to-all agree
            for r in range(world):
                send, _ = plan.split_sizes(r, 10)
                for d in range(world):
                    assert send[d] == plan.split_sizes(d, 10)[1][r]


def test_shard_subjects_round_robin():
    assert shard_subjects(64, 3, 8) == list(range(3, 64, 8))
    assert sorted(i for r in range(8) for i in shard_subjects(10, r, 8)) == list(range(10))
