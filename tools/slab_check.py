#!/usr/bin/env python
"""Multi-GPU check of the z-slab mode (run under torchrun, NCCL): every rank's gathered label map must agree with
the single-GPU PatchPredict result (>= 99.9 % of voxels; probabilities within fp32 re-association)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from segmentation_pipeline.distributed import CudaSlabOps, slab_predict  # noqa: E402
from segmentation_pipeline.grid import PatchGrid  # noqa: E402
from segmentation_pipeline.models import set_precision  # noqa: E402
from segmentation_pipeline.prediction import PatchPredict  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)
    dist.init_process_group("nccl", device_id=device)
    set_precision("bf16")
    model = bench.build_model().to(device)
    vol = bench.synthetic_volume(0).to(device)
    grid = PatchGrid(vol.shape[1:], 96, 48, "edge")
    with torch.no_grad():
        for _ in range(2):
            torch.cuda.synchronize(); dist.barrier()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            labels, probs = slab_predict(vol, grid, CudaSlabOps(model, 24), gather_probs=True)
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
        ref_probs, ref_labels = PatchPredict(patch_batch_size=24, patch_size=96, patch_overlap=48,
                                             padding_mode="edge").predict_volume(model, vol)
    agree = (labels == ref_labels).float().mean().item()
    err = ((probs - ref_probs).abs().max() / ref_probs.abs().max()).item()
    print(f"rank {rank}/{dist.get_world_size()}: slab mode {ms:.1f} ms, label agreement {agree:.6f}, prob rel err {err:.2e}")
    assert agree >= 0.999 and err < 1e-5
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
