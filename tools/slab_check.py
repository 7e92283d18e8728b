#!/usr/bin/env python
"""Multi-GPU check of the z-slab mode (run under torchrun, NCCL) on BASELINE config 3: the gathered label map and the
probabilities of every rank must be BIT-identical to the single-GPU PatchPredict result; prints the strong-scaling
numbers bench.py reports under "slab"."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from segmentation_pipeline.models import set_precision  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    set_precision("bf16")
    res = bench.run_slab_section(world, rank, device)
    if rank == 0:
        print("SLAB " + json.dumps(res))
        assert res["labels_bit_identical_to_1gpu"] and res["probs_bit_identical_to_1gpu"]
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
