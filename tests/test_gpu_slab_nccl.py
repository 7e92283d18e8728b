"""z-slab mode over NCCL on real GPUs (needs >= 2 devices on the box; skipped, saying so, on a single-GPU box): two
ranks partition one BASELINE-config-3 volume; labels and probabilities must be bit-identical to one GPU.  The host
logic of the same code is covered on the CPU (gloo, world size 2 / 3) by tests/test_distributed_cpu.py, and the
one-rank CUDA path by tests/test_gpu_models.py::test_slab_mode_single_rank_equals_patch_predict."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_nccl_slab_is_bit_identical_to_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip(f"{torch.cuda.device_count()} GPU on this box: the 2-rank NCCL run needs two")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29717", os.path.join(ROOT, "tools", "slab_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("SLAB ")]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads(line[0][5:])
    print(line[0])
    assert res["labels_bit_identical_to_1gpu"] and res["probs_bit_identical_to_1gpu"]
    assert res["patches_per_rank"] == [62, 63] and res["speedup"] > 1.5


@pytest.mark.gpu
def test_two_rank_ddp_training_step_all_reduces_the_device_gradients():
    """BASELINE config 5: the device training step under DistributedDataParallel (NCCL gradient all-reduce)."""
    if torch.cuda.device_count() < 2:
        pytest.skip(f"{torch.cuda.device_count()} GPU on this box: the 2-rank NCCL run needs two")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29719", os.path.join(ROOT, "tools", "train_ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("TRAINDDP ")]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads(line[0][9:])
    print(line[0])
    assert res["max_rel_err_vs_mean_of_local_grads"] <= 1e-5 and res["max_weight_drift_across_ranks"] == 0.0
    assert res["losses"][-1] < res["losses"][0]
