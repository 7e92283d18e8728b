#!/usr/bin/env python
"""Multi-GPU check of the training step (run under torchrun, NCCL): the device autograd Function under
DistributedDataParallel.  Every rank trains a copy of one small msseg2-style ModularUNet on its own batch; the
all-reduced gradients must equal the mean of the per-rank local gradients, and the weights must stay identical on all
ranks after the optimizer step (BASELINE config 5: data-parallel training with NCCL gradient all-reduce)."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from segmentation_pipeline import models as M  # noqa: E402
from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)                                    # same initial weights on every rank
    net = M.ModularUNet(2, 2, [8, 16, 16], 3, block_params={"residual": True}, downsample_class=M.BlurConv3d,
                        downsample_params={"kernel_size": 3, "stride": 2, "padding": 1},
                        upsample_class=M.BlurConvTranspose3d,
                        upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0}).to(device).train()
    local_net = copy.deepcopy(net)
    g = torch.Generator().manual_seed(10 + rank)            # different data per rank
    x = torch.randn(2, 2, 32, 32, 32, generator=g).to(device)
    y = torch.nn.functional.one_hot((torch.rand(2, 32, 32, 32, generator=g) < 0.2).long(), 2).movedim(-1, 1).float().to(device)
    criterion = HybridLogisticDiceLoss(logistic_class_weights=[1, 100])

    criterion(local_net(x), y)["loss"].backward()
    names = [n for n, p in local_net.named_parameters() if p.grad is not None]
    local_grads = {n: p.grad.clone() for n, p in local_net.named_parameters() if p.grad is not None}

    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9)
    losses = []
    worst = 0.0
    for it in range(3):
        loss = criterion(ddp(x), y)["loss"]
        opt.zero_grad()
        loss.backward()
        if it == 0:
            for n, p in net.named_parameters():
                if n not in local_grads:
                    continue
                mean = local_grads[n].clone()
                dist.all_reduce(mean, op=dist.ReduceOp.SUM)
                mean /= world
                scale = float(mean.abs().max()) + 1e-12
                worst = max(worst, float((p.grad - mean).abs().max()) / scale)
        opt.step()
        losses.append(float(loss.detach()))
    # weights identical on all ranks
    drift = 0.0
    for p in net.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        drift = max(drift, float((p.detach() - ref).abs().max()))
    res = {"world": world, "params_with_grad": len(names), "max_rel_err_vs_mean_of_local_grads": worst,
           "max_weight_drift_across_ranks": drift, "losses": losses}
    if rank == 0:
        print("TRAINDDP " + json.dumps(res))
    assert worst <= 1e-5 and drift == 0.0
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
