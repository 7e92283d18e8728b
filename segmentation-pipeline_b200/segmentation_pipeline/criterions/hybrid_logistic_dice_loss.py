"""HybridLogisticDiceLoss -- same constructor, attributes and returned dict as the reference's
criterions/hybrid_logistic_dice_loss.py:6-43, computed by libb200seg: one fused pass over prediction and target for the
four per-(n, c) sums the loss is made of (``b200seg_hybrid_loss_forward``) instead of five full-tensor passes with
temporaries, and one elementwise kernel for d loss / d prediction (``b200seg_hybrid_loss_backward``), wired into
autograd so that the reference trainer's ``loss_dict['loss'].backward()`` (segmentation_trainer.py:173-177) works.
``dice_loss`` and ``logistic_loss`` are reported values (no gradient), as the trainer only differentiates ``loss``.
CUDA tensors only: there is no CPU fallback."""
from __future__ import annotations

import torch
from torch import nn


class _HybridLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, dice_weight, class_weights, square_dice):
        import b200seg
        with b200seg.on_device(prediction):
            p = prediction.detach().to(torch.float32).contiguous()
            t = target.detach().to(torch.float32).contiguous()
            out3, sums = b200seg.hybrid_loss_forward(p, t, dice_weight, class_weights, square_dice)
        ctx.save_for_backward(p, t, sums)
        ctx.cfg = (dice_weight, class_weights, square_dice)
        ctx.mark_non_differentiable(out3)
        return out3[0].clone(), out3

    @staticmethod
    def backward(ctx, grad_loss, _grad_out3):
        import b200seg
        p, t, sums = ctx.saved_tensors
        dice_weight, class_weights, square_dice = ctx.cfg
        with b200seg.on_device(p):
            g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
            grad = b200seg.hybrid_loss_backward(p, t, sums, dice_weight, class_weights, square_dice, g)
        return grad, None, None, None, None


class HybridLogisticDiceLoss(nn.Module):
    def __init__(self, dice_weight=0.5, logistic_class_weights=None, square_dice=True):
        super().__init__()
        self.dice_weight = dice_weight
        self.logistic_class_weights = logistic_class_weights
        self.square_dice = square_dice

    def forward(self, prediction, target):
        if not (prediction.is_cuda and target.is_cuda):
            raise RuntimeError("HybridLogisticDiceLoss (b200) runs on CUDA tensors only (no CPU fallback)")
        if prediction.dim() != 5 or prediction.shape != target.shape:
            raise RuntimeError(f"expected prediction and target of one shape (N, C, W, H, D), got "
                               f"{tuple(prediction.shape)} and {tuple(target.shape)}")
        weights = None
        if self.logistic_class_weights is not None:
            weights = torch.tensor(self.logistic_class_weights, dtype=torch.float32, device=prediction.device).contiguous()
            if weights.numel() != prediction.shape[1]:
                raise RuntimeError("logistic_class_weights must have one entry per channel")
        loss, out3 = _HybridLoss.apply(prediction, target, float(self.dice_weight), weights, bool(self.square_dice))
        return {'loss': loss, "dice_loss": out3[1], "logistic_loss": out3[2]}
