"""The reference drivers' criterion as two kernels.

``run.py:382`` (and ``run_imageclef.py`` / ``run_mimic.py``) build ``nn.MultiLabelSoftMarginLoss()`` and call it on
``(logits, soft targets)`` at ``run.py:431``.  ATen evaluates it as ~20 pointwise / reduce launches forward plus backward over
the (B, 3000) logits; here the forward is one streaming reduction and the backward one pointwise pass
(``csrc/train_step.cu``).  Same constructor arguments and call signature as the torch module, so a driver swaps
``nn.MultiLabelSoftMarginLoss`` for ``vqa_b200.loss.MultiLabelSoftMarginLoss`` and nothing else.  CUDA only, like the
rest of the package.
"""
from __future__ import annotations

import torch

from . import kernels as kn


class _MLSMLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, scale):
        logits = logits.contiguous()
        target = target.contiguous()
        ctx.save_for_backward(logits, target)
        ctx.scale = scale
        return kn.mlsm_loss_fwd(logits, target, scale)

    @staticmethod
    def backward(ctx, grad_out):
        logits, target = ctx.saved_tensors
        return kn.mlsm_loss_bwd(logits, target, grad_out.contiguous(), ctx.scale), None, None


class MultiLabelSoftMarginLoss(torch.nn.Module):
    """``loss = mean_b mean_a -( y log sigmoid(x) + (1 - y) log sigmoid(-x) )`` (``reduction='mean'``, the reference's use) or the
    sum over the batch of the per-sample class means (``'sum'``).  Per-class ``weight`` and ``reduction='none'`` are not
    part of the reference's path and raise."""

    def __init__(self, weight=None, size_average=None, reduce=None, reduction: str = "mean"):
        super().__init__()
        if weight is not None or size_average is not None or reduce is not None:
            raise NotImplementedError("vqa_b200 MultiLabelSoftMarginLoss: weight / legacy size_average / reduce are not supported")
        if reduction not in ("mean", "sum"):
            raise NotImplementedError(f"vqa_b200 MultiLabelSoftMarginLoss: reduction={reduction!r} is not supported (mean, sum)")
        self.reduction = reduction

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if input.dim() != 2 or input.shape != target.shape:
            raise RuntimeError(f"MultiLabelSoftMarginLoss: expected (B, A) logits and targets of one shape, got {tuple(input.shape)} "
                               f"and {tuple(target.shape)}")
        B, A = input.shape
        scale = 1.0 / (A * B) if self.reduction == "mean" else 1.0 / A
        return _MLSMLossFn.apply(input, target, scale)
