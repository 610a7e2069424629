"""Drop-in for the loss every train script of the reference uses: ``PermutationLoss``
(``/root/reference/src/loss_func.py:8-59``; ``train.py:143``, ``train_new.py``, ``train_single_image.py``).

The reference loops over pairs in python (``binary_cross_entropy(reduction='sum')`` per valid block); here the whole
batch is one launch forward and one backward (``csrc/loss.cu``).  The eight other ThinkMatch losses of the reference
file are imported by nothing and are not rebuilt.
"""
import torch
import torch.nn as nn
from torch import Tensor

from fpmatch import ops


class _PermutationLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, n1, n2):
        pair = ops.permutation_loss_pairs(pred, gt, n1, n2)
        n_sum = n1.sum().to(torch.float32)
        ctx.save_for_backward(pred, gt, n1, n2, n_sum)
        return pair.sum() / n_sum

    @staticmethod
    def backward(ctx, g):
        pred, gt, n1, n2, n_sum = ctx.saved_tensors
        scale = (g / n_sum).reshape(1).to(torch.float32).contiguous()
        return ops.permutation_loss_bwd(pred, gt, n1, n2, scale), None, None, None


class PermutationLoss(nn.Module):
    r"""Binary cross entropy between a doubly-stochastic prediction and the ground-truth permutation, summed over each
    pair's valid :math:`n_1 \times n_2` block and divided by :math:`\sum_b n_{1,b}` (loss_func.py:26-59)."""

    check_range = True     # False: skip the reference's 0 <= pred <= 1 assertion (a host sync; illegal while a CUDA graph is captured)

    def __init__(self):
        super(PermutationLoss, self).__init__()

    def forward(self, pred_dsmat: Tensor, gt_perm: Tensor, src_ns: Tensor, tgt_ns: Tensor) -> Tensor:
        pred = pred_dsmat.to(dtype=torch.float32)
        gt = gt_perm.to(pred.device, torch.float32)
        if self.check_range:
            lo, hi = torch.aminmax(pred.detach())             # the reference asserts 0 <= pred <= 1 (one sync, as there)
            if not (lo >= 0 and hi <= 1 and bool(((gt >= 0) & (gt <= 1)).all())):
                print(pred_dsmat)
                raise AssertionError("pred_dsmat and gt_perm must lie in [0, 1]")
        n1 = src_ns.to(pred.device, torch.int64).contiguous()
        n2 = tgt_ns.to(pred.device, torch.int64).contiguous()
        return _PermutationLossFn.apply(pred.contiguous(), gt.contiguous(), n1, n2)
