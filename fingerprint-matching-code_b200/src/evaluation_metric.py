"""Drop-in for the matching metrics the reference's train / eval loops call after every forward
(``/root/reference/src/evaluation_metric.py:58-131,200-222``: ``matching_recall``, ``matching_precision``,
``matching_accuracy``).  One launch computes the per-pair sums the reference gathers in a python loop
(``csrc/loss.cu::matching_stats_kernel``).  The pck / clustering metrics and ``generate_roc_curve`` of the reference
file are not on the matching head's path and are not rebuilt.
"""
import torch
from torch import Tensor

from fpmatch import ops


def _check(pmat_pred: Tensor, pmat_gt: Tensor):
    assert torch.all((pmat_pred == 0) + (pmat_pred == 1)), 'pmat_pred can only contain 0/1 elements.'
    assert torch.all((pmat_gt == 0) + (pmat_gt == 1)), 'pmat_gt should only contain 0/1 elements.'
    assert torch.all(torch.sum(pmat_pred, dim=-1) <= 1) and torch.all(torch.sum(pmat_pred, dim=-2) <= 1)
    assert torch.all(torch.sum(pmat_gt, dim=-1) <= 1) and torch.all(torch.sum(pmat_gt, dim=-2) <= 1)


def _stats(pmat_pred, pmat_gt, ns):
    dev = pmat_pred.device
    pmat_gt = pmat_gt.to(dev)
    _check(pmat_pred, pmat_gt)
    return ops.matching_stats(pmat_pred.to(torch.float32).contiguous(), pmat_gt.to(torch.float32).contiguous(),
                              ns.to(dev, torch.int64).contiguous())


def matching_recall(pmat_pred: Tensor, pmat_gt: Tensor, ns: Tensor) -> Tensor:
    r""":math:`tr(X {X^{gt}}^\top) / \sum X^{gt}` per pair over the first ``ns[b]`` rows; 1 where the pair has no
    ground-truth match (evaluation_metric.py:58-90)."""
    st = _stats(pmat_pred, pmat_gt, ns)
    acc = st[:, 0] / st[:, 1]
    acc[torch.isnan(acc)] = 1
    return acc


def matching_precision(pmat_pred: Tensor, pmat_gt: Tensor, ns: Tensor) -> Tensor:
    r""":math:`tr(X {X^{gt}}^\top) / \sum X` per pair; 1 where nothing was predicted - 0/0 = NaN -> 1 as the reference
    sets it (evaluation_metric.py:93-125; only its ``*_varied`` variants use 0)."""
    st = _stats(pmat_pred, pmat_gt, ns)
    precision = st[:, 0] / st[:, 2]
    precision[torch.isnan(precision)] = 1
    return precision


def matching_accuracy(pmat_pred: Tensor, pmat_gt: Tensor, ns, idx: int) -> Tensor:
    """Wrapper of ``matching_recall`` with ``ns[idx]`` (evaluation_metric.py:200-222)."""
    return matching_recall(pmat_pred, pmat_gt, ns[idx])
