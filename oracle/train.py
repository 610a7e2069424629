"""CPU restatement of the reference's stage-1 training step.  TEST INFRASTRUCTURE ONLY (see package docstring).

Follows ``/root/reference/src/train/training_loop.py:22-64`` (zero_grad -> forward -> PermutationLoss (+ ks_loss +
cls_loss) -> backward -> clip_grad_norm_(5.0) in stage 1 -> AdamW.step) with the parameter groups of
``/root/reference/train.py:157-181`` (everything except the k-branch, the backbone and match_cls at LR, weight decay
1e-4; the backbone is outside the head and receives no update here because the step starts from its feature maps).
The forward is ``oracle.head.forward_head`` with the parameters left attached, so torch autograd differentiates the
restatement exactly as it differentiates the reference's python forward.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import head

Tensor = torch.Tensor


def permutation_loss(pred_dsmat: Tensor, gt_perm: Tensor, src_ns: Tensor, tgt_ns: Tensor,
                     dtype: torch.dtype = torch.float32) -> Tensor:
    """PermutationLoss.forward (/root/reference/src/loss_func.py:26-59): summed BCE over each pair's valid block,
    divided by the total number of source keypoints.  (``dtype`` = float64 only for the "truth" runs of the tests.)"""
    pred_dsmat = pred_dsmat.to(dtype)
    gt_perm = gt_perm.to(dtype)
    assert torch.all((pred_dsmat >= 0) * (pred_dsmat <= 1))
    loss = torch.tensor(0.0, dtype=dtype)
    n_sum = torch.zeros_like(loss)
    for b in range(pred_dsmat.shape[0]):
        r, c = int(src_ns[b]), int(tgt_ns[b])
        loss = loss + F.binary_cross_entropy(pred_dsmat[b, :r, :c], gt_perm[b, :r, :c], reduction="sum")
        n_sum = n_sum + float(src_ns[b])
    return loss / n_sum


K_PREFIXES = ("encoder_k.", "final_row.", "final_col.")
FROZEN_PREFIXES = K_PREFIXES + ("match_cls.", "node_layers.", "edge_layers.")


def trainable_names(state: Dict[str, Tensor], with_cls: bool = False, with_k: bool = False) -> List[str]:
    """Names the stage-1 optimizer updates (train.py:157-181), restricted to float parameters (not BN buffers).
    ``with_cls`` adds the match_cls parameters (train.py:239 gives them their own AdamW); ``with_k`` the AFA-U
    k-branch (encoder_k, final_row, final_col: trained from stage 2 on, train.py:183-215)."""
    frozen = tuple(q for q in FROZEN_PREFIXES if not (with_cls and q == "match_cls.") and not (with_k and q in K_PREFIXES))
    return [k for k, v in state.items() if v.is_floating_point() and not k.startswith(frozen)
            and "running_" not in k and "num_batches" not in k]


def loss_and_grads(state: Dict[str, Tensor], data: dict, fmaps, fmap_grads: bool = False,
                   dtype: torch.dtype = torch.float32, regression: bool = False):
    """One forward/backward of the stage-1 objective (PermutationLoss, + cls_loss when the batch carries labels);
    returns (loss, {name: grad}, [fmap grads], forward outputs).  ``dtype=torch.float64`` evaluates the same
    formulae in double precision: the tests use it to bound the fp32 oracle's own rounding noise."""
    p = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in state.items()}
    names = trainable_names(p, with_cls="label" in data, with_k=regression)
    for k in names:
        p[k].requires_grad_(True)
    fm = [(a.clone().to(dtype).requires_grad_(fmap_grads), b.clone().to(dtype).requires_grad_(fmap_grads))
          for a, b in fmaps]
    out = head.forward_head(p, data, fm, regression=regression, training=True, keep_graph=True, dtype=dtype)
    loss = permutation_loss(out["ds_mat"], data["gt_perm_mat"], data["ns"][0], data["ns"][1], dtype)
    # stage objective of training_loop.py:48-50: primary loss + ks_loss + cls_loss
    total = loss + (out["cls_loss"] if "cls_loss" in out else 0.0) + (out["ks_loss"] if regression else 0.0)
    total.backward()
    grads = {k: p[k].grad for k in names if p[k].grad is not None}
    fg = [(a.grad, b.grad) for a, b in fm] if fmap_grads else None
    return loss.detach(), grads, fg, out


def warmup_lr(t: int, base_lr: float = 1e-3, warmup_epochs: int = 10, steps_per_epoch: int = 75) -> float:
    """Learning rate of training step t under the reference's WarmupScheduler (utils/scheduler.py:11-13, stepped once
    per epoch, train.py:246-252,297) with 3 x num_iterations = 75 steps per epoch (training_loop.py:21-22)."""
    epoch = t // steps_per_epoch
    return base_lr * float(epoch + 1) / warmup_epochs if epoch < warmup_epochs else base_lr


def train_trajectory(state: Dict[str, Tensor], batches: List[dict], steps: int, lr: float = 1e-3,
                     weight_decay: float = 1e-4, clip: float = 5.0, lr_schedule=None) -> List[float]:
    """`steps` AdamW steps over a fixed cycle of batches; returns the per-step PermutationLoss values.
    ``lr_schedule``: optional callable step -> learning rate (e.g. ``warmup_lr``)."""
    p = {k: v.clone() for k, v in state.items()}
    names = trainable_names(p)
    params = [p[k].requires_grad_(True) for k in names]
    opt = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay)
    losses = []
    for it in range(steps):
        data = batches[it % len(batches)]
        if lr_schedule is not None:
            for gp in opt.param_groups:
                gp["lr"] = lr_schedule(it)
        opt.zero_grad()
        from fpmatch import synth          # only for clone_batch (pure python container copy)
        out = head.forward_head(p, synth.clone_batch(data), data["fmaps"], regression=False, training=True,
                                keep_graph=True)
        loss = permutation_loss(out["ds_mat"], data["gt_perm_mat"], data["ns"][0], data["ns"][1])
        total = loss + (out["cls_loss"] if "cls_loss" in out else 0.0)
        total.backward()
        torch.nn.utils.clip_grad_norm_([q for q in params if q.grad is not None], max_norm=clip)
        opt.step()
        losses.append(float(loss.detach()))
    return losses
