"""Drop-in for ``/root/reference/src/model/soft_topk.py`` (``soft_topk``, ``greedy_perm``, ``Sinkhorn_m``).

The reference loops over pairs in python, and its ``greedy_perm`` issues two reductions and one host sync
per candidate (``soft_topk.py:56-77``).  Here ``soft_topk`` is one launch with one CTA per pair
(``csrc/sinkhorn.cu``) and the greedy selection runs on the device (``csrc/lap.cu``).

Tie order: the reference ranks candidates with an unstable ``torch.argsort``; this implementation fixes
"value descending, flat index ascending" (= ``stable=True``), see SURVEY.md A.7.
"""
import torch
import torch.nn as nn
from torch import Tensor

from fpmatch import ops


def soft_topk(scores, ks, max_iter=10, tau=1., nrows=None, ncols=None, return_prob=False):
    r"""
    Topk-GM algorithm to suppress matches containing outliers.

    :param scores: :math:`(b\times n_1 \times n_2)` input 3d tensor
    :param ks: :math:`(b)` number of matches of each graph pair
    :param max_iter: maximum iterations (default: ``10``)
    :param tau: Sinkhorn temperature (default: ``1``)
    :param nrows: :math:`(b)` number of objects in dim1
    :param ncols: :math:`(b)` number of objects in dim2
    :param return_prob: whether to also return the soft matrix
    :return: the hard top-k matrix; with ``return_prob=True`` also the soft matrix
    """
    x = scores.detach().to(torch.float32).contiguous()
    dev = x.device
    differentiable = torch.is_grad_enabled() and scores.requires_grad
    B, R, C = x.shape
    if nrows is None:
        nrows = torch.full((B,), R, dtype=torch.int64, device=dev)
    if ncols is None:
        ncols = torch.full((B,), C, dtype=torch.int64, device=dev)
    nrows, ncols = nrows.to(dev), ncols.to(dev)
    ks = torch.as_tensor(ks, dtype=torch.float32, device=dev).reshape(-1)
    if differentiable:        # the soft matrix carries gradients back to `scores`, as the reference's does
        from fpmatch import autograd as fa
        output_s = fa.SoftTopkFn.apply(scores.to(torch.float32), ks, nrows, ncols, max_iter, tau)
    else:
        output_s = ops.soft_topk(x, ks, nrows, ncols, max_iter, tau)
    # hard selection: candidates ranked over the flattened (n1_b x n2_b block of the) soft matrix
    top_indices = _rank_valid_block(output_s.detach(), nrows, ncols)
    hard = ops.greedy_perm(torch.zeros_like(x), top_indices, ks)
    if return_prob:
        return hard, output_s
    return hard


def _rank_valid_block(output_s: Tensor, nrows: Tensor, ncols: Tensor) -> Tensor:
    """``argsort(output[:, :, 1], descending)`` of soft_topk.py:43.  The reference sorts a
    [b, n1max*n2max] buffer whose first n1_b*n2_b entries hold the pair's block in row-major order of
    width n2_b (the rest is exp(-inf) = 0), while greedy_perm decodes indices with the PADDED width;
    that index space is reproduced as is."""
    B, R, C = output_s.shape
    dev = output_s.device
    q = torch.arange(R * C, device=dev).view(1, -1).expand(B, -1)
    nb = ncols.view(B, 1).clamp(min=1)
    i = torch.div(q, nb, rounding_mode='floor').clamp(max=R - 1)
    j = q % nb
    valid = q < (nrows * ncols).view(B, 1)
    flat = torch.gather(output_s.reshape(B, -1), 1, i * C + j) * valid.to(output_s.dtype)
    return torch.argsort(flat, descending=True, dim=-1, stable=True)


def greedy_perm(x, top_indices, ks):
    r"""
    Greedy-topk algorithm to select matches with topk confidences (in place on ``x``, as the reference).

    :param x: :math:`(b\times n_1 \times n_2)` input 3d tensor
    :param top_indices: indices of topk matches
    :param ks: :math:`(b)` number of matches of each graph pair
    """
    if not x.is_contiguous() or x.dtype != torch.float32:
        raise RuntimeError('greedy_perm expects a contiguous float32 tensor')
    ks = torch.as_tensor(ks, dtype=torch.float32, device=x.device).reshape(-1)
    return ops.greedy_perm(x, top_indices.to(x.device), ks)


class Sinkhorn_m(nn.Module):
    r"""
    Sinkhorn with marginal distributions over a list of ``[n1_b*n2_b, 2]`` distance matrices
    (``soft_topk.py:80-255``).  Kept for API compatibility; ``forward`` evaluates the same recurrence with
    torch ops on whatever device the inputs live on (it is not on ``Net.forward``'s path - the fused
    ``soft_topk`` above is).
    """

    def __init__(self, max_iter: int = 10, tau: float = 1., epsilon: float = 1e-4,
                 log_forward: bool = True, batched_operation: bool = False):
        super(Sinkhorn_m, self).__init__()
        self.max_iter = max_iter
        self.tau = tau
        self.epsilon = epsilon
        self.log_forward = log_forward
        self.batched_operation = batched_operation

    def forward(self, s, row_prob: Tensor, col_prob: Tensor, nrows: Tensor = None, ncols: Tensor = None,
                dummy_row: bool = False) -> Tensor:
        if not self.log_forward:
            raise NotImplementedError
        return self.forward_log(s, row_prob, col_prob, nrows, ncols, dummy_row)

    def forward_log(self, s, row_prob, col_prob, nrows=None, ncols=None, dummy_row=True):
        batch_size = len(s)
        s = [s[i] / self.tau for i in range(batch_size)]
        log_row_prob = torch.log(row_prob).unsqueeze(2)
        log_col_prob = torch.log(col_prob).unsqueeze(1)
        ret = torch.full((batch_size, int(nrows.max() * ncols.max()), 2), -float('inf'),
                         device=s[0].device, dtype=s[0].dtype)
        for b in range(batch_size):
            log_s = s[b]
            n = int(nrows[b] * ncols[b])
            step = 0
            while step < self.max_iter or bool(torch.any(log_s > 0)):
                if step % 2 == 0:
                    log_s = log_s - torch.logsumexp(log_s, 1, keepdim=True) + log_row_prob[b, 0:n]
                else:
                    log_s = log_s - torch.logsumexp(log_s, 0, keepdim=True) + log_col_prob[b]
                log_s = torch.where(torch.isnan(log_s), torch.full_like(log_s, -float('inf')), log_s)
                step += 1
            ret[b, 0:n] = log_s
        return torch.exp(ret)
