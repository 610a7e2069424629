"""Drop-in for ``/root/reference/src/model/sinkhorn.py`` (``Sinkhorn``; same constructor and forward).

The reference forwards to ``pygmtools.sinkhorn(..., backend='pytorch')`` (``sinkhorn.py:85-87``), a python
loop over pairs and iterations.  Here one CTA per pair keeps the pair's matrix in shared memory for all
iterations (``csrc/sinkhorn.cu``).
"""
import torch
import torch.nn as nn
from torch import Tensor

from fpmatch import ops


class Sinkhorn(nn.Module):
    r"""
    Sinkhorn algorithm turns the input matrix into a bi-stochastic matrix (log-domain, temperature ``tau``).

    :param max_iter: maximum iterations (default: ``10``)
    :param tau: temperature (default: ``1``)
    :param epsilon: kept for signature compatibility (unused by the log-domain path, as in the reference)
    :param log_forward: ``True`` (default): log-domain kernel; ``False``: the reference's deprecated ``forward_ori``
     recurrence, evaluated with batched torch ops on the input's device (compatibility path, not a kernel)
    :param batched_operation: accepted and ignored - every pair is always processed concurrently and the
     result equals the reference's ``batched_operation=False`` arithmetic
    """
    def __init__(self, max_iter: int = 10, tau: float = 1., epsilon: float = 1e-4,
                 log_forward: bool = True, batched_operation: bool = False):
        super(Sinkhorn, self).__init__()
        self.max_iter = max_iter
        self.tau = tau
        self.epsilon = epsilon
        self.log_forward = log_forward
        if not log_forward:
            print('Warning: Sinkhorn algorithm without log forward is deprecated because log_forward is more stable.')
        self.batched_operation = batched_operation

    def forward(self, s: Tensor, nrows: Tensor = None, ncols: Tensor = None, dummy_row: bool = False) -> Tensor:
        if not self.log_forward:
            return self.forward_ori(s, nrows, ncols, dummy_row)
        return self.forward_log(s, nrows, ncols, dummy_row)

    def forward_ori(self, s, nrows=None, ncols=None, dummy_row=False):
        """The deprecated non-log recurrence of the reference (``sinkhorn.py:89-169``): per-pair softmax over the valid
        block at temperature ``tau``, dummy rows filled with ``epsilon``, ``+ epsilon``, then alternating row / column
        normalisation (even steps divide by the row sums) restricted to the valid block.  Same arithmetic with batched
        masks instead of the per-pair python loops and the ``b x n x n x n`` broadcast products; differentiable by
        torch autograd like the original."""
        if len(s.shape) == 2:
            s = s.unsqueeze(0)
            matrix_input = True
        elif len(s.shape) == 3:
            matrix_input = False
        else:
            raise ValueError('input data shape not understood.')
        B, R, C = s.shape
        dev = s.device
        as_t = lambda v, full: (torch.full((B,), full, dtype=torch.int64, device=dev) if v is None
                                else torch.as_tensor(v, dtype=torch.int64, device=dev))
        nrows, ncols = as_t(nrows, R), as_t(ncols, C)
        ri = torch.arange(R, device=dev).view(1, R, 1)
        ci = torch.arange(C, device=dev).view(1, 1, C)
        valid = (ri < nrows.view(B, 1, 1)) & (ci < ncols.view(B, 1, 1))
        sm = torch.softmax((s / self.tau).masked_fill(~valid, -float('inf')), dim=-1)
        s = torch.where(valid, sm, torch.zeros_like(sm))          # rows without a valid entry: softmax(-inf) = nan -> 0
        if dummy_row:
            pad = C - R
            s = torch.cat((s, torch.zeros((B, pad, C), dtype=s.dtype, device=dev)), dim=1)
            ori_nrows, nrows = nrows, ncols
            ri = torch.arange(s.shape[1], device=dev).view(1, -1, 1)
            dummy = (ri >= ori_nrows.view(B, 1, 1)) & (ri < nrows.view(B, 1, 1)) & (ci < ncols.view(B, 1, 1))
            s = torch.where(dummy, torch.full_like(s, self.epsilon), s)
            valid = (ri < nrows.view(B, 1, 1)) & (ci < ncols.view(B, 1, 1))
        s = s + self.epsilon
        vf = valid.to(s.dtype)
        for i in range(self.max_iter):
            if i % 2 == 0:
                tot = (s * (ci < ncols.view(B, 1, 1)).to(s.dtype)).sum(dim=2, keepdim=True)     # over the valid columns
            else:
                tot = (s * (ri < nrows.view(B, 1, 1)).to(s.dtype)).sum(dim=1, keepdim=True)     # over the valid rows
            s = s * torch.where(valid, 1 / tot.expand_as(s), torch.zeros_like(s)) * vf
        if dummy_row:
            if pad > 0:
                s = s[:, :-pad]
            keep = ~((ri[:, :s.shape[1]] >= ori_nrows.view(B, 1, 1)) & (ci < ncols.view(B, 1, 1)))
            s = s * keep.to(s.dtype)
        if matrix_input:
            s = s.squeeze(0)
        return s

    def forward_log(self, s, nrows=None, ncols=None, dummy_row=False):
        """Compute sinkhorn with row/column normalization in the log space."""
        if len(s.shape) == 2:
            s = s.unsqueeze(0)
            matrix_input = True
        elif len(s.shape) == 3:
            matrix_input = False
        else:
            raise ValueError('input data shape not understood.')
        nrows = nrows.to(s.device) if nrows is not None else None
        ncols = ncols.to(s.device) if ncols is not None else None
        if torch.is_grad_enabled() and s.requires_grad:
            # differentiable like the reference's pygmtools call (sinkhorn.py:85-87; used in training at gnn.py:219)
            from fpmatch import autograd as fa
            out = fa.SinkhornFn.apply(s.to(torch.float32), nrows, ncols, self.max_iter, self.tau, bool(dummy_row))
        else:
            x = s.detach().to(torch.float32).contiguous()
            out = ops.sinkhorn_log(x, nrows, ncols, self.max_iter, self.tau, dummy_row)
        if matrix_input:
            out = out.squeeze(0)
        return out


class GumbelSinkhorn(nn.Module):
    """
    Gumbel-Sinkhorn layer (``/root/reference/src/model/sinkhorn.py:172-233``): ``sample_num`` Gumbel-perturbed copies of
    every score matrix go through the log-domain Sinkhorn above as one batch of ``b * sample_num`` problems.  The noise
    is drawn with the reference's own expression from torch's generator, so a seeded run sees the same perturbation.

    :param max_iter: maximum iterations (default: ``10``)
    :param tau: temperature (default: ``1``)
    :param epsilon: kept for signature compatibility
    :param batched_operation: accepted and ignored (see ``Sinkhorn``)
    """
    def __init__(self, max_iter=10, tau=1., epsilon=1e-4, batched_operation=False):
        super(GumbelSinkhorn, self).__init__()
        self.sinkhorn = Sinkhorn(max_iter, tau, epsilon, batched_operation=batched_operation)

    def forward(self, s: Tensor, nrows: Tensor = None, ncols: Tensor = None, sample_num=5, dummy_row=False) -> Tensor:
        """:return: ``(b * sample_num, n1, n2)`` doubly-stochastic matrices, the samples of one input adjacent."""
        def sample_gumbel(t_like, eps=1e-20):
            u = torch.empty_like(t_like).uniform_()
            return -torch.log(-torch.log(u + eps) + eps)

        s_rep = torch.repeat_interleave(s, sample_num, dim=0)
        s_rep = s_rep + sample_gumbel(s_rep)
        nrows_rep = torch.repeat_interleave(nrows, sample_num, dim=0) if nrows is not None else None
        ncols_rep = torch.repeat_interleave(ncols, sample_num, dim=0) if ncols is not None else None
        return self.sinkhorn(s_rep, nrows_rep, ncols_rep, dummy_row)
