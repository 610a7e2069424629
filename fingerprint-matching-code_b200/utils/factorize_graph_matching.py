"""Drop-in for the live entry points of ``/root/reference/utils/factorize_graph_matching.py``.

``construct_sparse_aff_mat`` (:57-95) is kept verbatim in behaviour for callers that want the explicit
index lists; ``Net.forward`` does not call it - the association graph stays factorised on the GPU
(``csrc/gnn.cu``).  ``kronecker_torch`` / ``kronecker_sparse`` (:98-137) are host/data-pipeline helpers.
The dense ``construct_aff_mat`` / ``RebuildFGM`` path (:10-54,140-186) is dormant in the reference
(``ngm.py:294-315`` is commented out); it is rebuilt here as a scatter / gather pair over the Kronecker columns.
"""
import numpy as np
import scipy.sparse as ssp
import torch
from torch import Tensor


def construct_sparse_aff_mat(Ke: Tensor, Kp: Tensor, row_idx: Tensor, col_idx: Tensor):
    r"""
    Values and indices of the sparse affinity matrix
    :math:`K = diag(vec(K_p)) + (G_2 \otimes G_1) diag(vec(K_e)) (H_2 \otimes H_1)^\top`.

    :return: (K_value, row_idx, col_idx) with the :math:`n_1 n_2` diagonal entries appended
    """
    edge_value = torch.flatten(Ke)
    point_value = torch.flatten(Kp)
    K_value = torch.cat((edge_value, point_value), dim=0)
    diag = torch.linspace(0, point_value.shape[0] - 1, point_value.shape[0], device=row_idx.device)
    row_idx = torch.cat((row_idx, diag), dim=0)
    col_idx = torch.cat((col_idx, diag), dim=0)
    return K_value, row_idx, col_idx


def kronecker_torch(t1: Tensor, t2: Tensor) -> Tensor:
    r"""Batched dense Kronecker product of :math:`T_1` and :math:`T_2`."""
    batch_num = t1.shape[0]
    t1dim1, t1dim2 = t1.shape[1], t1.shape[2]
    t2dim1, t2dim2 = t2.shape[1], t2.shape[2]
    tt = torch.bmm(t1.reshape(batch_num, -1, 1), t2.reshape(batch_num, 1, -1))
    tt = tt.reshape(batch_num, t1dim1, t1dim2, t2dim1, t2dim2).permute([0, 1, 3, 2, 4])
    return tt.reshape(batch_num, t1dim1 * t2dim1, t1dim2 * t2dim2)


def kronecker_sparse(arr1: np.ndarray, arr2: np.ndarray):
    r"""scipy.sparse Kronecker product (cpu, data pipeline)."""
    return ssp.kron(ssp.coo_matrix(arr1), ssp.coo_matrix(arr2))


def construct_aff_mat(Ke: Tensor, Kp: Tensor, KroG, KroH, KroGt=None, KroHt=None) -> Tensor:
    r"""
    Dense affinity matrix :math:`K = diag(vec(K_p)) + (G_2 \otimes G_1) diag(vec(K_e)) (H_2 \otimes H_1)^\top`
    (factorize_graph_matching.py:10-54).  ``KroG``: ``CSRMatrix3d`` [b, n1n2, ne1ne2]; ``KroH``: ``CSCMatrix3d``
    [b, ne1ne2, n1n2] (the transposed Kronecker factor, as the reference's collate function builds it);
    ``KroGt`` / ``KroHt``: their ``transpose(keep_type=True)`` (computed when omitted).  Differentiable in Ke, Kp.
    """
    return RebuildFGM.apply(Ke, Kp, KroG, KroH, KroGt, KroHt)


class RebuildFGM(torch.autograd.Function):
    """factorize_graph_matching.py:140-186.  Forward: one scatter launch over the Kronecker columns instead of the
    reference's CSR.diag + CSR.CSC merge-join over all (n1n2)^2 outputs; backward: the matching gather for dKe and
    the diagonal for dKp (``csrc/sparse.cu``)."""

    @staticmethod
    def forward(ctx, Ke, Kp, Kro1, Kro2, Kro1T=None, Kro2T=None):
        from fpmatch import ops
        if Kro1T is None or Kro2T is None:
            Kro1T, Kro2T = Kro1.transpose(keep_type=True), Kro2.transpose(keep_type=True)
        ctx.K = (Kro1T, Kro2T)
        ctx.shapes = (tuple(Ke.shape), tuple(Kp.shape))
        B = Ke.shape[0]
        ke_vec = Ke.detach().transpose(1, 2).contiguous().view(B, -1).to(torch.float32)
        kp_vec = Kp.detach().transpose(1, 2).contiguous().view(B, -1).to(torch.float32)
        return ops.fgm_rebuild(Kro1T, Kro2T, ke_vec, kp_vec)

    @staticmethod
    def backward(ctx, dK):
        from src.sparse import bilinear_diag_torch
        Kro1T, Kro2T = ctx.K
        (B, e1, e2), (_, n1, n2) = ctx.shapes
        dKe = dKp = None
        if ctx.needs_input_grad[0]:
            dKe = bilinear_diag_torch(Kro1T, dK.contiguous(), Kro2T).view(B, e2, e1).transpose(1, 2)
        if ctx.needs_input_grad[1]:
            dKp = torch.diagonal(dK, dim1=-2, dim2=-1).reshape(B, n2, n1).transpose(1, 2)
        return dKe, dKp, None, None, None, None
