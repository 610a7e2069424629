"""Drop-in for the live part of ``/root/reference/src/sparse.py``: ``bilinear_diag_torch`` (:182-235), the
backward of ``RebuildFGM``.  The reference JIT-loads its ``bilinear_diag`` extension at import (:10-16); here the
product is ``fpm_bilinear_diag`` of ``libfpmatch_b200.so`` (``csrc/sparse.cu``).  The torch-COO helpers of the
reference file (``sbmm``, ``sdd_bmm_torch`` ...) are unreachable from the matching head and not rebuilt."""
import torch

from fpmatch import ops
from fpmatch.legacy_ext import bilinear_diag       # noqa: F401  (the reference binds its extension under this name, :12)
from src.sparse_torch import CSCMatrix3d, CSRMatrix3d


def bilinear_diag_torch(s_t1: CSRMatrix3d, d_t2: torch.Tensor, s_t3: CSCMatrix3d, device=None):
    """diag(t1 . t2 . t3) per batch entry: t1 CSR [B, x, f], t2 dense [B, f, f], t3 CSC [B, f, x] -> [B, x]."""
    if device is None:
        device = d_t2.device
    assert s_t1.shape[0] == d_t2.shape[0] == s_t3.shape[0], 'Batch size mismatch.'
    assert s_t1.shape[2] == d_t2.shape[1] and d_t2.shape[2] == s_t3.shape[1], 'Matrix shape mismatch'
    assert s_t1.shape[1] == s_t3.shape[2], 'the product is not square'
    return ops.bilinear_diag(s_t1.to(device), d_t2.to(device), s_t3.to(device))
