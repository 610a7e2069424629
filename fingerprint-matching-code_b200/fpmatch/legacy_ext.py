"""The reference's two JIT-compiled torch extensions under their own names and positional argument lists.

``/root/reference/src/sparse_torch/csx_matrix.py:10-17`` binds ``sparse_dot`` and ``/root/reference/src/sparse.py:12-16``
binds ``bilinear_diag`` as module-level objects whose functions take raw index / pointer / data tensors
(``src/extension/sparse_dot/sparse_dot.cpp:191-331``, ``src/extension/bilinear_diag/bilinear_diag.cpp:303-326``).  Code
written against those objects keeps working: ``src.sparse_torch.csx_matrix.sparse_dot`` and ``src.sparse.bilinear_diag``
are these namespaces, backed by ``libfpmatch_b200.so`` (``csrc/sparse.cu``) on the current stream.  Device rules follow
the reference: the dense-output products want CUDA tensors, sparse x sparse -> sparse refuses them.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import Tensor

from . import _lib
from .ops import _chk, _count, _stream


def _i64(t: Tensor, name: str):
    return _chk(t.contiguous(), name, torch.int64)


def csr_dot_csc_to_csr(t1_indices, t1_indptr, t1_data, t2_indices, t2_indptr, t2_data, batch_size, out_h, out_w):
    if t1_indices.is_cuda:
        raise RuntimeError("Unexpected cuda tensor in sparse dot sparse -> sparse computation.")   # sparse_dot.cpp:204
    raise RuntimeError("fpmatch: the reference's CPU-only sparse x sparse -> sparse kernel is not rebuilt "
                       "(no CPU path exists); use csr_dot_csc_to_dense on CUDA tensors")


def csr_dot_csc_to_dense(t1_indices, t1_indptr, t1_data, t2_indices, t2_indptr, t2_data, batch_size, out_h, out_w):
    if not t1_indices.is_cuda:
        raise RuntimeError("Unexpected cpu tensor in sparse dot sparse -> dense computation.")     # sparse_dot.cpp:225
    out = torch.empty((batch_size, out_h, out_w), dtype=torch.float32, device=t1_data.device)
    rc = _lib.lib().fpm_csr_dot_csc_dense(_i64(t1_indices, "t1_indices"), _i64(t1_indptr, "t1_indptr"),
                                          _chk(t1_data.contiguous(), "t1_data"), _i64(t2_indices, "t2_indices"),
                                          _i64(t2_indptr, "t2_indptr"), _chk(t2_data.contiguous(), "t2_data"),
                                          out.data_ptr(), int(batch_size), int(out_h), int(out_w), _stream())
    _lib.check(rc, "fpm_csr_dot_csc_dense"); _count()
    return out


def dense_dot_csc_to_dense(t1, t2_indices, t2_indptr, t2_data, batch_size, out_h, out_w, t1_w):
    if not t1.is_cuda:
        raise RuntimeError("Unexpected cpu tensor in dense dot sparse -> dense computation.")      # sparse_dot.cpp:242
    out = torch.empty((batch_size, out_h, out_w), dtype=torch.float32, device=t1.device)
    rc = _lib.lib().fpm_dense_dot_csc_dense(_chk(t1.contiguous(), "t1"), _i64(t2_indices, "t2_indices"),
                                            _i64(t2_indptr, "t2_indptr"), _chk(t2_data.contiguous(), "t2_data"),
                                            out.data_ptr(), int(batch_size), int(out_h), int(t1_w), int(out_w),
                                            _stream())
    _lib.check(rc, "fpm_dense_dot_csc_dense"); _count()
    return out


def csr_dot_diag_to_csr(t1_indices, t1_indptr, t1_data, t2, batch_size, out_h, out_w):
    """Returns ``[indices, indptr, data]`` of the product like the reference (copies of the pattern, scaled data)."""
    out = torch.empty_like(t1_data)
    rc = _lib.lib().fpm_csr_dot_diag(_i64(t1_indices, "t1_indices"), _i64(t1_indptr, "t1_indptr"),
                                     _chk(t1_data.contiguous(), "t1_data"), _chk(t2.contiguous(), "t2"), out.data_ptr(),
                                     int(batch_size), int(out_h), int(out_w), _stream())
    _lib.check(rc, "fpm_csr_dot_diag"); _count()
    return [t1_indices.clone(), t1_indptr.clone(), out]


def bilinear_diag_fn(t1_indices, t1_indptr, t1_data, t2, t3_indices, t3_indptr, t3_data, batch_size, xlen):
    """diag(t1 t2 t3) per batch entry -> [batch_size, xlen]  (bilinear_diag.cpp:303-326)."""
    out = torch.empty((batch_size, xlen), dtype=torch.float32, device=t2.device)
    rc = _lib.lib().fpm_bilinear_diag(_i64(t1_indices, "t1_indices"), _i64(t1_indptr, "t1_indptr"),
                                      _chk(t1_data.contiguous(), "t1_data"), _chk(t2.contiguous(), "t2"),
                                      _i64(t3_indices, "t3_indices"), _i64(t3_indptr, "t3_indptr"),
                                      _chk(t3_data.contiguous(), "t3_data"), out.data_ptr(), int(batch_size), int(xlen),
                                      int(t2.shape[-1]), _stream())
    _lib.check(rc, "fpm_bilinear_diag"); _count()
    return out


sparse_dot = SimpleNamespace(csr_dot_csc_to_csr=csr_dot_csc_to_csr, csr_dot_csc_to_dense=csr_dot_csc_to_dense,
                             dense_dot_csc_to_dense=dense_dot_csc_to_dense, csr_dot_diag_to_csr=csr_dot_diag_to_csr)
bilinear_diag = SimpleNamespace(bilinear_diag=bilinear_diag_fn)
