"""Drop-in for ``/root/reference/src/sparse_torch/csx_matrix.py``: batched CSR / CSC containers
(``CSRMatrix3d``, ``CSCMatrix3d``) with ``dot`` / ``dotdiag`` / ``transpose`` / ``concatenate``.

Same layout as the reference (csx_matrix.py:20-93): ``indices`` int64 [nnz] (local column / row ids), ``indptr``
int64 [B*h + 1] (CSR) or [B*w + 1] (CSC) holding GLOBAL offsets, ``data`` [nnz], ``shape`` (B, h, w).  The
reference JIT-compiles a torch extension at import (csx_matrix.py:7-17) whose kernels run on the legacy default
stream; here the products are entry points of ``libfpmatch_b200.so`` (``csrc/sparse.cu``) launched on the current
stream.  Construction from scipy matrices stays on the host, as in the reference's collate function
(src/gmdataset.py:631-642).
"""
import numpy as np
import scipy.sparse as ssp
import torch

from fpmatch import ops
from fpmatch.legacy_ext import sparse_dot          # noqa: F401  (the reference binds its extension under this name, :10)


def _to_tensor(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype) if dtype is not None else x.to(device)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


class CSXMatrix3d:
    sptype = None

    def __init__(self, inp, shape, device=None):
        if isinstance(inp, list) and len(inp) and isinstance(inp[0], ssp.spmatrix):
            assert len(shape) == 3, 'Only 3-dimensional tensor (bxhxw) is supported'
            ind, ptr, dat, off = [], [], [], 0
            for b in range(shape[0]):
                m = ssp.coo_matrix(inp[b])
                if m.shape != tuple(shape[1:3]):         # smaller entries are zero padded to the batch shape
                    m = ssp.coo_matrix((m.data, (m.row, m.col)), shape=tuple(shape[1:3]))
                m.eliminate_zeros()
                sp = (m.tocsc() if self.sptype == 'csc' else m.tocsr()).astype(inp[b].dtype)
                sp.sort_indices()
                ind.append(sp.indices); ptr.append(sp.indptr[:-1].astype(np.int64) + off); dat.append(sp.data)
                off += int(sp.indptr[-1])
            ptr.append(np.array([off], dtype=np.int64))
            inp = [np.concatenate(ind), np.concatenate(ptr), np.concatenate(dat)]
        elif not isinstance(inp, list):
            raise ValueError('Data type {} not understood.'.format(type(inp)))
        ind, ptr, dat = inp
        if isinstance(ind, torch.Tensor) and device is None:
            device = ind.device
        self.indices = _to_tensor(ind, torch.int64, device)
        self.indptr = _to_tensor(ptr, torch.int64, device)
        self.data = _to_tensor(dat, None, device)
        self.shape = tuple(int(s) for s in shape)

    # ------------------------------------------------------------------ basic protocol
    def __len__(self):
        return self.shape[0]

    @property
    def device(self):
        return self.indices.device

    def _major(self):
        """Length of the compressed dimension (rows for CSR, columns for CSC)."""
        return self.shape[1] if self.sptype == 'csr' else self.shape[2]

    def get_batch(self, item):
        """(indices, indptr, data) of batch entry ``item`` (int or [start, stop) pair), indptr re-based to 0."""
        if isinstance(item, int):
            start, stop = item, item + 1
        else:
            start, stop = item
        m = self._major()
        ptr = self.indptr[start * m: stop * m + 1]
        lo, hi = int(ptr[0]), int(ptr[-1])
        return self.indices[lo:hi], ptr - lo, self.data[lo:hi]

    def __getitem__(self, item):
        if isinstance(item, int):
            return self.__class__(list(self.get_batch(item)), shape=[1] + list(self.shape[1:3]))
        if isinstance(item, slice):
            idx = list(range(*item.indices(self.shape[0])))
            parts = [self[b] for b in idx]
            return concatenate(*parts) if parts else None
        raise ValueError('Index type {} not supported.'.format(type(item)))

    def to(self, tgt):
        out = self.__class__([self.indices, self.indptr, self.data], shape=self.shape)
        if isinstance(tgt, torch.dtype):
            out.data = self.data.to(tgt)
        else:
            out.indices, out.indptr, out.data = self.indices.to(tgt), self.indptr.to(tgt), self.data.to(tgt)
        return out

    def cuda(self):
        return self.to(torch.device('cuda'))

    def cpu(self):
        return self.to(torch.device('cpu'))

    def numpy(self):
        return self.indices.cpu().numpy(), self.indptr.cpu().numpy(), self.data.cpu().numpy()

    def as_ssp(self):
        """List of scipy matrices, one per batch entry."""
        ctor = ssp.csr_matrix if self.sptype == 'csr' else ssp.csc_matrix
        out = []
        for b in range(self.shape[0]):
            ind, ptr, dat = (t.cpu().numpy() for t in self.get_batch(b))
            out.append(ctor((dat, ind, ptr), shape=self.shape[1:3]))
        return out

    def as_list(self, mask=None):
        mask = [1, 1, 1] if mask is None else mask
        return [t for t, m in zip((self.indices, self.indptr, self.data), mask) if m]

    def as_sparse_torch(self):
        coo = [m.tocoo() for m in self.as_ssp()]
        idx = np.concatenate([np.stack([np.full(c.nnz, b), c.row, c.col]) for b, c in enumerate(coo)], 1)
        val = np.concatenate([c.data for c in coo])
        return torch.sparse_coo_tensor(torch.as_tensor(idx), torch.as_tensor(val), self.shape).to(self.device)

    def to_dense(self):
        return self.as_sparse_torch().to_dense()

    def shape_eq(self, other):
        return tuple(self.shape) == tuple(other.shape)

    def diagonal(self):
        return torch.diagonal(self.to_dense(), dim1=-2, dim2=-1)

    @classmethod
    def from_dense(cls, dense_tensor, device=None):
        mats = [ssp.coo_matrix(d.detach().cpu().numpy()) for d in dense_tensor]
        return cls(mats, shape=tuple(dense_tensor.shape), device=device if device is not None else dense_tensor.device)

    # ------------------------------------------------------------------ structure changes
    def _expanded_major_ids(self):
        """Global id (b * major + r) of the compressed row/column of every stored entry."""
        counts = self.indptr[1:] - self.indptr[:-1]
        return torch.repeat_interleave(torch.arange(counts.numel(), device=self.device), counts,
                                       output_size=int(self.indices.numel()))

    def _recompress(self, new_cls, new_shape):
        """Same entries, compressed along the OTHER dimension (CSR <-> CSC of the same matrix)."""
        B, major, minor = self.shape[0], self._major(), (self.shape[2] if self.sptype == 'csr' else self.shape[1])
        gid = self._expanded_major_ids()
        b, r = torch.div(gid, major, rounding_mode='floor'), gid % major
        key = (b * minor + self.indices) * major + r            # sort by (batch, other dim, this dim)
        order = torch.argsort(key)
        counts = torch.bincount(b[order] * minor + self.indices[order], minlength=B * minor)
        ptr = torch.zeros(B * minor + 1, dtype=torch.int64, device=self.device)
        ptr[1:] = torch.cumsum(counts, 0)
        return new_cls([r[order], ptr, self.data[order]], shape=new_shape)


class CSCMatrix3d(CSXMatrix3d):
    sptype = 'csc'

    def __init__(self, inp, shape=None, device=None):
        if shape is None:
            shape = _infer_shape(inp)
        super().__init__(inp, shape, device)

    def transpose(self, keep_type=False):
        tshape = (self.shape[0], self.shape[2], self.shape[1])
        if not keep_type:        # CSC of M is CSR of M^T: same arrays
            return CSRMatrix3d([self.indices, self.indptr, self.data], shape=tshape)
        return self.transpose()._recompress(CSCMatrix3d, tshape)

    def Tdot(self, other, *args, **kwargs):
        """self^T . other"""
        return self.transpose().dot(other, *args, **kwargs)


class CSRMatrix3d(CSXMatrix3d):
    sptype = 'csr'

    def __init__(self, inp, shape=None, device=None):
        if shape is None:
            shape = _infer_shape(inp)
        super().__init__(inp, shape, device)

    def transpose(self, keep_type=False):
        tshape = (self.shape[0], self.shape[2], self.shape[1])
        if not keep_type:
            return CSCMatrix3d([self.indices, self.indptr, self.data], shape=tshape)
        return self.transpose()._recompress(CSRMatrix3d, tshape)

    def dot(self, other, *args, **kwargs):
        return dot(self, other, *args, **kwargs)

    def dotdiag(self, other):
        """self . diag(other), other [B, w] -> CSR with the same structure (csx_matrix.py:434-465)."""
        assert other.shape[0] == self.shape[0] and other.shape[1] == self.shape[2], 'Shape mismatch'
        data = ops.csr_dot_diag(self.indices, self.indptr, self.data, other, self.shape)
        return CSRMatrix3d([self.indices, self.indptr, data], shape=self.shape)


def _infer_shape(inp):
    if isinstance(inp, list) and len(inp) and isinstance(inp[0], ssp.spmatrix):
        return (len(inp), max(m.shape[0] for m in inp), max(m.shape[1] for m in inp))
    raise ValueError('shape must be given for raw index / pointer / data input')


def dot(t1, t2, dense_output=False):
    """CSR . CSC or dense . CSC (csx_matrix.py:468-503)."""
    if isinstance(t1, CSRMatrix3d) and isinstance(t2, CSCMatrix3d):
        assert t1.shape[0] == t2.shape[0] and t1.shape[2] == t2.shape[1], 'Shape mismatch'
        if not dense_output:
            # the reference's CUDA path raises here too (sparse_dot.cpp:204); its CPU-only std::list kernel is not rebuilt
            raise NotImplementedError('sparse x sparse -> sparse is not implemented on CUDA; use dense_output=True')
        return ops.csr_dot_csc_dense(t1, t2)
    if isinstance(t1, torch.Tensor) and isinstance(t2, CSCMatrix3d):
        assert t1.shape[0] == t2.shape[0] and t1.shape[2] == t2.shape[1], 'Shape mismatch'
        if not dense_output:
            raise NotImplementedError('Sparse output is not implemented.')
        return ops.dense_dot_csc_dense(t1, t2)
    raise ValueError('Data type not understood.')


def concatenate(*mats, device=None):
    """Stack along the batch dimension; all inputs must share type and the padded (max) shape is used."""
    if device is None:
        device = mats[0].device
    cls = type(mats[0])
    h = max(m.shape[1] for m in mats); w = max(m.shape[2] for m in mats)
    ind, ptr, dat, off, B = [], [], [], 0, 0
    for m in mats:
        assert type(m) is cls, 'Matrices must be the same type'
        major, tgt = m._major(), (h if cls.sptype == 'csr' else w)
        p = m.indptr.to(device)
        for b in range(m.shape[0]):
            pb = p[b * major: (b + 1) * major + 1]
            base = int(pb[0])
            ptr.append(pb[:-1] - base + off)
            if tgt > major:     # padded rows / columns are empty
                ptr.append(torch.full((tgt - major,), int(pb[-1]) - base + off, dtype=torch.int64, device=device))
            off += int(pb[-1]) - base
        ind.append(m.indices.to(device)); dat.append(m.data.to(device)); B += m.shape[0]
    ptr.append(torch.tensor([off], dtype=torch.int64, device=device))
    return cls([torch.cat(ind), torch.cat(ptr), torch.cat(dat)], shape=(B, h, w))
