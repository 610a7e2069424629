"""Host logic of the batched CSR / CSC containers (drop-in for /root/reference/src/sparse_torch/csx_matrix.py)
against scipy.sparse: construction, layout, transposes, concatenation, slicing.  No GPU needed; the products are
covered by tests/test_gpu_sparse.py."""
import numpy as np
import scipy.sparse as ssp
import torch

from src.sparse_torch import CSCMatrix3d, CSRMatrix3d, concatenate


def rand_mats(B, h, w, density, seed):
    rng = np.random.RandomState(seed)
    return [ssp.random(h, w, density=density, random_state=rng, dtype=np.float32, format="coo") for _ in range(B)]


def test_layout_matches_reference_convention():
    mats = rand_mats(3, 5, 7, 0.3, 0)
    m = CSRMatrix3d(mats, shape=(3, 5, 7))
    assert m.indices.dtype == torch.int64 and m.indptr.dtype == torch.int64
    assert m.indptr.numel() == 3 * 5 + 1 and int(m.indptr[-1]) == sum(x.nnz for x in mats)
    off = 0
    for b, x in enumerate(mats):
        c = x.tocsr(); c.sort_indices()
        ind, ptr, dat = m.get_batch(b)
        assert np.array_equal(ind.numpy(), c.indices) and np.array_equal(ptr.numpy(), c.indptr)
        assert np.array_equal(dat.numpy(), c.data)
        assert int(m.indptr[b * 5]) == off
        off += c.nnz
    c = CSCMatrix3d(mats, shape=(3, 5, 7))
    assert c.indptr.numel() == 3 * 7 + 1


def test_roundtrip_and_transposes():
    mats = rand_mats(4, 6, 9, 0.25, 1)
    for cls in (CSRMatrix3d, CSCMatrix3d):
        m = cls(mats, shape=(4, 6, 9))
        for a, b in zip(m.as_ssp(), mats):
            assert np.allclose(a.toarray(), b.toarray())
        t = m.transpose()                       # same arrays, other type
        assert t.sptype != m.sptype and t.shape == (4, 9, 6)
        for a, b in zip(t.as_ssp(), mats):
            assert np.allclose(a.toarray(), b.toarray().T)
        k = m.transpose(keep_type=True)         # re-compressed along the other dimension
        assert k.sptype == m.sptype and k.shape == (4, 9, 6)
        for a, b in zip(k.as_ssp(), mats):
            assert np.allclose(a.toarray(), b.toarray().T)
            assert a.has_sorted_indices or np.all(np.diff(a.indices) != 0)
        assert np.allclose(m.to_dense().numpy(), np.stack([x.toarray() for x in mats]))


def test_concatenate_slicing_padding_from_dense():
    a = CSRMatrix3d(rand_mats(2, 4, 5, 0.4, 2), shape=(2, 4, 5))
    b = CSRMatrix3d(rand_mats(1, 3, 5, 0.4, 3), shape=(1, 3, 5))
    c = concatenate(a, b)
    assert c.shape == (3, 4, 5) and c.indptr.numel() == 3 * 4 + 1
    d = c.to_dense().numpy()
    assert np.allclose(d[:2], a.to_dense().numpy())
    assert np.allclose(d[2, :3], b.to_dense().numpy()[0]) and np.all(d[2, 3] == 0)
    one = c[1]
    assert one.shape == (1, 4, 5) and np.allclose(one.to_dense().numpy()[0], d[1])
    two = c[0:2]
    assert two.shape == (2, 4, 5) and np.allclose(two.to_dense().numpy(), d[:2])
    dense = torch.tensor(d)
    f = CSCMatrix3d.from_dense(dense)
    assert np.allclose(f.to_dense().numpy(), d)
    # a smaller scipy entry is zero padded to the declared batch shape
    g = CSRMatrix3d([ssp.coo_matrix(np.ones((2, 2), np.float32)), ssp.coo_matrix(np.ones((3, 4), np.float32))])
    assert g.shape == (2, 3, 4) and g.indptr.numel() == 7 and g.to_dense()[0, 2].abs().sum() == 0
    assert len(g) == 2 and g.shape_eq(g)


def test_kronecker_incidence_layout_is_what_the_head_consumes():
    """CSCMatrix3d(kron(G2,G1)).indices must be the KGHs_sparse index list of src/gmdataset.py:623-642."""
    from fpmatch import synth
    from utils.factorize_graph_matching import kronecker_sparse
    data = synth.make_batch(2, 6, seed=4, with_kron=True)
    G1, G2, H1, H2 = data["Gs"][0], data["Gs"][1], data["Hs"][0], data["Hs"][1]
    for b in range(2):
        KG = CSCMatrix3d([kronecker_sparse(G2[b].numpy(), G1[b].numpy()).astype(np.float32)])
        KH = CSCMatrix3d([kronecker_sparse(H2[b].numpy(), H1[b].numpy()).astype(np.float32)])
        idxG, idxH = data["KGHs_sparse"][b]
        e1, e2 = int(data["es"][0][b]), int(data["es"][1][b])
        e1max = G1.shape[2]
        # valid columns t = k2 * e1max + k1 (k1 < e1, k2 < e2) hold exactly one entry each
        cols = (torch.arange(e2)[:, None] * e1max + torch.arange(e1)[None, :]).reshape(-1)
        gi = KG.indices[KG.indptr[cols]]
        hi = KH.indices[KH.indptr[cols]]
        assert torch.equal(gi, idxG) and torch.equal(hi, idxH)
