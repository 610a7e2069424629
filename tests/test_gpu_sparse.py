"""GPU parity of the batched CSR / CSC products and the dense factorised-graph-matching affinity (SURVEY.md
section 8 row A14) against scipy / dense torch evaluations of the same formulae."""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as ssp
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
DEV = "cuda"


def report(name, **kv):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "parity_report.jsonl", "a") as f:
        f.write(json.dumps({"test": name, **kv}) + "\n")


def rand_mats(B, h, w, density, seed):
    rng = np.random.RandomState(seed)
    return [ssp.random(h, w, density=density, random_state=rng, dtype=np.float32, format="coo") for _ in range(B)]


def test_sparse_products_match_scipy():
    from src.sparse_torch import CSCMatrix3d, CSRMatrix3d, dot
    from src.sparse import bilinear_diag_torch
    B, h, k, w = 3, 17, 23, 11
    A = rand_mats(B, h, k, 0.3, 0); Bm = rand_mats(B, k, w, 0.3, 1)
    a = CSRMatrix3d(A, shape=(B, h, k)).cuda(); b = CSCMatrix3d(Bm, shape=(B, k, w)).cuda()
    ref = np.stack([(x @ y).toarray() for x, y in zip(A, Bm)])
    out = dot(a, b, dense_output=True)
    e1 = np.abs(out.cpu().numpy() - ref).max()
    diag = torch.randn(B, k, generator=torch.Generator().manual_seed(2))
    ad = a.dotdiag(diag.to(DEV))
    refd = np.stack([(x @ ssp.diags(diag[i].numpy())).toarray() for i, x in enumerate(A)])
    e2 = np.abs(ad.cpu().to_dense().numpy() - refd).max()
    D = torch.randn(B, h, k, generator=torch.Generator().manual_seed(3))
    e3 = np.abs(dot(D.to(DEV), b, dense_output=True).cpu().numpy()
                - np.stack([D[i].numpy() @ Bm[i].toarray() for i in range(B)])).max()
    # bilinear_diag: diag(S1 T S3), S1 CSR [x,f], T [f,f], S3 CSC [f,x]
    x, f = 13, 9
    S1 = rand_mats(B, x, f, 0.4, 4); S3 = rand_mats(B, f, x, 0.4, 5)
    T = torch.randn(B, f, f, generator=torch.Generator().manual_seed(6))
    got = bilinear_diag_torch(CSRMatrix3d(S1, shape=(B, x, f)).cuda(), T.to(DEV), CSCMatrix3d(S3, shape=(B, f, x)).cuda())
    refb = np.stack([np.diag(S1[i].toarray() @ T[i].numpy() @ S3[i].toarray()) for i in range(B)])
    e4 = np.abs(got.cpu().numpy() - refb).max()
    report("sparse_products", csr_csc=float(e1), dotdiag=float(e2), dense_csc=float(e3), bilinear_diag=float(e4))
    assert max(e1, e2, e3, e4) < 1e-5
    with pytest.raises(NotImplementedError):
        dot(a, b)                               # sparse output: CPU-only in the reference as well
    with pytest.raises(RuntimeError):
        dot(CSRMatrix3d(A, shape=(B, h, k)), CSCMatrix3d(Bm, shape=(B, k, w)), dense_output=True)   # CPU tensors


@pytest.mark.parametrize("n,ragged", [(6, False), (9, True)])
def test_construct_aff_mat_forward_backward(n, ragged):
    """K = diag(vec Kp) + (G2 x G1) diag(vec Ke) (H2 x H1)^T built the reference's way (collate: src/gmdataset.py:
    644-650) versus a dense torch evaluation, values and gradients."""
    from fpmatch import synth
    from src.sparse_torch import CSRMatrix3d
    from utils.factorize_graph_matching import construct_aff_mat, kronecker_sparse, kronecker_torch
    B = 3
    data = synth.make_batch(B, n, seed=8, ragged=ragged, with_kron=False)
    G1, G2, H1, H2 = data["Gs"][0], data["Gs"][1], data["Hs"][0], data["Hs"][1]
    K1G = CSRMatrix3d([kronecker_sparse(x.numpy(), y.numpy()).astype(np.float32) for x, y in zip(G2, G1)])
    K1H = CSRMatrix3d([kronecker_sparse(x.numpy(), y.numpy()).astype(np.float32) for x, y in zip(H2, H1)]).transpose()
    n1, n2, e1, e2 = G1.shape[1], G2.shape[1], G1.shape[2], G2.shape[2]
    gen = torch.Generator().manual_seed(9)
    Ke = torch.randn(B, e1, e2, generator=gen); Kp = torch.randn(B, n1, n2, generator=gen)
    gK = torch.randn(B, n1 * n2, n1 * n2, generator=gen)
    # dense reference
    ke_r = Ke.clone().requires_grad_(True); kp_r = Kp.clone().requires_grad_(True)
    KG, KH = kronecker_torch(G2, G1), kronecker_torch(H2, H1)
    Kref = torch.bmm(KG * ke_r.transpose(1, 2).reshape(B, 1, -1), KH.transpose(1, 2)) \
        + torch.diag_embed(kp_r.transpose(1, 2).reshape(B, -1))
    (Kref * gK).sum().backward()
    ke_g = Ke.to(DEV).requires_grad_(True); kp_g = Kp.to(DEV).requires_grad_(True)
    K = construct_aff_mat(ke_g, kp_g, K1G.cuda(), K1H.cuda())
    (K * gK.to(DEV)).sum().backward()
    ef = (K.detach().cpu() - Kref.detach()).abs().max().item()
    eke = (ke_g.grad.cpu() - ke_r.grad).abs().max().item()
    ekp = (kp_g.grad.cpu() - kp_r.grad).abs().max().item()
    # the generic route of the reference (CSR.diag then CSR.CSC -> dense) must agree with the scatter
    generic = K1G.cuda().dotdiag(Ke.to(DEV).transpose(1, 2).contiguous().view(B, -1)).dot(K1H.cuda(), dense_output=True)
    eg = (generic.cpu() + torch.diag_embed(Kp.transpose(1, 2).reshape(B, -1)) - Kref.detach()).abs().max().item()
    report("construct_aff_mat", n=n, ragged=ragged, forward=ef, dKe=eke, dKp=ekp, generic_route=eg)
    assert max(ef, eke, ekp, eg) < 1e-5


@pytest.mark.parametrize("edge_emb,sk", [(False, 1), (True, 0), (False, 0)])
def test_dense_gnn_layer_matches_reference_formulation(edge_emb, sk):
    """GNNLayer (dense NGM-v1 message passing, gnn.py:11-87): forward and gradients of the
    fused row kernels against the reference's permute + matmul formulation evaluated on the CPU (oracle/ops.py)."""
    from oracle import ops as oo
    from src.model.gnn import GNNLayer
    torch.manual_seed(1)
    in_n, out_e = (1, 16) if not edge_emb else (8, 8)
    layer = GNNLayer(in_n, 1, out_e + sk, out_e, sk_channel=sk, sk_iter=20, sk_tau=0.05, edge_emb=edge_emb)
    g = torch.Generator().manual_seed(2)
    b, n1m, n2m = 3, 6, 7
    N = n1m * n2m
    n1 = torch.tensor([6, 4, 5]); n2 = torch.tensor([7, 7, 3])
    A = (torch.rand(b, N, N, generator=g) < 0.25).float()
    A[0, 5] = 0                                                     # an isolated node: row sum 0 -> normalised to 0
    W = (torch.rand(b, N, N, 1, generator=g) - 0.3) * A.unsqueeze(-1)
    x = torch.rand(b, N, in_n, generator=g)
    up = torch.randn(b, N, out_e + sk, generator=g)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    Wc = W.clone().requires_grad_(True); xc = x.clone().requires_grad_(True)
    Wn_ref, xn_ref = oo.gnn_layer_dense(p, A, Wc, xc, n1, n2, True, 20, 0.05)
    (xn_ref * up).sum().backward()
    layer = layer.to(DEV)
    Wd = W.to(DEV).requires_grad_(True); xd = x.to(DEV).requires_grad_(True)
    Wn, xn = layer(A.to(DEV), Wd, xd, n1.to(DEV), n2.to(DEV), norm=True)
    (xn * up.to(DEV)).sum().backward()
    rel = lambda a, r: ((a.detach().cpu() - r.detach()).abs().max() / r.detach().abs().max().clamp_min(1e-12)).item()
    errs = {"x_new": rel(xn, xn_ref), "W_new": rel(Wn, Wn_ref), "dx": rel(xd.grad, xc.grad), "dW": rel(Wd.grad, Wc.grad)}
    for k, q in layer.named_parameters():
        if p[k].grad is not None and p[k].grad.abs().max() > 0:
            errs["d" + k] = rel(q.grad, p[k].grad)
    report("dense_gnn_layer", edge_emb=edge_emb, sk=sk, **errs)
    tol = 2e-3 if sk else 1e-4            # tau = 0.05 Sinkhorn amplifies fp32 rounding of its input
    assert max(errs.values()) < tol, errs
    with torch.no_grad():                 # inference path (no autograd) gives the same values
        _, xn2 = layer(A.to(DEV), W.to(DEV), x.to(DEV), n1.to(DEV), n2.to(DEV))
    assert (xn2 - xn.detach()).abs().max().item() < 1e-5


def test_legacy_extension_namespaces_match_containers():
    """The reference's raw-tensor entry points (sparse_dot.*, bilinear_diag.bilinear_diag) against the container
    methods that the parity tests above pin to scipy."""
    from src.sparse import bilinear_diag, bilinear_diag_torch
    from src.sparse_torch import CSCMatrix3d, CSRMatrix3d, dot
    from src.sparse_torch.csx_matrix import sparse_dot
    B, h, k, w = 3, 13, 19, 7
    a = CSRMatrix3d(rand_mats(B, h, k, 0.3, 10), shape=(B, h, k)).cuda()
    b = CSCMatrix3d(rand_mats(B, k, w, 0.3, 11), shape=(B, k, w)).cuda()
    got = sparse_dot.csr_dot_csc_to_dense(a.indices, a.indptr, a.data, b.indices, b.indptr, b.data, B, h, w)
    assert torch.equal(got, dot(a, b, dense_output=True))
    D = torch.randn(B, h, k, generator=torch.Generator().manual_seed(12)).to(DEV)
    got = sparse_dot.dense_dot_csc_to_dense(D, b.indices, b.indptr, b.data, B, h, w, k)
    assert torch.equal(got, dot(D, b, dense_output=True))
    diag = torch.randn(B, k, generator=torch.Generator().manual_seed(13)).to(DEV)
    ind, ptr, dat = sparse_dot.csr_dot_diag_to_csr(a.indices, a.indptr, a.data, diag, B, h, k)
    ref = a.dotdiag(diag)
    assert torch.equal(ind, ref.indices) and torch.equal(ptr, ref.indptr) and torch.equal(dat, ref.data)
    with pytest.raises(RuntimeError, match="Unexpected cuda tensor"):
        sparse_dot.csr_dot_csc_to_csr(a.indices, a.indptr, a.data, b.indices, b.indptr, b.data, B, h, w)
    x, f = 11, 9
    s1 = CSRMatrix3d(rand_mats(B, x, f, 0.4, 14), shape=(B, x, f)).cuda()
    s3 = CSCMatrix3d(rand_mats(B, f, x, 0.4, 15), shape=(B, f, x)).cuda()
    T = torch.randn(B, f, f, generator=torch.Generator().manual_seed(16)).to(DEV)
    got = bilinear_diag.bilinear_diag(s1.indices, s1.indptr, s1.data, T, s3.indices, s3.indptr, s3.data, B, x)
    assert torch.equal(got, bilinear_diag_torch(s1, T, s3))
