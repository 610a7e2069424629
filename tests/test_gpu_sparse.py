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
