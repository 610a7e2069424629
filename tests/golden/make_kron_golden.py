"""Golden vectors for the Kronecker index lists ``KGHs_sparse`` of collate_fn (/root/reference/src/gmdataset.py:623-642),
produced by the REFERENCE's own ``kronecker_sparse`` (utils/factorize_graph_matching.py:125-137), ``CSCMatrix3d`` and
``construct_sparse_aff_mat`` (:57-95), imported in this container with the JIT build of its CUDA extension stubbed out
(``torch.utils.cpp_extension.load`` -> dummy; the container classes only use it for products, not for ``.indices``).
Covers complete AND partial ground-truth permutations (G2 = perm^T G1, H2 = perm^T H1, gmdataset.py:345-352): the case
in which kron(G2,G1) and kron(H2,H1) drop different columns and ngm.py:339-342 truncates to a common length.
Run from the repo root:  python tests/golden/make_kron_golden.py   -> tests/golden/kron_partial.pt"""
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.utils.cpp_extension as cpp

ROOT = Path(__file__).resolve().parents[2]
cpp.load = lambda *a, **k: types.SimpleNamespace()            # the extension does not compile against torch 2.11
sys.path.insert(0, "/root/reference")
from src.sparse_torch import CSCMatrix3d                       # noqa: E402  (reference)
from utils.factorize_graph_matching import construct_sparse_aff_mat, kronecker_sparse   # noqa: E402  (reference)
sys.path.pop(0)
for m in [k for k in sys.modules if k == "src" or k.startswith("src.") or k == "utils" or k.startswith("utils.")]:
    del sys.modules[m]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
from fpmatch import synth                                       # noqa: E402

cases = {}
for tag, kw in {"complete": dict(partial=0, n=10, seed=2), "partial2": dict(partial=2, n=12, seed=3),
                "partial5": dict(partial=5, n=14, seed=6)}.items():
    d = synth.make_batch(4, kw["n"], seed=kw["seed"], partial=kw["partial"], imposter_every=3, with_kron=False,
                         with_dense_gh=True)
    G1, G2 = d["Gs"]; H1, H2 = d["Hs"]
    idxG, idxH, rows, cols, lens = [], [], [], [], []
    g1, g2 = d["pyg_graphs"]
    for b in range(4):
        # gmdataset.py:627-637, verbatim call pattern
        K1G = [kronecker_sparse(x, y).astype(np.float32) for x, y in zip(G2[b].unsqueeze(0), G1[b].unsqueeze(0))]
        K1H = [kronecker_sparse(x, y).astype(np.float32) for x, y in zip(H2[b].unsqueeze(0), H1[b].unsqueeze(0))]
        kg = CSCMatrix3d(K1G).indices
        kh = CSCMatrix3d(K1H).transpose().indices
        idxG.append(kg.clone()); idxH.append(kh.clone())
        # ngm.py:328-342 with Ke_b [e1_pyg, e2_pyg], Kp_b [n1_b, n2_b] of the right sizes
        e1 = int(g1.eptr[b + 1] - g1.eptr[b]); e2 = int(g2.eptr[b + 1] - g2.eptr[b])
        n1b, n2b = int(d["ns"][0][b]), int(d["ns"][1][b])
        K_value, row_idx, col_idx = construct_sparse_aff_mat(torch.zeros(e1, e2), torch.zeros(n1b, n2b), kg, kh)
        common_len = min(row_idx.numel(), col_idx.numel(), K_value.numel())
        rows.append(row_idx[:common_len].long()); cols.append(col_idx[:common_len].long()); lens.append(common_len)
    i32 = lambda ts: [t.to(torch.int32) for t in ts]
    cases[tag] = {"kw": kw, "common_len": lens, "idxG": i32(idxG), "idxH": i32(idxH), "row": i32(rows), "col": i32(cols)}
    print(tag, [int(x.numel()) for x in idxG], [int(x.numel()) for x in idxH], lens)
torch.save(cases, ROOT / "tests" / "golden" / "kron_partial.pt")
