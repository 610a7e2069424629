"""CPU tests (no GPU): the oracle against the committed golden vectors (which were produced by the
reference's own modules, see tests/golden/make_golden.py), oracle self-consistency, and host-side logic."""
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"

from oracle import ops as oo          # noqa: E402
from oracle import head, lap           # noqa: E402


def test_feature_align_golden():
    fx = torch.load(GOLD / "feature_align.pt")
    for tag, c in fx.items():
        assert torch.equal(oo.feature_align(c["fmap"], c["P"], c["ns"], (320, 240)), c["out"]), tag
        assert torch.equal(oo.feature_align_loop(c["fmap"], c["P"], c["ns"], (320, 240)), c["out"]), tag


def test_feature_align_quirk_known_answer():
    """with a map that encodes its own (row, col), image corners land on these taps."""
    Hf, Wf = 15, 20
    rows = torch.arange(Hf, dtype=torch.float32).view(Hf, 1).expand(Hf, Wf)
    cols = torch.arange(Wf, dtype=torch.float32).view(1, Wf).expand(Hf, Wf)
    fmap = torch.stack([rows, cols])[None]
    P = torch.tensor([[[0.0, 0.0], [160.0, 120.0], [319.0, 239.0], [319.0, 0.0], [0.0, 239.0]]])
    out = oo.feature_align(fmap, P, torch.tensor([5]), (320, 240))[0]
    np.testing.assert_allclose(out[1].numpy(), [0.0, 7.0, 14.453125, 14.453125, 0.0], atol=1e-4)
    np.testing.assert_allclose(out[0].numpy(), [0.0, 9.5, 14.0, 0.0, 14.0], atol=1e-4)


def test_hungarian_golden():
    fx = torch.load(GOLD / "hungarian.pt")
    assert torch.equal(oo.hungarian(fx["s"], fx["n1"], fx["n2"]), fx["out"])


def test_lap_restatement_matches_scipy_bit_exact():
    """oracle/lap_ref.c (the algorithm the GPU kernel re-states) against the live scipy on tie-heavy input."""
    import scipy.optimize as opt
    rng = np.random.RandomState(0)
    checked = 0
    for t in range(1500):
        n1, n2 = rng.randint(1, 30), rng.randint(1, 30)
        kind = t % 5
        if kind == 0: m = rng.rand(n1, n2)
        elif kind == 1: m = rng.randint(0, 3, (n1, n2)).astype(np.float64)
        elif kind == 2: m = rng.rand(n1, n2) * (rng.rand(n1, n2) < 0.1)
        elif kind == 3: m = np.zeros((n1, n2))
        else: m = np.round(rng.rand(n1, n2), 1)
        cost = -m.astype(np.float32)
        r, c = opt.linear_sum_assignment(cost)
        rr, cc = lap.solve(cost)
        assert np.array_equal(r, rr) and np.array_equal(c, cc), (t, n1, n2)
        checked += 1
    assert checked == 1500


def test_soft_topk_golden():
    fx = torch.load(GOLD / "soft_topk.pt")
    out = oo.soft_topk_prob(fx["scores"], fx["ks"], 10, float(fx["tau"]), fx["nrows"], fx["ncols"])
    assert torch.equal(out, fx["prob"])
    assert torch.equal(oo.hungarian(fx["prob"], fx["nrows"], fx["ncols"]), fx["hungarian"])
    top = torch.argsort(fx["hungarian"].mul(fx["prob"]).reshape(fx["prob"].shape[0], -1), descending=True, dim=-1, stable=True)
    assert torch.equal(oo.greedy_perm(torch.zeros_like(fx["prob"]), top, fx["ks"]), fx["greedy"])
    assert torch.equal(oo.greedy_topk_fast(fx["hungarian"], fx["prob"], fx["ks"]), fx["greedy"])


def test_soft_topk_k_zero_gives_zero_matrix():
    s = torch.rand(1, 6, 6)
    out = oo.soft_topk_prob(s, torch.tensor([0.0]), 10, 0.01, torch.tensor([6]), torch.tensor([6]))
    assert (out == 0).all()


def test_affinity_golden():
    fx = torch.load(GOLD / "affinity.pt")
    for X, Y, w, ref in zip(fx["Xs"], fx["Ys"], fx["Ws"], fx["out"]):
        assert torch.equal(oo.affinity(X, Y, w, fx["A_weight"], fx["A_bias"]), ref)


def test_sinkhorn_known_answers():
    fx = torch.load(GOLD / "sinkhorn_kat.pt")
    for tag, c in fx["cases"].items():
        out = oo.sinkhorn(fx["s"], fx["n1"], fx["n2"], dummy_row=c["dummy_row"], max_iter=c["max_iter"], tau=c["tau"])
        assert (out.double() - c["out"]).abs().max() < 5e-5, tag


def test_sinkhorn_doubly_stochastic_property():
    g = torch.Generator().manual_seed(0)
    s = torch.randn(3, 8, 8, generator=g)
    out = oo.sinkhorn(s, None, None, dummy_row=False, max_iter=60, tau=1.0)
    assert torch.allclose(out.sum(1), torch.ones(3, 8), atol=1e-4)
    assert torch.allclose(out.sum(2), torch.ones(3, 8), atol=1e-3)


def test_afau_golden():
    fx = torch.load(GOLD / "afau.pt")
    p = {"encoder_k." + k: v.float() for k, v in fx["state"].items()}
    r, c = oo.afau_encoder(fx["row"].float(), fx["col"].float(), fx["cost"], p)
    assert (r - fx["out_row"]).abs().max() < 5e-5 and (c - fx["out_col"]).abs().max() < 5e-5


def test_sage_aggregation_equals_factorised_form():
    """The Kronecker factorisation the CUDA kernel relies on against the explicit index lists."""
    from fpmatch import synth
    data = synth.make_batch(3, 9, seed=4, ragged=True, n_min=5, with_kron=True)
    n1max, n2max = data["Ps"][0].shape[1], data["Ps"][1].shape[1]
    N = n1max * n2max
    for b in range(3):
        idxG, idxH = data["KGHs_sparse"][b]
        n1b, n2b = int(data["ns"][0][b]), int(data["ns"][1][b])
        diag = torch.arange(n1b * n2b)
        row, col = torch.cat((idxG, diag)), torch.cat((idxH, diag))
        x = torch.randn(N, 3, dtype=torch.float64)
        ref = oo.sage_mean_aggregate(x, row, col, N)
        A1, A2 = data["As"][0][b].double(), data["As"][1][b].double()        # A[i, j] = 1 iff edge i -> j
        X = x.view(n2max, n1max, 3)                                          # [i2, i1, c]
        agg = torch.einsum("ab,acd,ce->bed", A2, X, A1)                      # sum over in-neighbours
        cnt = torch.einsum("ab,ce->be", A2, A1).reshape(N).clone()           # indeg2(j2) * indeg1(j1)
        agg = agg.reshape(N, 3).clone()
        agg[: n1b * n2b] += x[: n1b * n2b]
        cnt[: n1b * n2b] += 1
        fact = agg / cnt.clamp(min=1)[:, None]
        assert torch.allclose(ref, fact, atol=1e-12)


def effective_structure(t1, t2, e1_pyg, e2_pyg, n1b, n2b):
    """Host mirror of ``assoc_effective_kernel`` (csrc/gnn.cu): what the reference's separately compacted, truncated
    index lists (gmdataset.py:623-642, ngm.py:333-342) mean as a factorised structure."""
    gs1, hd1 = t1[0][t1[0] >= 0].tolist(), t1[1][t1[1] >= 0].tolist()
    gs2, hd2 = t2[0][t2[0] >= 0].tolist(), t2[1][t2[1] >= 0].tolist()
    E1, E2 = min(len(gs1), len(hd1)), min(len(gs2), len(hd2))
    L, nd = E1 * E2, n1b * n2b
    common = min(L + nd, e1_pyg * e2_pyg + nd)
    cut = min(L, common)
    afull, ccut = (cut // E1, cut % E1) if E1 else (0, 0)
    ndiag = max(0, min(nd, common - L))
    eff1 = list(zip(gs1[:E1], hd1[:E1]))
    eff2 = list(zip(gs2[:afull], hd2[:afull]))
    part = (gs2[afull], hd2[afull], ccut) if afull < E2 and ccut > 0 else None
    return eff1, eff2, part, ndiag


@pytest.mark.parametrize("partial,n,seed", [(0, 9, 4), (1, 10, 5), (2, 12, 3), (4, 14, 6), (6, 16, 7)])
def test_partial_permutation_lists_equal_effective_factorised_structure(partial, n, seed):
    """With a PARTIAL ground-truth permutation G2 / H2 lose different columns, the reference's two index lists are
    compacted independently and cut to len(K_value) (SURVEY 5 'common_len').  The CUDA path evaluates the same
    aggregation from an effective Kronecker structure + a cut-off block + a shortened diagonal; this checks that
    derivation against torch_sparse-style aggregation over the explicit (truncated) lists."""
    from fpmatch import synth
    B = 4
    data = synth.make_batch(B, n, seed=seed, partial=partial, with_kron=True)
    n1max, n2max = data["Ps"][0].shape[1], data["Ps"][1].shape[1]
    N = n1max * n2max
    g1, g2 = data["pyg_graphs"]
    saw_part = False
    for b in range(B):
        idxG, idxH = data["KGHs_sparse"][b]
        n1b, n2b = int(data["ns"][0][b]), int(data["ns"][1][b])
        e1b, e2b = int(g1.eptr[b + 1] - g1.eptr[b]), int(g2.eptr[b + 1] - g2.eptr[b])
        diag = torch.arange(n1b * n2b)
        row, col = torch.cat((idxG, diag)), torch.cat((idxH, diag))
        common = min(row.numel(), col.numel(), e1b * e2b + n1b * n2b)           # ngm.py:339
        row, col = row[:common], col[:common]
        x = torch.randn(N, 3, dtype=torch.float64)
        ref = oo.sage_mean_aggregate(x, row, col, N)

        eff1, eff2, part, ndiag = effective_structure(data["edge_lists"][0][b], data["edge_lists"][1][b], e1b, e2b, n1b, n2b)
        saw_part |= part is not None
        X = x.view(n2max, n1max, 3)
        agg = torch.zeros(n2max, n1max, 3, dtype=torch.float64)
        cnt = torch.zeros(n2max, n1max, dtype=torch.float64)
        M1 = torch.zeros(n1max, n1max, dtype=torch.float64)                     # multiplicity of i1 -> j1
        for s_, d_ in eff1:
            M1[s_, d_] += 1
        M2 = torch.zeros(n2max, n2max, dtype=torch.float64)
        for s_, d_ in eff2:
            M2[s_, d_] += 1
        agg += torch.einsum("ab,acd,ce->bed", M2, X, M1)
        cnt += torch.einsum("ab,ce->be", M2, M1)
        if part is not None:
            ps2, pd2, ccut = part
            for s_, d_ in eff1[:ccut]:
                agg[pd2, d_] += X[ps2, s_]
                cnt[pd2, d_] += 1
        agg = agg.reshape(N, 3); cnt = cnt.reshape(N)
        agg[:ndiag] += x[:ndiag]; cnt[:ndiag] += 1
        fact = agg / cnt.clamp(min=1)[:, None]
        assert torch.allclose(ref, fact, atol=1e-12), (b, (ref - fact).abs().max())
    if partial >= 4:
        assert saw_part, "the cut was expected to end inside a Kronecker column block for some pair"


def test_oracle_head_runs_and_is_consistent_between_fp32_and_fp64():
    from fpmatch import synth
    from src.model.ngm import Net
    torch.manual_seed(0)
    net = Net(regression=True).eval()
    sd = net.state_dict()
    data = synth.make_batch(3, 12, seed=1, with_kron=True)
    o32 = head.forward_head(sd, synth.clone_batch(data), data["fmaps"])
    o64 = head.forward_head(sd, synth.clone_batch(data), data["fmaps"], dtype=torch.float64)
    assert (o32["ds_mat"] - o64["ds_mat"].float()).abs().max() < 1e-4
    assert torch.equal(o32["perm_mat"], o64["perm_mat"])
    assert o32["perm_mat"].sum((1, 2)).tolist() == [float(k) for k in o32["k_int"]]
    # loop-structured feature_align (what the CPU baseline times) gives the same result
    o32b = head.forward_head(sd, synth.clone_batch(data), data["fmaps"], feature_align_loops=True)
    assert torch.equal(o32["ds_mat"], o32b["ds_mat"])


def test_graph_oracle_matches_scipy_delaunay():
    """oracle/graphs.py: the empty-circle restatement (what the CUDA kernel evaluates) against scipy's Delaunay
    called the way the reference calls it (utils/build_graphs.py:78-100)."""
    from oracle import graphs as og
    rng = np.random.RandomState(0)
    for t in range(60):
        n = int(rng.choice([1, 2, 3, 4, 5, 8, 17, 50, 100]))
        P = np.stack([rng.uniform(0, 320, n), rng.uniform(0, 240, n)], 1)
        if t % 2:
            P = P.astype(np.float32)
        assert (og.delaunay_adjacency(P) == og.delaunay_adjacency_ref(P.astype(np.float64))).all()
    P = np.stack([rng.uniform(0, 320, 400), rng.uniform(0, 240, 400)], 1)
    A = og.delaunay_adjacency(P)
    assert (A == og.delaunay_adjacency_ref(P)).all() and 2300 < A.sum() < 2400
    # degenerate inputs: collinear -> fully connected, duplicate -> isolated, both as scipy / the reference do
    col = np.array([[0, 0], [1, 1], [2, 2], [3, 3.0]])
    assert (og.delaunay_adjacency(col) == og.fully_connect(col)).all()
    assert (og.delaunay_adjacency_ref(col) == og.fully_connect(col)).all()
    dup = np.stack([rng.uniform(0, 320, 12), rng.uniform(0, 240, 12)], 1); dup[9] = dup[2]
    A = og.delaunay_adjacency(dup)
    assert A[9].sum() == 0 and (A == og.delaunay_adjacency_ref(dup)).all()
    # build_graphs / to_pyg_graph statements: row-major edge order, G H^T = A, pseudo-coordinates in [0, 1]
    P = np.stack([rng.uniform(0, 320, 30), rng.uniform(0, 240, 30)], 1)
    A, G, H, e = og.build_graphs(P, 30, 32, None, "tri", True)
    assert e == int(A.sum()) and (G[:30] @ H[:30].T == A).all()
    x, ei, ea = og.pyg_graph(A, P)
    assert (np.diff(ei[0]) >= 0).all() and ea.min() >= 0 and ea.max() <= 1 and x.dtype == np.float32
    # genuine pairs: A2 = perm^T A1 perm
    perm = np.eye(30, dtype=np.float32)[rng.permutation(30)]
    A2, _, _ = og.permute_adjacency(A, perm)
    assert (A2 == perm.T @ A @ perm).all()


def test_dense_gnn_layer_oracle_equals_einsum():
    """oracle.ops.gnn_layer_dense keeps the reference's permute + matmul expression (gnn.py:66); an independent einsum
    of the same sum must agree, for the single-channel edge tensor and for one channel per node feature."""
    torch.manual_seed(0)
    b, N, F_ = 2, 12, 4
    A = (torch.rand(b, N, N) < 0.3).float()
    x = torch.rand(b, N, 3)
    lin = lambda i, o: (torch.randn(o, i) * 0.3, torch.randn(o) * 0.1)
    p = {}
    for name, i in (("n_func", 3), ("n_self_func", 3)):
        p[name + ".0.weight"], p[name + ".0.bias"] = lin(i, F_)
        p[name + ".2.weight"], p[name + ".2.bias"] = lin(F_, F_)
    mlp = lambda name, t: torch.relu(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(
        t, p[name + ".0.weight"], p[name + ".0.bias"])), p[name + ".2.weight"], p[name + ".2.bias"]))
    An = torch.nn.functional.normalize(A, p=1, dim=2)
    for fe in (1, F_):
        W = torch.rand(b, N, N, fe)
        _, out = oo.gnn_layer_dense(p, A, W, x, torch.tensor([3, 3]), torch.tensor([4, 4]), True)
        x1 = mlp("n_func", x)
        ref = torch.einsum("bij,bijc,bjc->bic", An, W.expand(b, N, N, F_), x1) + mlp("n_self_func", x)
        assert (out - ref).abs().max() < 1e-5


def test_index_lists_and_common_len_cut_match_the_reference_golden():
    """tests/golden/kron_partial.pt was produced by the reference's own kronecker_sparse / CSCMatrix3d /
    construct_sparse_aff_mat (make_kron_golden.py) for complete and PARTIAL ground-truth permutations.  Checked here:
    the generator's KGHs_sparse restatement (fpmatch.synth.add_kron), the [idx; diag] lists with the ngm.py:339 cut as
    oracle/head.py forms them, and the effective factorised structure the CUDA path derives from the edge tables."""
    from fpmatch import synth
    fx = torch.load(GOLD / "kron_partial.pt")
    for tag, c in fx.items():
        kw = c["kw"]
        d = synth.make_batch(4, kw["n"], seed=kw["seed"], partial=kw["partial"], imposter_every=3, with_kron=True)
        g1, g2 = d["pyg_graphs"]
        n1max = d["Ps"][0].shape[1]
        for b in range(4):
            idxG, idxH = d["KGHs_sparse"][b]
            assert torch.equal(idxG.int(), c["idxG"][b]) and torch.equal(idxH.int(), c["idxH"][b]), (tag, b)
            n1b, n2b = int(d["ns"][0][b]), int(d["ns"][1][b])
            e1b, e2b = int(g1.eptr[b + 1] - g1.eptr[b]), int(g2.eptr[b + 1] - g2.eptr[b])
            diag = torch.arange(n1b * n2b)
            row, col = torch.cat((idxG, diag)), torch.cat((idxH, diag))
            common = min(row.numel(), col.numel(), e1b * e2b + n1b * n2b)       # oracle/head.py
            assert common == c["common_len"][b], (tag, b)
            assert torch.equal(row[:common].int(), c["row"][b]) and torch.equal(col[:common].int(), c["col"][b])
            # the effective structure reproduces the reference's (row, col) multiset exactly
            eff1, eff2, part, ndiag = effective_structure(d["edge_lists"][0][b], d["edge_lists"][1][b], e1b, e2b, n1b, n2b)
            pairs = [(s2 * n1max + s1, d2 * n1max + d1) for (s2, d2) in eff2 for (s1, d1) in eff1]
            if part is not None:
                ps2, pd2, ccut = part
                pairs += [(ps2 * n1max + s1, pd2 * n1max + d1) for (s1, d1) in eff1[:ccut]]
            pairs += [(p, p) for p in range(ndiag)]
            ref_pairs = list(zip(c["row"][b].tolist(), c["col"][b].tolist()))
            assert sorted(pairs) == sorted(ref_pairs), (tag, b)
