"""Generate (and self-check) the golden vectors under tests/golden/ from the REFERENCE's own modules.

Run in the build container only (``/root/reference`` is absent on the GPU box):

    python tests/golden/make_golden.py

It imports the reference modules that still import here - ``utils/hungarian.py``, ``utils/feature_align.py``,
``src/model/soft_topk.py``, ``src/model/afau.py``, ``src/model/affinity_layer.py`` - straight from
``/root/reference`` (nothing is copied), runs them on seeded inputs, checks the oracle restatement
(``oracle/ops.py``) against them, and stores inputs + reference outputs as small ``.pt`` fixtures.  The
pieces owned by absent third-party packages (pygmtools Sinkhorn, PyG SplineConv / SAGEConv) cannot be pinned
this way; for them the script stores hand-derivable known-answer cases computed in float64 numpy.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REF))          # reference's `utils` / `src` packages
sys.path.insert(1, str(ROOT))         # `oracle`

from utils.feature_align import feature_align as ref_feature_align          # noqa: E402
from utils.hungarian import hungarian as ref_hungarian                      # noqa: E402
from src.model.soft_topk import soft_topk as ref_soft_topk, greedy_perm as ref_greedy_perm  # noqa: E402
from src.model.afau import Encoder as RefEncoder                            # noqa: E402
from src.model.affinity_layer import InnerProductWithWeightsAffinity as RefAffinity  # noqa: E402

from oracle import ops as oo                                                 # noqa: E402


def save(name, **tensors):
    torch.save(tensors, HERE / f"{name}.pt")
    size = os.path.getsize(HERE / f"{name}.pt")
    print(f"  wrote {name}.pt ({size / 1024:.1f} KiB)")


def golden_feature_align():
    print("feature_align")
    g = torch.Generator().manual_seed(11)
    cases = {}
    for tag, (C, Hf, Wf) in {"nodes": (8, 15, 20), "edges": (6, 8, 10)}.items():
        B, n = 3, 14
        fmap = torch.randn(B, C, Hf, Wf, generator=g)
        P = torch.rand(B, n, 2, generator=g) * torch.tensor([320.0, 240.0])
        # corner / edge keypoints exercise the clamp + post-fetch edge rule
        P[0, :6] = torch.tensor([[0, 0], [319.99, 239.99], [319.99, 0], [0, 239.99], [160, 120], [5.3, 238.7]])
        ns = torch.tensor([14, 9, 1])
        ref = ref_feature_align(fmap, P, ns, (320, 240))
        a = oo.feature_align_loop(fmap, P, ns, (320, 240))
        b = oo.feature_align(fmap, P, ns, (320, 240))
        assert torch.equal(ref, a), "oracle loop != reference"
        assert torch.equal(ref, b), "oracle vectorised != reference"
        cases[tag] = dict(fmap=fmap, P=P, ns=ns, out=ref)
    save("feature_align", **cases)


def tie_heavy_matrices(rng, count, nmax):
    mats = []
    for t in range(count):
        n1, n2 = rng.randint(1, nmax + 1), rng.randint(1, nmax + 1)
        kind = t % 5
        if kind == 0:
            m = rng.rand(n1, n2)
        elif kind == 1:
            m = rng.randint(0, 3, (n1, n2)).astype(np.float64)
        elif kind == 2:
            m = rng.rand(n1, n2) * (rng.rand(n1, n2) < 0.1)        # ~90 % exact zeros
        elif kind == 3:
            m = np.zeros((n1, n2))
        else:
            m = np.round(rng.rand(n1, n2), 1)
        mats.append(m.astype(np.float32))
    return mats


def golden_hungarian():
    print("hungarian")
    rng = np.random.RandomState(5)
    mats = tie_heavy_matrices(rng, 60, 24)
    R = max(m.shape[0] for m in mats); C = max(m.shape[1] for m in mats)
    s = torch.zeros(len(mats), R, C)
    n1 = torch.zeros(len(mats), dtype=torch.long); n2 = torch.zeros(len(mats), dtype=torch.long)
    for b, m in enumerate(mats):
        s[b, :m.shape[0], :m.shape[1]] = torch.from_numpy(m)
        n1[b], n2[b] = m.shape
    ref = ref_hungarian(s, n1, n2)
    mine = oo.hungarian(s, n1, n2)
    assert torch.equal(ref, mine)
    save("hungarian", s=s, n1=n1, n2=n2, out=ref)


def golden_soft_topk():
    print("soft_topk / greedy_perm")
    g = torch.Generator().manual_seed(3)
    B, R, C = 5, 9, 11
    nrows = torch.tensor([9, 7, 9, 4, 8]); ncols = torch.tensor([11, 11, 6, 9, 8])
    scores = torch.rand(B, R, C, generator=g)
    ks = torch.tensor([3.0, 0.0, 4.4, 2.5, 7.9])
    tau = 0.01
    hard, prob = ref_soft_topk(scores, ks, 10, tau, nrows, ncols, True)
    mine = oo.soft_topk_prob(scores, ks, 10, tau, nrows, ncols)
    assert torch.equal(prob, mine), (prob - mine).abs().max()
    # greedy_perm: reference loop vs the oracle's restatement on a stable candidate order
    x = ref_hungarian(prob, nrows, ncols)
    top = torch.argsort(x.mul(prob).reshape(B, -1), descending=True, dim=-1, stable=True)
    gp_ref = ref_greedy_perm(torch.zeros_like(prob), top, ks)
    gp_mine = oo.greedy_perm(torch.zeros_like(prob), top, ks)
    gp_fast = oo.greedy_topk_fast(x, prob, ks)
    assert torch.equal(gp_ref, gp_mine) and torch.equal(gp_ref, gp_fast)
    save("soft_topk", scores=scores, ks=ks, nrows=nrows, ncols=ncols, tau=torch.tensor(tau), prob=prob,
         hard=hard, hungarian=x, greedy=gp_ref)


def golden_afau():
    print("afau Encoder")
    torch.manual_seed(7)
    enc = RefEncoder().eval()
    B, n1, n2 = 2, 7, 9
    g = torch.Generator().manual_seed(8)
    row = torch.randn(B, n1, 600, generator=g) * 0.1
    col = torch.randn(B, n2, 600, generator=g) * 0.1
    cost = torch.rand(B, n1, n2, generator=g)
    with torch.no_grad():
        r_ref, c_ref = enc(row, col, cost)
        p = {"encoder_k." + k: v for k, v in enc.state_dict().items()}
        r, c = oo.afau_encoder(row, col, cost, p)
        assert (r - r_ref).abs().max() < 2e-5 and (c - c_ref).abs().max() < 2e-5, \
            ((r - r_ref).abs().max(), (c - c_ref).abs().max())
        # the structured inputs of ngm.py:392-399: zero rows, one-hot columns
        row0 = torch.zeros(B, n1, 600); col0 = torch.zeros(B, n2, 600)
        for b, nb in enumerate([9, 6]):
            col0[b, torch.arange(nb), torch.arange(nb)] = 1
        r0_ref, c0_ref = enc(row0, col0, cost)
        r0, c0 = oo.afau_encoder(row0, col0, cost, p)
        # InstanceNorm over near-constant channels (one-hot + bias) amplifies fp32 rounding ~300x: both the
        # reference and the oracle sit ~3e-5..6e-5 from a float64 evaluation here, on values of size ~2.8
        assert (r0 - r0_ref).abs().max() < 2e-4 and (c0 - c0_ref).abs().max() < 2e-4
    # only the small mixed-score parameters and the outputs are stored; the 600-wide linears are
    # re-created from the seed by the test (torch.manual_seed(7); RefEncoder-compatible init order)
    small = {k: v for k, v in enc.state_dict().items()}
    save("afau", state={k: v.half() if v.numel() > 20000 else v for k, v in small.items()},
         row=row.half(), col=col.half(), cost=cost, note="big tensors stored in fp16; outputs below were "
         "computed from the fp16-rounded values cast back to fp32")
    # recompute outputs from the rounded values so the fixture is self-consistent
    sd = {k: (v.half().float() if v.numel() > 20000 else v) for k, v in small.items()}
    enc.load_state_dict(sd)
    with torch.no_grad():
        r_ref, c_ref = enc(row.half().float(), col.half().float(), cost)
        r0_ref, c0_ref = enc(row0, col0, cost)
    fx = torch.load(HERE / "afau.pt")
    fx.update(out_row=r_ref, out_col=c_ref, out_row_struct=r0_ref, out_col_struct=c0_ref,
              n2_struct=torch.tensor([9, 6]))
    torch.save(fx, HERE / "afau.pt")
    print(f"  afau.pt now {os.path.getsize(HERE / 'afau.pt') / 1024:.1f} KiB")


def golden_affinity():
    print("affinity layer")
    torch.manual_seed(9)
    aff = RefAffinity(32, 16).eval()
    g = torch.Generator().manual_seed(10)
    Xs = [torch.randn(n, 16, generator=g) for n in (5, 7, 3)]
    Ys = [torch.randn(n, 16, generator=g) for n in (6, 7, 4)]
    Ws = torch.randn(3, 32, generator=g)
    with torch.no_grad():
        ref = aff(Xs, Ys, Ws)
        mine = [oo.affinity(X, Y, w, aff.A.weight, aff.A.bias) for X, Y, w in zip(Xs, Ys, Ws)]
    for a, b in zip(ref, mine):
        assert torch.equal(a, b)
    save("affinity", Xs=Xs, Ys=Ys, Ws=Ws, A_weight=aff.A.weight.detach(), A_bias=aff.A.bias.detach(), out=ref)


def golden_sinkhorn_kat():
    """Known-answer cases for the pygmtools-owned Sinkhorn, computed independently in float64 numpy from the
    published algorithm (PARITY UNPINNED against pygmtools itself - see oracle/__init__.py)."""
    print("sinkhorn known-answer (float64 numpy)")

    def np_sinkhorn(s, n1, n2, max_iter, tau, dummy_row):
        out = np.zeros_like(s, dtype=np.float64)
        for b in range(s.shape[0]):
            r, c = int(n1[b]), int(n2[b])
            m = s[b, :r, :c].astype(np.float64)
            tr = r > c
            if tr:
                m = m.T; r, c = c, r
            L = m / tau
            if dummy_row:
                L = np.concatenate([L, np.full((c - r, c), -100.0)], 0)
            for i in range(max_iter):
                ax = 1 if i % 2 == 0 else 0
                mx = L.max(axis=ax, keepdims=True)
                L = L - (mx + np.log(np.exp(L - mx).sum(axis=ax, keepdims=True)))
            res = np.exp(L[:r])
            if tr:
                res = res.T
            out[b, :res.shape[0], :res.shape[1]] = res
        return out

    rng = np.random.RandomState(21)
    s = rng.randn(6, 7, 9).astype(np.float32)
    n1 = np.array([7, 5, 3, 7, 6, 2]); n2 = np.array([9, 9, 4, 6, 6, 2])   # includes n1 > n2 (per-sample transpose)
    cases = {}
    for tag, (it, tau, dummy) in {"it10_tau1_dummy": (10, 1.0, True), "it20_tau005_dummy": (20, 0.05, True),
                                  "it10_tau05_nodummy": (10, 0.5, False)}.items():
        ref = np_sinkhorn(s, n1, n2, it, tau, dummy)
        mine = oo.sinkhorn(torch.from_numpy(s), torch.from_numpy(n1), torch.from_numpy(n2), dummy_row=dummy,
                           max_iter=it, tau=tau)
        err = np.abs(mine.numpy() - ref).max()
        assert err < 5e-5, (tag, err)
        cases[tag] = dict(out=torch.from_numpy(ref), max_iter=it, tau=tau, dummy_row=dummy)
    save("sinkhorn_kat", s=torch.from_numpy(s), n1=torch.from_numpy(n1), n2=torch.from_numpy(n2), cases=cases)


if __name__ == "__main__":
    golden_feature_align()
    golden_hungarian()
    golden_soft_topk()
    golden_affinity()
    golden_sinkhorn_kat()
    golden_afau()
    print("all golden vectors generated and the oracle agrees with the reference on them")
