"""GPU parity at pore-level sizes (BASELINE.json config 4: 400 keypoints/image): the paths that leave shared
memory (Sinkhorn / soft-top-k global scratch, LAP cost matrix read through L2) and size-independent
properties at full size."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
DEV = "cuda"


def report(name, **kv):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "parity_report.jsonl", "a") as f:
        f.write(json.dumps({"test": name, **kv}) + "\n")


@pytest.mark.parametrize("path", ["cluster", "global"])
def test_soft_topk_beyond_shared_memory(path, monkeypatch):
    """Plans beyond one CTA's shared memory: the 8-CTA cluster kernel (slices in distributed shared memory) and the
    global-workspace kernel it falls back to for still larger plans (forced here through FPMATCH_STK_GLOBAL)."""
    from fpmatch import ops
    from oracle import ops as oo
    if path == "global":
        monkeypatch.setenv("FPMATCH_STK_GLOBAL", "1")
    g = torch.Generator().manual_seed(0)
    B, n = 3, 260
    n1 = torch.tensor([260, 200, 230]); n2 = torch.tensor([260, 260, 190])
    ss = oo.sinkhorn(torch.randn(B, n, n, generator=g), n1, n2, dummy_row=True, max_iter=10, tau=0.05)
    ks = torch.tensor([100.0, 0.0, 57.5])
    ref = oo.soft_topk_prob(ss, ks, 10, 0.01, n1, n2)
    out = ops.soft_topk(ss.to(DEV), ks.to(DEV), n1.to(DEV), n2.to(DEV), 10, 0.01)
    err = (out.cpu() - ref).abs().max().item()
    report("soft_topk_beyond_smem", path=path, max_abs=err)
    assert err < 1e-4


def test_head_400_keypoints_matches_oracle():
    from fpmatch import synth
    from oracle import head
    from src.model.ngm import Net
    torch.manual_seed(0)
    net = Net(regression=True).eval()
    data = synth.make_batch(2, 400, seed=4, with_kron=True, with_dense_gh=False)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = head.forward_head(sd, synth.clone_batch(data), data["fmaps"], regression=True)
    net = net.to(DEV)
    with torch.no_grad():
        out = net(synth.batch_to(synth.clone_batch(data), DEV))
    ds_err = (out["ds_mat"].cpu() - ref["ds_mat"]).abs().max().item()
    k_err = (out["k_prob"].cpu() - ref["k_prob"]).abs().max().item()
    perm_equal = bool(torch.equal(out["perm_mat"].cpu(), ref["perm_mat"]))
    report("head_n400", ds_mat=ds_err, k_prob=k_err, perm_equal=perm_equal,
           ss=(out["_fpm_inter"]["ss"].cpu() - ref["inter"]["ss"]).abs().max().item())
    assert ds_err < 1e-4 and k_err < 1e-4 and perm_equal


def test_full_size_properties_256_pairs_100_keypoints():
    """At BASELINE.json's full size the oracle is too slow; check the properties the domain offers."""
    from fpmatch import synth
    from src.model.ngm import Net
    torch.manual_seed(0)
    net = Net(regression=True).eval().to(DEV)
    data = synth.batch_to(synth.make_batch(256, 100, seed=1234, with_dense_gh=False), DEV)
    with torch.no_grad():
        out = net(data)
    ds, perm, ks = out["ds_mat"], out["perm_mat"], out["k_prob"]
    inter = out["_fpm_inter"]
    assert torch.isfinite(ds).all() and ds.min() >= 0 and ds.max() <= 1 + 1e-5
    # perm_mat is a partial permutation with exactly round(k * min(n1, n2)) ones
    assert ((perm == 0) | (perm == 1)).all()
    assert perm.sum(1).max() <= 1 and perm.sum(2).max() <= 1
    k_int = torch.round(inter["k_scaled"])
    assert torch.equal(perm.sum((1, 2)), k_int)
    # the final Sinkhorn output is row-stochastic over the valid block (10 iterations end on a column step:
    # columns sum to <= 1, rows approximately)
    ss = inter["ss"]
    assert (ss.sum(1) <= 1 + 1e-4).all()
    # soft top-k moves at most mass k into the "max" anchor column (exactly k unless the closing
    # row-normalisation of soft_topk.py:232 fired, which can only shrink entries)
    assert (ds.sum((1, 2)) <= inter["k_scaled"] * (1 + 2e-3) + 1e-2).all()
    # hungarian on ds_mat is a full assignment containing the kept top-k matches
    from utils.hungarian import hungarian
    hung = hungarian(ds, data["ns"][0], data["ns"][1])
    assert torch.equal(hung.sum((1, 2)), torch.minimum(data["ns"][0], data["ns"][1]).float())
    kept_positive = (perm * ds) > 0
    assert (hung[kept_positive] == 1).all()
