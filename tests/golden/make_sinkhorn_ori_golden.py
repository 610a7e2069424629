"""Golden vectors for the deprecated non-log Sinkhorn (`Sinkhorn(log_forward=False)`), produced by the REFERENCE's own
`forward_ori` (/root/reference/src/model/sinkhorn.py:89-169) imported in this container with `pygmtools` stubbed (the
module imports it at the top but `forward_ori` never calls it).
Run from the repo root:  python tests/golden/make_sinkhorn_ori_golden.py   -> tests/golden/sinkhorn_ori.pt"""
import importlib.util
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.modules.setdefault("pygmtools", types.ModuleType("pygmtools"))
spec = importlib.util.spec_from_file_location("ref_sinkhorn", "/root/reference/src/model/sinkhorn.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

g = torch.Generator().manual_seed(11)
cases = {}
for tag, (B, R, C, dummy, iters, tau) in {"square": (3, 7, 7, False, 10, 0.5), "ragged_dummy": (4, 6, 9, True, 20, 0.2),
                                          "ragged": (3, 8, 8, False, 11, 1.0)}.items():
    s = torch.randn(B, R, C, generator=g)
    nrows = torch.randint(max(2, R // 2), R + 1, (B,), generator=g)
    ncols = torch.randint(max(R, C // 2), C + 1, (B,), generator=g) if dummy else torch.randint(max(2, C // 2), C + 1, (B,), generator=g)
    if tag == "square":
        nrows = torch.full((B,), R); ncols = torch.full((B,), C)
    sk = mod.Sinkhorn(max_iter=iters, tau=tau, epsilon=1e-4, log_forward=False)
    sr = s.clone().requires_grad_(True)
    out = sk(sr, nrows, ncols, dummy_row=dummy)
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    cases[tag] = {"s": s, "nrows": nrows, "ncols": ncols, "dummy_row": dummy, "max_iter": iters, "tau": tau,
                  "out": out.detach(), "w": w, "grad": sr.grad}
torch.save(cases, ROOT / "tests" / "golden" / "sinkhorn_ori.pt")
print({k: float(v["out"].sum()) for k, v in cases.items()})
