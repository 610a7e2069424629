"""Debug aid: tensor-core GEMM modes at awkward M (1 full tile + a few rows), real and random operands."""
import sys, os, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import ops

torch.manual_seed(0)
dev = "cuda"
res = []
for M in (80, 128, 131, 135, 143, 200, 256, 259, 300, 400):
    for N, K in ((19968, 768), (520, 768)):
        A = torch.randn(M, K, device=dev) * 0.06
        Bt = (torch.rand(N, K, device=dev) - 0.5) * 0.0144
        ref = (A.double() @ Bt.double().T)
        for mode in ("3xf16", "3xtf32", "fp32"):
            C = ops.gemm_nt(A, Bt, mode=mode)
            err = (C.double() - ref).abs()
            worst = err.max().item()
            idx = int(err.argmax())
            res.append({"M": M, "N": N, "K": K, "mode": mode, "max_abs": worst, "row": idx // N, "col": idx % N,
                        "rows_bad": int((err.amax(1) > 1e-6).sum()), "ref_max": ref.abs().max().item()})
            print(res[-1], flush=True)
json.dump(res, open(ROOT / "gpurun_out" / "debug_gemm_ragged.json", "w"))
