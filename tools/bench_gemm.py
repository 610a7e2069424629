"""Times the slab GEMM (M=25600, N=19968, K=768) in every mode / kernel with CUDA events (inputs are far larger than
L2 across iterations: the 2 GB output alone evicts everything).  Usage: python tools/bench_gemm.py"""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from fpmatch import ops

M, N, K = 25600, 19968, 768
dev = "cuda"
A = torch.randn(M, K, device=dev) * 0.06
Bt = (torch.rand(N, K, device=dev) - 0.5) * 0.0144
out = torch.empty(M, N, device=dev)
res = []
for mode, pair in (("3xf16", True), ("3xf16", False), ("3xtf32", True), ("3xtf32", False), ("tf32", False)):
    ops.set_gemm_pair(pair)
    for _ in range(3):
        ops.gemm_nt(A, Bt, out=out, mode=mode, weight_operand=True)
    ops.gemm_profile_start()
    for _ in range(10):
        ops.gemm_nt(A, Bt, out=out, mode=mode, weight_operand=True)
    torch.cuda.synchronize()
    ev = ops.gemm_profile_stop()
    ms = sorted(e[4].elapsed_time(e[5]) for e in ev)
    rec = {"mode": mode, "kernel": "pair" if pair else "single", "ms_median": ms[len(ms) // 2], "ms_min": ms[0],
           "tflops_algorithmic": 2.0 * M * N * K / (ms[len(ms) // 2] / 1e3) / 1e12}
    res.append(rec)
    print(json.dumps(rec), flush=True)
ops.set_gemm_pair(True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_gemm.json").write_text(json.dumps(res))
