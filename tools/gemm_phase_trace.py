#!/usr/bin/env python
"""Per-CTA phase timeline of the tcgen05 GEMM at the SplineConv shape (debug / profiling aid).

Records {smid, t_entry, t_setup_done, t_mainloop_done, t_end} for every CTA through fpm_gemm_set_trace and
prints the median duration of each phase plus the gap between consecutive CTAs on the same SM.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "fingerprint-matching-code_b200"))

import torch  # noqa: E402


def main():
    from fpmatch import _lib, ops
    M, N, K = 25600, 19968, 768
    A = torch.randn(M, K, device="cuda") * 0.1
    Bt = torch.randn(N, K, device="cuda") * 0.01
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    out = {}
    for mode in sys.argv[1:] or ["3xf16", "3xtf32", "tf32"]:
        for _ in range(2):
            ops.gemm_nt(A, Bt, mode=mode, weight_operand=True)
        buf = torch.zeros(tiles * 5, dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().fpm_gemm_set_trace(buf.data_ptr(), tiles), "set_trace")
        ops.gemm_nt(A, Bt, mode=mode, weight_operand=True)
        torch.cuda.synchronize()
        _lib.lib().fpm_gemm_set_trace(None, 0)
        t = buf.view(tiles, 5).cpu()
        t = t[t[:, 1] > 0]
        sm, t0, t1, t2, t3 = t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]
        med = lambda x: float(x.float().median())
        gaps = []
        for s in sm.unique().tolist():
            sel = (sm == s).nonzero().flatten()
            order = sel[torch.argsort(t0[sel])]
            g = t0[order][1:] - t3[order][:-1]
            gaps.append(g)
        gaps = torch.cat(gaps)
        res = {"ctas": int(t.shape[0]), "total_us": float(t3.max() - t0.min()) / 1e3,
               "setup_us": med(t1 - t0) / 1e3, "mainloop_us": med(t2 - t1) / 1e3, "epilogue_us": med(t3 - t2) / 1e3,
               "cta_us": med(t3 - t0) / 1e3, "gap_between_ctas_us": med(gaps) / 1e3,
               "gap_p90_us": float(gaps.float().quantile(0.9)) / 1e3}
        out[mode] = res
        print(mode, json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "gemm_phase_trace.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
