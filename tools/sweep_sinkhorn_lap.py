#!/usr/bin/env python
"""BASELINE.json config 5: batched Sinkhorn + exact GPU LAP sweep, n = 16..512, batch 1..4096, against the
reference's CPU path (pygmtools-style per-pair Sinkhorn loop + scipy linear_sum_assignment, 1 core each as
utils/hungarian.py runs it with nproc=1).

Writes gpurun_out/sweep_sinkhorn_lap.json (+ a markdown table on stdout).  GPU times are CUDA-event times
over `reps` launches with inputs resident; CPU times are measured on a bounded sample of at most
`--cpu-sample` matrices per size and reported per matrix.
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def gpu_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="16,32,64,100,128,256,400,512")
    ap.add_argument("--batches", default="1,8,64,256,1024,4096")
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--max-elems", type=float, default=6e8, help="skip (B, n) above this many matrix elements")
    args = ap.parse_args()
    from fpmatch import ops
    from oracle import ops as oo
    dev = "cuda"
    rows = []
    for n in [int(x) for x in args.ns.split(",")]:
        g = torch.Generator().manual_seed(n)
        cs = min(args.cpu_sample, 4 if n >= 400 else args.cpu_sample)
        s_cpu = torch.randn(cs, n, n, generator=g)
        nn_ = torch.full((cs,), n, dtype=torch.long)
        t0 = time.perf_counter(); ss_cpu = oo.sinkhorn(s_cpu, nn_, nn_, dummy_row=True, max_iter=10, tau=0.01)
        cpu_sk = (time.perf_counter() - t0) / cs * 1e3
        t0 = time.perf_counter(); x_cpu = oo.hungarian(ss_cpu, nn_, nn_)
        cpu_lap = (time.perf_counter() - t0) / cs * 1e3
        # parity on the sample before timing
        sg = s_cpu.to(dev); ng = nn_.to(dev)
        ss_g = ops.sinkhorn_log(sg, ng, ng, 10, 0.01, True)
        hung_g, _ = ops.lap_topk(ss_cpu.to(dev), ng, ng, want_hungarian=True)
        sk_err = (ss_g.cpu() - ss_cpu).abs().max().item()
        lap_equal = bool(torch.equal(hung_g.cpu(), x_cpu))
        for B in [int(x) for x in args.batches.split(",")]:
            if B * n * n > args.max_elems:
                continue
            s = torch.randn(B, n, n, device=dev)
            nb = torch.full((B,), n, dtype=torch.long, device=dev)
            reps = 20 if B * n * n < 5e7 else 5
            t_sk = gpu_time(lambda: ops.sinkhorn_log(s, nb, nb, 10, 0.01, True), reps)
            ss = ops.sinkhorn_log(s, nb, nb, 10, 0.01, True)
            t_lap = gpu_time(lambda: ops.lap_topk(ss, nb, nb, want_hungarian=True), reps)
            rows.append({"n": n, "batch": B, "gpu_sinkhorn_ms": t_sk, "gpu_lap_ms": t_lap,
                         "gpu_sinkhorn_us_per_matrix": t_sk / B * 1e3, "gpu_lap_us_per_matrix": t_lap / B * 1e3,
                         "cpu_sinkhorn_ms_per_matrix": cpu_sk, "cpu_scipy_lap_ms_per_matrix": cpu_lap,
                         "speedup_sinkhorn": cpu_sk / (t_sk / B), "speedup_lap": cpu_lap / (t_lap / B),
                         "sinkhorn_bytes_GBs": 2 * B * n * n * 4 / (t_sk * 1e-3) / 1e9,
                         "sinkhorn_max_abs_err_vs_oracle": sk_err, "lap_bit_exact_vs_scipy": lap_equal})
            del s, ss
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep_sinkhorn_lap.json").write_text(json.dumps({"cpu_threads": torch.get_num_threads(), "rows": rows}, indent=1))
    print("| n | batch | GPU Sinkhorn ms | GPU LAP ms | CPU Sinkhorn ms/mat | scipy LAP ms/mat | x Sinkhorn | x LAP | LAP exact |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['n']} | {r['batch']} | {r['gpu_sinkhorn_ms']:.3f} | {r['gpu_lap_ms']:.3f} | {r['cpu_sinkhorn_ms_per_matrix']:.3f} | "
              f"{r['cpu_scipy_lap_ms_per_matrix']:.3f} | {r['speedup_sinkhorn']:.0f} | {r['speedup_lap']:.0f} | {r['lap_bit_exact_vs_scipy']} |")


if __name__ == "__main__":
    main()
