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


def tie_stress(B, n, g, device="cpu"):
    """soft-top-k-like matrices: >= 90 % exact zeros, the rest in (0, 1] with many repeated values (1.0, 0.5)."""
    s = torch.rand(B, n, n, generator=g)
    keep = torch.rand(B, n, n, generator=g) < 0.08
    vals = torch.where(torch.rand(B, n, n, generator=g) < 0.5, torch.ones(()), torch.round(s * 4) / 4)
    return (vals * keep).to(device)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="16,32,64,100,128,256,400,512")
    ap.add_argument("--batches", default="1,8,64,256,1024,4096")
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--max-elems", type=float, default=1.1e9, help="skip (B, n) above this many matrix elements")
    ap.add_argument("--variants", default="square10,square20,ragged10,tie_lap")
    args = ap.parse_args()
    from fpmatch import ops
    from oracle import ops as oo
    dev = "cuda"
    rows = []
    # warm the CPU side (first-call costs of torch / scipy used to land in the n = 16 row: 19 ms "per matrix")
    w = torch.randn(4, 24, 24)
    wn = torch.full((4,), 24, dtype=torch.long)
    for _ in range(3):
        oo.hungarian(oo.sinkhorn(w, wn, wn, dummy_row=True, max_iter=10, tau=0.01), wn, wn)
    for variant in args.variants.split(","):
        iters = 20 if variant == "square20" else 10
        for n in [int(x) for x in args.ns.split(",")]:
            g = torch.Generator().manual_seed(n)
            cs = min(args.cpu_sample, 4 if n >= 400 else args.cpu_sample)
            if variant == "ragged10":
                n1c = torch.randint(max(1, n // 2), n + 1, (cs,), generator=g)
                n2c = torch.maximum(n1c, torch.randint(max(1, n // 2), n + 1, (cs,), generator=g))
            else:
                n1c = n2c = torch.full((cs,), n, dtype=torch.long)
            if variant == "tie_lap":
                ss_cpu = tie_stress(cs, n, g)
                cpu_sk, sk_err = None, None
            else:
                s_cpu = torch.randn(cs, n, n, generator=g)
                oo.sinkhorn(s_cpu[:1], n1c[:1], n2c[:1], dummy_row=True, max_iter=iters, tau=0.01)       # warm this size
                t0 = time.perf_counter()
                ss_cpu = oo.sinkhorn(s_cpu, n1c, n2c, dummy_row=True, max_iter=iters, tau=0.01)
                cpu_sk = (time.perf_counter() - t0) / cs * 1e3
                ss_g = ops.sinkhorn_log(s_cpu.to(dev), n1c.to(dev), n2c.to(dev), iters, 0.01, True)
                sk_err = (ss_g.cpu() - ss_cpu).abs().max().item()
            oo.hungarian(ss_cpu[:1], n1c[:1], n2c[:1])                                                    # warm this size
            t0 = time.perf_counter(); x_cpu = oo.hungarian(ss_cpu, n1c, n2c)
            cpu_lap = (time.perf_counter() - t0) / cs * 1e3
            hung_g, _ = ops.lap_topk(ss_cpu.to(dev), n1c.to(dev), n2c.to(dev), want_hungarian=True)
            lap_equal = bool(torch.equal(hung_g.cpu(), x_cpu))
            for B in [int(x) for x in args.batches.split(",")]:
                if B * n * n > args.max_elems:
                    continue
                gg = torch.Generator(device=dev).manual_seed(B + n)
                if variant == "ragged10":
                    n1 = torch.randint(max(1, n // 2), n + 1, (B,), device=dev, generator=gg)
                    n2 = torch.maximum(n1, torch.randint(max(1, n // 2), n + 1, (B,), device=dev, generator=gg))
                else:
                    n1 = n2 = torch.full((B,), n, dtype=torch.long, device=dev)
                reps = 20 if B * n * n < 5e7 else 4
                t_sk = None
                if variant == "tie_lap":
                    ss = tie_stress(min(B, 64), n, g).to(dev).repeat((B + 63) // 64, 1, 1)[:B].contiguous()
                else:
                    s = torch.randn(B, n, n, device=dev, generator=gg)
                    t_sk = gpu_time(lambda: ops.sinkhorn_log(s, n1, n2, iters, 0.01, True), reps)
                    ss = ops.sinkhorn_log(s, n1, n2, iters, 0.01, True)
                    del s
                t_lap = gpu_time(lambda: ops.lap_topk(ss, n1, n2, want_hungarian=True), reps)
                rows.append({"variant": variant, "n": n, "batch": B, "sinkhorn_iters": iters,
                             "gpu_sinkhorn_ms": t_sk, "gpu_lap_ms": t_lap,
                             "gpu_sinkhorn_us_per_matrix": t_sk / B * 1e3 if t_sk else None,
                             "gpu_lap_us_per_matrix": t_lap / B * 1e3,
                             "cpu_sinkhorn_ms_per_matrix": cpu_sk, "cpu_scipy_lap_ms_per_matrix": cpu_lap,
                             "speedup_sinkhorn": cpu_sk / (t_sk / B) if t_sk else None, "speedup_lap": cpu_lap / (t_lap / B),
                             "sinkhorn_bytes_GBs": 2 * B * n * n * 4 / (t_sk * 1e-3) / 1e9 if t_sk else None,
                             "sinkhorn_max_abs_err_vs_oracle": sk_err, "lap_bit_exact_vs_scipy": lap_equal})
                del ss
                torch.cuda.empty_cache()
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep_sinkhorn_lap.json").write_text(json.dumps({"cpu_threads": torch.get_num_threads(), "rows": rows}, indent=1))
    f = lambda v, fmt: "-" if v is None else format(v, fmt)
    print("| variant | n | batch | GPU Sinkhorn ms | GPU LAP ms | CPU Sinkhorn ms/mat | scipy LAP ms/mat | x Sinkhorn | x LAP | LAP exact |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['variant']} | {r['n']} | {r['batch']} | {f(r['gpu_sinkhorn_ms'], '.3f')} | {r['gpu_lap_ms']:.3f} | "
              f"{f(r['cpu_sinkhorn_ms_per_matrix'], '.3f')} | {r['cpu_scipy_lap_ms_per_matrix']:.3f} | "
              f"{f(r['speedup_sinkhorn'], '.0f')} | {r['speedup_lap']:.0f} | {r['lap_bit_exact_vs_scipy']} |")


if __name__ == "__main__":
    main()
