#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (ncu gpu__time_duration list) and gpurun_out/<tag>_prof.ncu-rep (ncu --set full)
into the markdown summaries committed under profiles/.   Usage: python tools/summarize_profiles.py <tag> [forward_index]"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
GO, PR = ROOT / "gpurun_out", ROOT / "profiles"


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:110]


def launches():
    f = GO / f"{tag}_launches.csv"
    if not f.exists():
        return
    lines = [l for l in f.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    seq = [(short(r["Kernel Name"]), float(r["Metric Value"]) / 1e6) for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
    # one forward = the span between consecutive global_max_kernel launches 2 apart (two images per forward)
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("fpm::global_max_kernel")][::2]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else min(3, len(starts) - 2)
    a, b = starts[k], starts[k + 1]
    agg = OrderedDict()
    for n, t in seq[a:b]:
        e = agg.setdefault(n, [0.0, 0])
        e[0] += t; e[1] += 1
    total = sum(v[0] for v in agg.values())
    out = [f"# ncu launch list, one forward of the matching head (B=256 pairs, n=100) - {tag}", "",
           f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3 --no-cpu-baseline` "
           f"({len(seq)} launches captured; forward #{k} shown: launches {a}..{b - 1}).",
           "Times are serialised, cold-cache kernel durations: compare SHARES, not absolutes.", "",
           f"Total {total:.3f} ms in {b - a} launches.", "", "| ms | share | launches | kernel |", "|---|---|---|---|"]
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| {t:.3f} | {100 * t / total:.1f}% | {c} | `{n}` |")
    (PR / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    print("\n".join(out[:30]))


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
           "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]


def full():
    reps = sorted(GO.glob(f"{tag}_prof*.ncu-rep"))       # one or several captures of the same build (different kernel sets)
    if not reps:
        return
    hdr, units, data = None, None, []
    for rep in reps:
        r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if hdr is None:
            hdr, units = rows[0], rows[1]
        elif rows[0] != hdr:                                # same ncu, same --set: the columns agree; be safe anyway
            pos = {h: i for i, h in enumerate(rows[0])}
            rows = rows[:2] + [[d[pos[h]] if h in pos else "" for h in hdr] for d in rows[2:]]
        data += rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    out = [f"# ncu --set full, one launch of each hot kernel inside bench.py (B=256, n=100) - {tag}", "",
           "`ncu --set full --clock-control none --import-source on`; values are per launch (ncu replays each kernel ~39 passes; "
           "durations here are under the profiler - bench numbers come from CUDA events, not from this file).", ""]
    # one entry per (kernel, grid size): the longest launch of that shape (several GEMMs share the persistent grid)
    def dur(d):
        try:
            return float(d[idx["gpu__time_duration.sum"]].replace(",", ""))
        except Exception:
            return 0.0
    best = {}
    for d in data:
        key = (short(d[idx["Kernel Name"]]), d[idx["launch__grid_size"]] if "launch__grid_size" in idx else "")
        if key not in best or dur(d) > dur(best[key]):
            best[key] = d
    for (name, _), d in best.items():
        out.append(f"## `{name}`  grid {d[idx['Grid Size']] if 'Grid Size' in idx else ''} block {d[idx['Block Size']] if 'Block Size' in idx else ''}")
        out.append("")
        out.append("| metric | value | unit |"); out.append("|---|---|---|")
        for m in cols:
            out.append(f"| {m} | {d[idx[m]]} | {units[idx[m]]} |")
        try:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d[idx["dram__bytes_read.sum"]].replace(",", "")) * mult[units[idx["dram__bytes_read.sum"]]]
            wr = float(d[idx["dram__bytes_write.sum"]].replace(",", "")) * mult[units[idx["dram__bytes_write.sum"]]]
            out.append(f"| dram traffic (read + write) | {(rd + wr) / 1e6:.1f} | Mbyte |")
        except Exception:
            pass
        out.append("")
    (PR / f"{tag}_ncu_full.md").write_text("\n".join(out) + "\n")
    print("\n".join(out[:60]))
    # machine-readable pointer for bench.py's roofline.traffic: the slab GEMM = the longest persistent-grid launch of
    # gemm_tc_pair_kernel in the capture
    import json
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tmul = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

    def val(d, m, table):
        return float(d[idx[m]].replace(",", "")) * table[units[idx[m]]]
    gemms = [d for (name, _), d in best.items() if "gemm_tc_pair_kernel" in name]
    rec = {"tag": tag, "file": f"profiles/{tag}_ncu_full.md", "kernels": {}}
    for (name, grid), d in best.items():
        try:
            rec["kernels"][f"{name} grid={grid}"] = {
                "time_us": val(d, "gpu__time_duration.sum", tmul),
                "dram_bytes": val(d, "dram__bytes_read.sum", mult) + val(d, "dram__bytes_write.sum", mult)}
        except Exception:
            pass
    if gemms:
        d = max(gemms, key=dur)
        rec["slab_gemm"] = {"kernel": short(d[idx["Kernel Name"]]), "grid": d[idx["launch__grid_size"]],
                            "time_us": val(d, "gpu__time_duration.sum", tmul),
                            "dram_bytes": val(d, "dram__bytes_read.sum", mult) + val(d, "dram__bytes_write.sum", mult)}
    (PR / "LATEST_NCU.json").write_text(json.dumps(rec, indent=1) + "\n")


launches()
full()
