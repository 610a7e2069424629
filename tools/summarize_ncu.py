#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rN_launches.md [forward_index]
  python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/rN_ncu_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, which=3):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")))
            for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    # one forward = from a global_max_kernel pair to the next pair (2 launches per forward, one per image)
    marks = [i for i, (n, _) in enumerate(rows) if "global_max_kernel" in n]
    starts = marks[0::2]
    start = starts[min(which, len(starts) - 2)]
    end = starts[min(which, len(starts) - 2) + 1]
    agg = collections.OrderedDict()
    for n, t in rows[start:end]:
        n = re.sub(r"^void ", "", re.sub(r"\(.*", "", n))
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list, one forward of the matching head (B=256 pairs, n=100)\n\n"
                f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` "
                f"({len(rows)} launches captured; forward #{which} shown: launches {start}..{end - 1}).\n"
                f"Times are serialised, cold-cache kernel durations: compare SHARES, not absolutes.\n\n"
                f"Total {tot / 1e6:.3f} ms in {end - start} launches.\n\n| ms | share | launches | kernel |\n|---|---|---|---|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t / 1e6:.3f} | {100 * t / tot:.1f}% | {c} | `{n[:110]}` |\n")
    print(open(dst).read())


METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {m: hdr.index(m) for m, _ in METRICS if m in hdr}
    kn = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\nOne row per profiled launch (`--clock-control none`); "
                f"dram bytes are per launch.\n\n| kernel | " + " | ".join(lbl for m, lbl in METRICS if m in col) + " |\n")
        f.write("|---|" + "---|" * len(col) + "\n")
        for d in data:
            name = re.sub(r"^void ", "", re.sub(r"\(.*", "", d[kn]))[:48]
            cells = []
            for m, _ in METRICS:
                if m in col:
                    v, u = d[col[m]], units[col[m]]
                    try:
                        v = f"{float(v.replace(',', '')):.3f}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {u}".strip())
            f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst, int(sys.argv[4]) if len(sys.argv) > 4 else 3)
    else:
        full(src, dst)
