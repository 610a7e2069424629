#!/usr/bin/env python
"""SASS census of the built library: per kernel, how many tcgen05 / TMEM / TMA instructions the cubin holds.

    python tools/sass_census.py <tag>      ->  profiles/<tag>_sass_census.txt

`cuobjdump -sass` over fingerprint-matching-code_b200/fpmatch/libfpmatch_b200.so (the file the GPU box loads).  The
mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as proof of the Blackwell paths: UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st, TMEM), UTMALDG / UTMASTG (TMA bulk tensor copies), UTCBAR (tcgen05.commit
-> mbarrier), SYNCS (mbarrier try_wait / arrive), plus the legacy HMMA / IMMA that would betray an mma.sync kernel."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "fingerprint-matching-code_b200" / "fpmatch" / "libfpmatch_b200.so"
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOM", "SYNCS",
       "HMMA", "IMMA", "FFMA", "DFMA", "REDUX", "MUFU"]
tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for p in PAT:
            if op.startswith(p):
                per[cur][p] += 1
                break
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
out = [f"# SASS census of libfpmatch_b200.so ({tag}): instruction counts per kernel (cuobjdump -sass, sm_100a)", "",
       "kernel | total | " + " | ".join(PAT), "---|---|" + "---|" * len(PAT)]
tot = collections.Counter()
for (k, c), name in zip(per.items(), demangle):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    out.append(f"`{name[:90]}` | {c['_total']} | " + " | ".join(str(c[p]) if c[p] else "." for p in PAT))
    tot.update(c)
out.append(f"**all {len(per)} kernels** | {tot['_total']} | " + " | ".join(str(tot[p]) for p in PAT))
tc = [re.sub(r"\(.*", "", n).replace("void ", "") for (k, c), n in zip(per.items(), demangle) if c["UTCHMMA"]]
out += ["", f"kernels issuing tcgen05.mma (UTCHMMA): {len(tc)}", *[f"  {n}" for n in tc]]
dst = ROOT / "profiles" / f"{tag}_sass_census.txt"
dst.write_text("\n".join(out) + "\n")
print("\n".join(out[-(len(tc) + 4):]))
print("wrote", dst)
