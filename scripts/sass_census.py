#!/usr/bin/env python
"""SASS census of libbevr.so: per kernel, counts of the instructions that identify the hardware path it uses
(tcgen05 = UTCHMMA / LDTM / UTCBAR, TMA = UBLKCP / UTMA*, legacy tensor cores = HMMA.*TF32, vector REDs, packed FFMA2,
async copies LDGSTS).  Runs without a GPU (cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bevrender_b200", "libbevr.so")
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("UBLKCP", r"\bUBLKCP"),
       ("UTMA*", r"\bUTMA"), ("HMMA.TF32", r"\bHMMA\.\S*TF32"), ("REDG.F32x4", r"\bREDG\.\S*F32x4"),
       ("REDG.F32", r"\bREDG\.\S*F32(?!x)"), ("ATOMS", r"\bATOMS"), ("FFMA2/FMUL2", r"\bF(FMA|MUL)2"),
       ("LDGSTS", r"\bLDGSTS"), ("SHFL", r"\bSHFL"), ("MUFU.EX2", r"\bMUFU\.EX2")]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("bevr::", "").replace("void ", "")
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        total[name] += 1
        for key, pat in PAT:
            if re.search(pat, line):
                counts[name][key] += 1
keys = [k for k, _ in PAT]
print("| kernel | SASS instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for n, c in counts.items():
    print(f"| `{n}` | {total[n]} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
