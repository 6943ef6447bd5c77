"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: executed warp
instructions and stall samples (where does the kernel spend its issue slots?)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
inst, samp, text = defaultdict(float), defaultdict(float), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    if r[1]:
        text[key] = r[1].strip()[:110]
    try:
        inst[key] += float(d.get("Instructions Executed") or 0)
        samp[key] += float(d.get("# Samples") or 0)
    except ValueError:
        pass
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {tot_i:.3e}, stall samples {tot_s:.0f}")
for key in sorted(inst, key=lambda k: -inst[k])[:top]:
    print(f"{100 * inst[key] / tot_i:5.1f}% inst {100 * samp[key] / max(tot_s, 1):5.1f}% samples  {key[0]}:{key[1]:<5d} {text.get(key, '')}")
