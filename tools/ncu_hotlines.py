#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (captured with
--import-source on, compiled with -lineinfo).
Usage: ncu_hotlines.py report.ncu-rep [kernel-substring] [top-n] [instr|samples]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by = sys.argv[4] if len(sys.argv) > 4 else "instr"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kern = hdr = fname = None
agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kern = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif kern and want in kern and hdr and r[0].isdigit():
        i, j = hdr.index("Instructions Executed"), hdr.index("# Samples")
        a = agg.setdefault((kern, fname, int(r[0]), r[1]), [0, 0])
        a[0] += int(r[i]) if r[i].isdigit() else 0
        a[1] += int(r[j]) if r[j].isdigit() else 0
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print(f"kernels matching '{want}': {sorted(set(k[0] for k in agg))}")
print(f"total warp instructions {tot}, stall samples {tots}")
key = (lambda kv: -kv[1][0]) if by == "instr" else (lambda kv: -kv[1][1])
for (k, f, ln, src), (i, sm) in sorted(agg.items(), key=key)[:topn]:
    print(f"{f}:{ln:4d} inst {100 * i / max(tot, 1):5.1f}% samp {100 * sm / max(tots, 1):5.1f}%  {src.strip()[:100]}")
