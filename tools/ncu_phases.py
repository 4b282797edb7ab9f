#!/usr/bin/env python3
"""Split an .ncu-rep's SASS page of one kernel at its BAR.SYNC instructions and print, per segment (= pipeline phase),
executed warp-instructions, stall samples and the most-sampled instructions with their stall reasons.
usage: tools/ncu_phases.py report.ncu-rep [voxels] [top-N per segment]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
vox = float(sys.argv[2]) if len(sys.argv) > 2 else 2 * 192 ** 3
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[start]
iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[start + 1:] if len(r) > iSrc and r[0].startswith("0x")]
base = int(body[0][0], 16)
segs, cur = [], []
for r in body:
    cur.append(r)
    if re.search(r"\bBAR\.SYNC|\bEXIT\b", r[iSrc]) and not r[iSrc].strip().startswith("@"):
        segs.append(cur); cur = []
if cur:
    segs.append(cur)
tot = sum(int(r[iS] or 0) for r in body) or 1
for k, sg in enumerate(segs):
    n = sum(int(r[iE] or 0) for r in sg); sm = sum(int(r[iS] or 0) for r in sg)
    if n == 0:
        continue
    reasons = collections.Counter()
    for r in sg:
        for i, c in stall_cols:
            reasons[c[6:]] += int(r[i] or 0)
    print(f"== segment {k}: +{int(sg[0][0], 16) - base:#x} .. +{int(sg[-1][0], 16) - base:#x}  warp-instr {n} ({n * 32 / vox:.1f}/voxel)  samples {sm} ({100 * sm / tot:.1f}%)")
    print("   stalls: " + ", ".join(f"{a}={b}" for a, b in reasons.most_common(7)))
    for r in sorted(sorted(sg, key=lambda r: -int(r[iS] or 0))[:topn], key=lambda r: int(r[0], 16)):
        st = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:3]
        print(f"   +{int(r[0], 16) - base:#07x} {r[iSrc].strip()[:60]:60s} exec {r[iE]:>8s} smp {r[iS]:>5s}  " + " ".join(f"{c}={v}" for v, c in st if v))
