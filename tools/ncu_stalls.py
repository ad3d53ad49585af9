#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep source page: python tools/ncu_stalls.py rep kernel-regex [N]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "Source" in r)
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
iS, iX, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
print("kernel regex %s: %d instructions, %d samples" % (pat, len(data), tot))
top = sorted([(int(r[iS]), i) for i, r in enumerate(data) if r[iS].isdigit()], reverse=True)[:n]
for sm, i in sorted(top, key=lambda x: x[1]):
    print("%5d  %-78s %6d %5.1f%%  exec %s" % (i, data[i][isrc].strip()[:78], sm, 100.0 * sm / max(tot, 1), data[i][iX]))
