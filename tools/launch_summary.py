"""Per-kernel table of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
    python tools/launch_summary.py <launches.csv> <steps profiled> ["title"]
Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes."""
import csv
import sys
from collections import defaultdict

path, n_steps = sys.argv[1], int(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else path
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = defaultdict(list)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    per[r[ik].split("(")[0][:96]].append(v)
# steps actually profiled = launches of the once-per-step combine kernel (bench.py runs a few un-timed steps as well)
for key in ("elbo_combine_kernel", "iwae_kernel", "dreg_stage2_kernel"):
    c = sum(len(v) for k, v in per.items() if key in k)
    if c:
        n_steps = c
        break
tot = sum(sum(v) for v in per.values()) / n_steps
ours = sum(sum(v) for k, v in per.items() if "mmvae::" in k) / n_steps
print("# %s" % title)
print("# %d steps profiled; %.1f us of kernels per step, %.1f us (%.1f %%) in mmvae:: kernels; %d launches per step, %d of them mmvae::" % (
    n_steps, tot, ours, 100 * ours / tot, sum(len(v) for v in per.values()) // n_steps,
    sum(len(v) for k, v in per.items() if "mmvae::" in k) // n_steps))
print("%-96s %7s %10s %9s %7s" % ("kernel", "n/step", "us/step", "avg us", "share"))
for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print("%-96s %7.1f %10.1f %9.1f %6.1f%%" % (name, len(v) / n_steps, sum(v) / n_steps, sum(v) / len(v),
                                               100 * sum(v) / n_steps / tot))
