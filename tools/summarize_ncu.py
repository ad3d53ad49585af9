"""Turn the ncu outputs of tools/profile_step.sh into the text summaries committed under profiles/.
    python tools/summarize_ncu.py   (needs ncu on PATH to read the .ncu-rep; reads gpurun_out/, writes profiles/)"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")


def launches():
    rows = [r for r in csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "launches_r1.csv")) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = defaultdict(list)
    order = []
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v  # -> us
        name = r[ik].split("(")[0][:84]
        per[name].append(v)
        order.append(name)
    n_steps = max(1, sum(1 for n in order if "iwae_kernel" in n))
    tot = sum(sum(v) for v in per.values()) / n_steps
    bench = json.load(open(os.path.join(OUT, "r1_bench_c2_n1.json")))
    lines = ["# r1: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --streams 1`",
             "# (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised, single stream: compare SHARES)",
             "# workload c2_moe_iwae_cdsprites_l5, B=256, K=30, fp32; %d steps profiled; %.1f us of kernels per step" % (n_steps, tot),
             "# bench.py (CUDA events, graph replay, 3 streams): %.3f ms/step -- two streaming kernels overlap there, so the step is" % bench["ms_per_step"],
             "# SHORTER than the serialised kernel sum; dominant kernel share of the serialised sum below",
             "%-84s %7s %12s %10s %7s" % ("kernel", "n/step", "us/step", "avg us", "share")]
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%-84s %7.1f %12.1f %10.1f %6.1f%%" % (name, len(v) / n_steps, sum(v) / n_steps, sum(v) / len(v),
                                                            100 * sum(v) / n_steps / tot))
    open(os.path.join(OUT, "r1_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def full():
    rep = os.path.join(ROOT, "gpurun_out", "prof_r1_final.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# r1 ncu --set full --clock-control none --import-source on (gpurun_out/prof_r1_final.ncu-rep, scratch)",
           "# bench.py C2 (MMVAE IWAE K=30, B=256, fp32), one eager single-stream step; per-launch values", ""]
    traffic = OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out.append(name)
        for w in WANT:
            if w in idx:
                out.append("    %-82s %s %s" % (w, r[idx[w]], units[idx[w]]))
        def mb(key):
            v, u = float(r[idx[key]].replace(",", "")), units[idx[key]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tb = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        key = None
        if "loglik_kernel<float, float, 0, 0" in name or "loglik_kernel<float, float, (int)0, (int)0" in name:
            key = "mmvae_loglik_rowreduce_fwd"
        if "loglik_kernel<float, float, 0, 1" in name or "loglik_kernel<float, float, (int)0, (int)1" in name:
            key = "mmvae_loglik_rowreduce_bwd"
        if key and key not in traffic:
            traffic[key] = tb
    open(os.path.join(OUT, "r1_ncu_full_c2_kernels.txt"), "w").write("\n".join(out) + "\n")
    if traffic:
        traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, r1 final kernels; "
                            "algorithmic: fwd 390.1 MB, bwd 767.6 MB (bwd writes partly still dirty in the 126 MB L2 at kernel end)")
        json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    launches()
    full()
