#!/bin/bash
# r2: ncu launch lists (gpu__time_duration.sum, --clock-control none) of one eager single-stream step of every bench
# workload family; run under gpurun on ONE GPU.  Kernel times under ncu are cold-cache and serialised.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
prof() {  # name, bench args...
  local name=$1; shift
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$name.csv \
      python bench.py "$@" --steps 2 --warmup 1 --no-graph --streams 1 --no-cpu-baseline --no-e2e --no-roofline-timer > gpurun_out/ncu_$name.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_$name.csv 3 "bench.py $* (eager, one stream, 1 warm-up + 2 timed steps)" > gpurun_out/launches_$name.txt 2>&1
  head -16 gpurun_out/launches_$name.txt
}
prof c2 --workload c2_moe_iwae_cdsprites_l5
prof c4lat16k --workload c4_moe_dreg_latent_only --batch 16384
prof c5bf16 --workload c5_dmvae_elbo_cub --batch 4096 --dtype bf16
prof c1b32 --workload c1_poe_elbo_cdsprites_l1
