#!/bin/bash
# r2: one `ncu --set full --clock-control none --import-source on` capture per kernel family on its own workload (run
# under gpurun, ONE GPU; gpurun copies back at most 64 MiB, hence the small launch counts).  .ncu-rep files land in gpurun_out/ (scratch); tools/ncu_metrics.py turns them into the text
# summaries committed under profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
full() {  # name, kernel regex, launch-skip, launch-count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip --launch-count $cnt \
      -f -o gpurun_out/prof_r2_$name python bench.py "$@" --steps 2 --warmup 1 --no-graph --streams 1 --no-cpu-baseline \
      --no-e2e --no-roofline-timer > gpurun_out/ncu_full_$name.log 2>&1
  tail -1 gpurun_out/ncu_full_$name.log
}
full c4lat16k "moe_(fwd|bwd)_flat" 4 2 --workload c4_moe_dreg_latent_only --batch 16384
full c5bf16 "catce_pairs" 6 1 --workload c5_dmvae_elbo_cub --batch 4096 --dtype bf16
full c2 "loglik_kernel" 8 3 --workload c2_moe_iwae_cdsprites_l5
ls -la gpurun_out/prof_r2_*.ncu-rep
