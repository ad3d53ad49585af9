#!/bin/bash
# ncu evidence for the C2 step (run under gpurun, ONE GPU): (1) launch list with per-launch durations, (2) one
# --set full capture of every kernel of one eager step.  Kernels are serialised by ncu, so the step runs single-stream
# and eager here; numbers printed by bench.py under ncu are not bench values.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --streams 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mmvae|loglik|catce|moe_|iwae|prior_scale|partial_sum|reduce_sum" \
    --launch-skip 36 --launch-count 18 -f -o gpurun_out/prof_r1_final $B > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r1_final.ncu-rep gpurun_out/launches_r1.csv
