set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"catce_resident" --launch-skip 6 --launch-count 1 \
      -f -o gpurun_out/prof_r2_c5res python bench.py --workload c5_dmvae_elbo_cub --batch 4096 --dtype bf16 --steps 2 --warmup 1 --no-graph --streams 1 --no-cpu-baseline \
      --no-e2e --no-roofline-timer > gpurun_out/ncu_full_c5res.log 2>&1
tail -1 gpurun_out/ncu_full_c5res.log
