#!/bin/bash
# r2: build variants of the whole library with MoE tuning macros (in parallel) and time the MoE kernels alone at the
# C4 latent shape through tools/moe_bench.py (MMVAE_B200_LIB selects the variant).  Run: `gpurun -- tools/tune_moe.sh`.
set -u
cd "$(dirname "$0")/.."
SRC=multimodal-vae-comparison_b200/csrc
OUT=gpurun_out/tune_moe; mkdir -p $OUT; rm -f $OUT/lib_*.so
declare -A V
V[b3_f4]=""
V[b4_f4]="-DMMVAE_MOE_BWD_STAGES=4"
V[b3_f6]="-DMMVAE_MOE_FWD_STAGES=6"
V[b2_f3]="-DMMVAE_MOE_BWD_STAGES=2 -DMMVAE_MOE_FWD_STAGES=3"
V[b4_f8]="-DMMVAE_MOE_BWD_STAGES=4 -DMMVAE_MOE_FWD_STAGES=8"
for k in "${!V[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude ${V[$k]} -shared \
    $SRC/loglik.cu $SRC/catce.cu $SRC/osigma.cu $SRC/latent.cu $SRC/moe.cu $SRC/combine.cu -o $OUT/lib_$k.so -lcudart > $OUT/build_$k.log 2>&1 &
done
rm -f $OUT/lib_*.so  # keep gpurun_out/ small (64 MiB merge limit)
wait
for k in "${!V[@]}"; do
  echo "== $k ${V[$k]}"
  MMVAE_B200_LIB=$PWD/$OUT/lib_$k.so python tools/moe_bench.py --big 16384 2>&1 | grep "bench M=2 B=16384"
done
rm -f $OUT/lib_*.so  # keep gpurun_out/ small (64 MiB merge limit)
