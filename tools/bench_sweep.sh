#!/bin/bash
# r2: the C4 sweep BASELINE.json configs[3] names (MMVAE DReG K=50 D=64, batch 1k..64k) plus the C2 strong split
# (configs[1]: batch 256 over the GPUs), at N ranks of one box.  Usage (under gpurun [--gpus N]): tools/bench_sweep.sh N
set -u
N=${1:-1}
cd "$(dirname "$0")/.."
OUT=gpurun_out/sweep_n$N.jsonl; : > $OUT
ERR=gpurun_out/sweep_n$N.err; : > $ERR
run() {
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" >> $OUT 2>> $ERR
  else
    python bench.py "$@" >> $OUT 2>> $ERR
  fi
}
# global batches 1k .. 64k split over the N ranks (strong scaling of each sweep point; at N = 1 this is the plain sweep)
for GB in 1024 2048 4096 8192 16384 32768 65536; do
  run --workload c4_moe_dreg_latent_only --global-batch $GB --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity
done
for GB in 1024 4096 16384; do
  if [ $((GB / N)) -le 8192 ]; then  # reconstructions + gradients: 3.1 MB per local sample
    run --workload c4_moe_dreg_mnistsvhn --global-batch $GB --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity
  fi
done
# weak scaling of the sweep's largest per-GPU point and the C2 strong split, with the sharded-parity record
run --workload c4_moe_dreg_latent_only --batch 16384 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
run --workload c2_moe_iwae_cdsprites_l5 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity
if [ "$N" -gt 1 ]; then
  # headline workload, weak scaling: collectives fused over peer memory (default) vs NCCL, with the plugin-level e2e
  run --workload c2_moe_iwae_cdsprites_l5 --steps 20 --warmup 5 --no-cpu-baseline
  run --workload c2_moe_iwae_cdsprites_l5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --nccl-only
  run --workload c4_moe_dreg_latent_only --batch 16384 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --nccl-only
fi
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith('{'):
        continue
    d = json.loads(l)
    c, r = d['config'], d['roofline']
    print("%-26s N=%d global B=%-6d (%5d/GPU) %-6s %12.0f samples/s  %8.3f ms/step  step %5.1f%% of HBM peak/GPU  parity_n %s  e2e %s  [%s]" % (
        c['workload'], d['n_gpus'], c['global_batch'], c['batch_per_gpu'], d['scaling'], d['value'], d['ms_per_step'],
        100 * r['step']['frac'], (d.get('parity_n') or {}).get('max_rel'), ('%.0f' % d['e2e']['value']) if 'e2e' in d else '-',
        d['run']['grad_sync'][:28]))
PY
