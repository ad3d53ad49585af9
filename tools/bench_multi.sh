#!/bin/bash
# r2: multi-GPU bench (run under `gpurun --gpus N`): at N = 2 the 2-rank NCCL / peer-memory parity test first; then the
# headline workload, weak scaling, collectives fused over peer memory (the driver's command) vs NCCL; then ONE process
# group sweeping the C4 DReG latent-only step over global batches 1k..64k (strong) + 16k per GPU (weak), the C4 step with
# likelihoods and the C2 strong split (--sweep: a torchrun start-up per point would cost more than the points).
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/multi_n$N.jsonl; : > $OUT
ERR=gpurun_out/multi_n$N.err; : > $ERR
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" >> $OUT 2>> $ERR
  echo "rc=$? $*" >> $ERR
}
if [ "$N" -eq 2 ]; then
  timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -m gpu -x -q 2>&1 | tail -3
fi
run --steps 20 --warmup 5 --no-cpu-baseline
run --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --nccl-only
L=c4_moe_dreg_latent_only
run --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --sweep $L:1024,$L:2048,$L:4096,$L:8192,$L:16384,$L:32768,$L:65536,$L:16384:w,c4_moe_dreg_mnistsvhn:16384,c4_moe_dreg_mnistsvhn:65536,c2_moe_iwae_cdsprites_l5:256
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith('{'):
        continue
    d = json.loads(l)
    c, r = d['config'], d['roofline']
    print("%-26s N=%d global B=%-6d (%5d/GPU) %-6s %12.0f samples/s  %8.4f ms/step  step %5.1f%% of HBM peak/GPU  parity_n %s  e2e %s  [%s]" % (
        c['workload'], d['n_gpus'], c['global_batch'], c['batch_per_gpu'], d['scaling'], d['value'], d['ms_per_step'],
        100 * r['step']['frac'], (d.get('parity_n') or {}).get('max_rel'), ('%.0f' % d['e2e']['value']) if 'e2e' in d else '-',
        d['run']['collectives'][:40]))
PY
grep -h "rc=[1-9]" $ERR | head; tail -3 $ERR
