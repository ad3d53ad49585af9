#!/bin/bash
# r1 helper: bench every BASELINE.json configuration at N=1 (run under gpurun); one JSON line per workload
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
python bench.py --workload c2_moe_iwae_cdsprites_l5 --steps 30 --warmup 5 >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c1_poe_elbo_cdsprites_l1 --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c1_poe_elbo_cdsprites_l1 --batch 4096 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c3_mopoe_elbo_sprites --batch 256 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c3_mopoe_elbo_sprites --batch 4096 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c4_moe_dreg_mnistsvhn --batch 1024 --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c4_moe_dreg_mnistsvhn --batch 8192 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python bench.py --workload c5_dmvae_elbo_cub --batch 4096 --dtype bf16 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    r=d['roofline']
    print("%-28s B=%-6d %s  %10.0f samples/s  %.3f ms/step  step %.1f%% of HBM peak | %s %.0f GB/s (%.0f%%) | e2e %s" % (
        d['config']['workload'], d['config']['batch_per_gpu'], d['dtype'], d['value'], d['ms_per_step'], 100*r['step']['frac'],
        r['kernel'].replace('mmvae_',''), r['achieved'] or 0, 100*(r['frac'] or 0), ("%.0f"%d['e2e']['value']) if 'e2e' in d else '-'))
PY
