#!/bin/bash
# r2: every BASELINE.json configuration at N = 1 (run under gpurun), one JSON line per workload, EACH with the reference's
# CPU path (oracle port, all host cores, same shapes; bounded sample stated in cpu_baseline.sample) timed beside it.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/bench_all.jsonl; : > $OUT
ERR=gpurun_out/bench_all.err; : > $ERR
b() { python bench.py "$@" --cpu-budget-s 5 >> $OUT 2>> $ERR; }
b --workload c2_moe_iwae_cdsprites_l5 --steps 30 --warmup 5
b --workload c1_poe_elbo_cdsprites_l1 --steps 50 --warmup 5
b --workload c1_poe_elbo_cdsprites_l1 --batch 4096 --steps 30 --warmup 5
b --workload c3_mopoe_elbo_sprites --batch 16 --steps 50 --warmup 5
b --workload c3_mopoe_elbo_sprites --batch 256 --steps 30 --warmup 5
b --workload c3_mopoe_elbo_sprites --batch 4096 --steps 20 --warmup 5 --no-e2e
b --workload c3_mopoe_elbo_vilanro --steps 50 --warmup 5
b --workload c3_poe_elbo_vilanro --batch 4096 --steps 20 --warmup 5 --no-e2e
b --workload c4_moe_dreg_mnistsvhn --batch 1024 --steps 20 --warmup 5
b --workload c4_moe_dreg_mnistsvhn --batch 8192 --steps 10 --warmup 3 --no-e2e
b --workload c4_moe_dreg_latent_only --batch 16384 --steps 20 --warmup 5 --no-e2e
b --workload c5_dmvae_elbo_cub --batch 4096 --dtype bf16 --steps 30 --warmup 5
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    r=d['roofline']; c=d.get('cpu_baseline') or {}
    print("%-28s B=%-6d %s  %10.0f samples/s  %.3f ms/step  step %.1f%% of HBM peak | %s %.0f GB/s (%.0f%%) | e2e %s | cpu %s (%s cores; %s)" % (
        d['config']['workload'], d['config']['batch_per_gpu'], d['dtype'], d['value'], d['ms_per_step'], 100*r['step']['frac'],
        r['kernel'].replace('mmvae_',''), r['achieved'] or 0, 100*(r['frac'] or 0), ("%.0f"%d['e2e']['value']) if 'e2e' in d else '-',
        ("%.0f" % c['value']) if c else '-', c.get('cores','-'), c.get('sample','-')[:40]))
PY
