"""Context number (not a test, not the product): the oracle restatement -- the reference's algorithm written with the
same eager torch ops, device resident -- timed on one B200 on the leaf-tensor protocol of bench.py (SURVEY 8d item
"reference eager torch on 1 B200 ... the restatement that keeps BCE on device").  The unmodified reference itself
additionally bounces BCE through the CPU (objectives.py:405-406), so it is slower than this.

    python tests/eager_torch_gpu_timing.py [workload] [batch]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mmvae_b200.workloads as W  # noqa: E402
from oracle import leafstep  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_moe_iwae_cdsprites_l5"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else None
    cfg, t = W.make_leaves(name, B=B, seed=1234)
    t = {k: ([x.cuda() for x in v] if isinstance(v, list) else v.cuda()) for k, v in t.items()}
    for _ in range(3):
        leafstep.run(cfg, t, device="cuda")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    torch.cuda.reset_peak_memory_stats()
    a.record()
    for _ in range(n):
        leafstep.run(cfg, t, device="cuda")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(json.dumps({"what": "eager torch restatement on GPU (device resident)", "workload": name, "batch": cfg["B"],
                      "ms_per_step": ms, "samples_per_s": cfg["B"] / ms * 1e3,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
