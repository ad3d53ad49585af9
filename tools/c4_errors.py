#!/usr/bin/env python
"""r2 helper: deviation of the CUDA leaf step AND of the fp32 oracle (= the reference's arithmetic) from the fp64 oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmvae_b200.workloads as W  # noqa: E402
from oracle import leafstep  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


for name, B in (("c4_moe_dreg_mnistsvhn", 6), ("c4_moe_dreg_latent_only", 37), ("c2_moe_iwae_cdsprites_l5", 4),
                ("c2_moe_iwae_cdsprites_l5", 16)):
    for seed in (77, 78, 79):
        cfg, t = W.make_leaves(name, B=B, seed=seed)
        t["pz_logits"] = torch.randn(1, cfg["D"], generator=torch.Generator().manual_seed(5)) * 0.3
        l64, g64 = leafstep.run(cfg, t, beta=1.3, dtype=torch.float64)
        l32, g32 = leafstep.run(cfg, t, beta=1.3, dtype=torch.float32)
        step = W.LeafStep(cfg, t, beta=1.3)
        loss = step.run()
        torch.cuda.synchronize()
        ours = {"mu": step.mu.grad, "s": step.s.grad, "pz_logits": step.pz_logits.grad}
        for i, r in enumerate(step.recon):
            ours["recon%d" % i] = r.grad
        print("%s B=%d seed %d: loss ours %.1e ref32 %.1e | " % (name, B, seed, rel(loss, l64), rel(l32, l64)) +
              " ".join("%s %.1e/%.1e" % (k, rel(ours[k], g64[k]), rel(g32[k], g64[k])) for k in ours if g64[k] is not None))
