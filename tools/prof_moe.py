#!/usr/bin/env python
"""r2 helper for `ncu -k regex:moe_.*flat`: a few fwd+bwd launches of the MoE kernels at the C4 latent shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmvae_b200.ops as ops  # noqa: E402
import mmvae_b200.synthetic as syn  # noqa: E402

M, B, D, K = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 64, 50
dist = sys.argv[2] if len(sys.argv) > 2 else "laplace"
g = torch.Generator().manual_seed(1)
dev = "cuda"
mu = torch.randn(M, B, D, generator=g).to(dev).requires_grad_(True)
s = (torch.softmax(torch.randn(M, B, D, generator=g), -1) + 1e-6).to(dev).requires_grad_(True)
eps = torch.stack([syn.make_noise(g, dist, (K, B, D)) for _ in range(M)]).to(dev)
dz = torch.randn(M, K, B, D, device=dev) * 0.1
dlq = torch.randn(M, M, K, B, device=dev)
dlpz = torch.randn(M, K, B, device=dev)
mu0, s0 = torch.zeros(1, D, device=dev), torch.ones(1, D, device=dev)
for _ in range(3):
    z, lq, lpz = ops.moe_logdens(mu, s, mu0, s0, eps, [1 if dist == "laplace" else 0] * M, True)
    torch.autograd.backward([z, lq, lpz], [dz, dlq, dlpz])
    mu.grad = s.grad = None
torch.cuda.synchronize()
