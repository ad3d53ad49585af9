#!/usr/bin/env python
"""r2 helper (run under gpurun): the MoE sample + log-density kernels alone, at the C4 / C2 latent shapes.
Prints per-direction time, algorithmic GB/s (fwd: eps read + z written; bwd: eps + dz read) and a parity check of the
same call against a torch fp64 evaluation at a small batch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmvae_b200.ops as ops  # noqa: E402
import mmvae_b200.synthetic as syn  # noqa: E402
from oracle import refmath  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def parity(M, B, D, K, dist, rk):
    g = torch.Generator().manual_seed(7)
    post = [syn.make_posterior(g, B, D) for _ in range(M)]
    lg = torch.randn(1, D, generator=g) * 0.3
    noise = [syn.make_noise(g, dist, (K, B, D)) for _ in range(M)]
    w_z = torch.randn(M, K, B, D, generator=g)
    w_lq = torch.randn(M, M, K, B, generator=g)
    w_lp = torch.randn(M, K, B, generator=g)
    out = {}
    for dev, dt in (("cpu", torch.float64), ("cuda", torch.float32)):
        mus = [p[0].to(dev, dt).requires_grad_(True) for p in post]
        ss = [p[1].to(dev, dt).requires_grad_(True) for p in post]
        l = lg.to(dev, dt).requires_grad_(True)
        mu0, s0 = torch.zeros_like(l), torch.softmax(l, 1) * D
        if dev == "cpu":
            z = torch.stack([refmath.rsample(dist, mus[m], ss[m], noise[m].to(dt)) for m in range(M)])
            lq = torch.stack([torch.stack([refmath.log_prob(dist, z[r], mus[j], ss[j]).sum(-1) for j in range(M)])
                              for r in range(M)])
            lpz = torch.stack([refmath.normal_log_prob(z[r], mu0, s0).sum(-1) for r in range(M)])
        else:
            z, lq, lpz = ops.moe_logdens(torch.stack(mus), torch.stack(ss), mu0, s0, torch.stack(noise).to(dev),
                                         [1 if dist == "laplace" else 0] * M, True)
        tot = (z * w_z.to(dev, dt)).sum() + (lq * w_lq.to(dev, dt)).sum() + (lpz * w_lp.to(dev, dt)).sum()
        tot.backward()
        out[dev] = (z, lq, lpz, torch.stack([m.grad for m in mus]), torch.stack([x.grad for x in ss]), l.grad)
    names = ["z", "lq", "lpz", "dmu", "ds", "dlogits"]
    print("parity M=%d B=%d D=%d K=%d %s: " % (M, B, D, K, dist) +
          "  ".join("%s %.1e" % (n, rel(a, b)) for n, a, b in zip(names, out["cuda"], out["cpu"])))


def bench(M, B, D, K, dist, iters=20):
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    mu = torch.randn(M, B, D, generator=g).to(dev).requires_grad_(True)
    s = (torch.softmax(torch.randn(M, B, D, generator=g), -1) + 1e-6).to(dev).requires_grad_(True)
    lg = torch.zeros(1, D, device=dev, requires_grad=True)
    eps = torch.stack([syn.make_noise(g, dist, (K, B, D)) for _ in range(M)]).to(dev)
    dz = torch.randn(M, K, B, D, device=dev) * 0.1
    dlq = torch.randn(M, M, K, B, device=dev)
    dlpz = torch.randn(M, K, B, device=dev)
    codes = [1 if dist == "laplace" else 0] * M
    mu0 = torch.zeros(1, D, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf, tb = [], []
    for it in range(iters + 3):
        s0 = torch.softmax(lg, 1) * D
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda._sleep(200000)
        a.record()
        z, lq, lpz = ops.moe_logdens(mu, s, mu0, s0, eps, codes, True)
        b.record()
        flush.zero_()
        torch.cuda._sleep(200000)
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        torch.autograd.backward([z, lq, lpz], [dz, dlq, dlpz])
        b2.record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(a.elapsed_time(b))
            tb.append(a2.elapsed_time(b2))
        mu.grad = s.grad = lg.grad = None
    big = M * K * B * D * 4
    f, bw = sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2]
    print("bench M=%d B=%d D=%d K=%d %-7s fwd %.3f ms %.0f GB/s | bwd(+prior sum, torch glue) %.3f ms %.0f GB/s" % (
        M, B, D, K, dist, f, 2 * big / f / 1e6, bw, 2 * big / bw / 1e6))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, default=16384)
    a = ap.parse_args()
    for cfg in [(2, 5, 64, 4, "laplace"), (2, 37, 64, 50, "laplace"), (2, 6, 16, 3, "normal"), (3, 9, 20, 5, "laplace"),
                (2, 300, 128, 2, "normal"), (1, 40, 32, 7, "laplace"), (3, 4, 100, 3, "normal")]:
        parity(*cfg, rk=False)
    bench(2, a.big, 64, 50, "laplace")
    bench(2, a.big, 64, 50, "normal")
    bench(2, 256, 16, 30, "normal")
    bench(2, 4096, 16, 30, "normal")
