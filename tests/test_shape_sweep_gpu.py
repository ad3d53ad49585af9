"""GPU: seeded random shape sweep of the likelihood kernels against the oracle evaluated in fp64 -- exercises every
dispatch path of loglik.cu / catce.cu (128-bit vs scalar loads, rows split over CTAs, several short rows per CTA,
L2-tiled row order, row strides, bf16, TMA / cp.async / plain staging, 1..8 warps per row)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import refmath  # noqa: E402


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


def _cases(seed, n):
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        K = rnd.choice([1, 1, 2, 3, 5])
        B = rnd.choice([1, 2, 3, 7, 8, 16, 33, 64])
        P = rnd.choice([1, 3, 4, 7, 8, 63, 64, 100, 189, 256, 784, 1000, 1024, 3072, 4097, 12288, 20000])
        lt = rnd.choice(["bce", "lprob", "lprob", "mse", "l1", "bce_logits"])
        lik = rnd.choice(["normal", "laplace"])
        dt = rnd.choice([torch.float32, torch.float32, torch.bfloat16])
        pad = rnd.choice([0, 0, 0, 5, 8])  # row stride larger than P
        out.append((K, B, P, lt, lik, dt, pad))
    return out


@pytest.mark.parametrize("K,B,P,lt,lik,dt,pad", _cases(2024, 48))
def test_loglik_random_shapes(K, B, P, lt, lik, dt, pad):
    import mmvae_b200.ops as ops
    g = torch.Generator().manual_seed(K * 1000003 + B * 7919 + P)
    rows = K * B
    full = torch.randn(rows, P + pad, generator=g)
    if lt == "bce":
        full = torch.sigmoid(full).clamp(1e-6, 1 - 1e-6)
    full = full.to(dt)
    t = torch.rand(B, P, generator=g)
    w = torch.randn(rows, generator=g)
    xo = full.double().requires_grad_(True)
    ref = refmath.lpx_rows(lt, xo[:, :P], t.double(), 0.7, K, lik)
    (ref * w.double()).sum().backward()
    xc = full.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc[:, :P], t.cuda(), lt, lik, 0.7)
    (out * w.cuda()).sum().backward()
    vt, gt = (2e-5, 2e-5) if dt == torch.float32 else (2e-4, 1.5e-2)
    assert _rel(out, ref) < vt, "value"
    assert _rel(xc.grad[:, :P], xo.grad[:, :P]) < gt, "grad"
    assert float(xc.grad[:, P:].abs().sum()) == 0.0  # padding columns untouched
    xc2 = full.cuda().requires_grad_(True)
    S, rows_f = ops.loglik_weighted_sum(xc2[:, :P], t.cuda(), lt, lik, 0.7, w_rows=w.cuda())
    S.backward()
    assert torch.equal(rows_f, out.detach()), "fused rows differ from forward rows"
    assert _rel(xc2.grad[:, :P], xo.grad[:, :P]) < gt, "fused grad"


def _catce_cases(seed, n):
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        K = rnd.choice([1, 1, 2, 4])
        B = rnd.choice([1, 3, 4, 8, 16, 20])
        C = rnd.choice([2, 7, 9, 45, 64, 100, 246])
        d = rnd.choice([1, 1, 6, 27, 27, 33, 64])
        dt = rnd.choice([torch.float32, torch.float32, torch.bfloat16])
        out.append((K, B, C, d, dt))
    return out


@pytest.mark.parametrize("K,B,C,d,dt", _catce_cases(7, 32))
def test_catce_random_shapes(K, B, C, d, dt):
    import mmvae_b200.ops as ops
    g = torch.Generator().manual_seed(K * 1000003 + B * 7919 + C * 31 + d)
    rows = K * B
    shape = (C, d) if d > 1 else (C,)
    x = torch.randn(rows, *shape, generator=g).to(dt)
    t = torch.rand(B, *shape, generator=g)
    t = t / t.sum(1, keepdim=True)
    w = torch.randn(rows, generator=g)
    xo = x.double().requires_grad_(True)
    ref = refmath.lpx_rows("category_ce", xo, t.double(), 1.3, K)
    (ref * w.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.catce_rows(xc, t.cuda(), 1.3)
    (out * w.cuda()).sum().backward()
    vt, gt = (2e-5, 2e-5) if dt == torch.float32 else (2e-4, 1.5e-2)
    assert _rel(out, ref) < vt
    assert _rel(xc.grad, xo.grad) < gt
    xc2 = x.cuda().requires_grad_(True)
    S, _ = ops.catce_weighted_sum(xc2, t.cuda(), 1.3, w_rows=w.cuda())
    S.backward()
    assert _rel(xc2.grad, xo.grad) < gt
