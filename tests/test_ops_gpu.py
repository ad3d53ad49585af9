"""GPU parity: every C-ABI kernel (through mmvae_b200.ops) against the oracle restatement on the same seeded inputs.
Tolerances follow BASELINE.json north_star: 1e-5 relative in fp32, 1e-2 in bf16; index maps bit exact."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import refmath  # noqa: E402

FP32_TOL = 1e-5
BF16_TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.fixture(scope="module")
def ops():
    import mmvae_b200.ops as o
    return o


def _mk(ltype, rows, shape, g, dtype=torch.float32):
    x = torch.randn(rows, *shape, generator=g)
    if ltype in ("bce",):
        x = torch.sigmoid(x).clamp(1e-6, 1 - 1e-6)
    return x.to(dtype)


ELEMENTWISE = [("bce", "normal"), ("lprob", "normal"), ("lprob", "laplace"), ("mse", "normal"), ("l1", "normal")]


@pytest.mark.parametrize("ltype,lik", ELEMENTWISE)
@pytest.mark.parametrize("K,B,shape", [(3, 5, (3, 8, 8)),      # vector path, one CTA per row
                                       (1, 4, (7, 11)),        # P = 77: scalar path
                                       (2, 3, (3, 64, 64)),    # few long rows: split rows + finalize
                                       (1, 2, (4, 9, 64, 64))])  # very long rows
def test_loglik_rows_fp32(ops, ltype, lik, K, B, shape):
    g = torch.Generator().manual_seed(11)
    x = _mk(ltype, K * B, shape, g)
    t = torch.rand(B, *shape, generator=g)
    w = torch.randn(K * B, generator=g)
    lam = 0.37
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows(ltype, xo, t, lam, K, lik)
    (ref * w.to(ref.dtype)).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc, t.cuda(), ltype, lik, lam)
    (out * w.cuda()).sum().backward()
    assert out.shape == ref.shape
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL


@pytest.mark.parametrize("ltype,lik", ELEMENTWISE)
@pytest.mark.parametrize("per_row", [False, True])
def test_loglik_fused_fp32(ops, ltype, lik, per_row):
    g = torch.Generator().manual_seed(12)
    K, B, shape = 2, 6, (3, 16, 16)
    x = _mk(ltype, K * B, shape, g)
    t = torch.rand(B, *shape, generator=g)
    w = torch.randn(K * B, generator=g) if per_row else None
    wc = -0.25
    lam = 1.7
    xo = x.clone().requires_grad_(True)
    ref_rows = refmath.lpx_rows(ltype, xo, t, lam, K, lik)
    wo = w.clone().requires_grad_(True) if per_row else None
    refS = (ref_rows * wo.to(ref_rows.dtype)).sum() if per_row else wc * ref_rows.sum()
    (3.0 * refS).backward()
    xc = x.cuda().requires_grad_(True)
    wcu = w.cuda().requires_grad_(True) if per_row else None
    S, rows = ops.loglik_weighted_sum(xc, t.cuda(), ltype, lik, lam, w_rows=wcu, w_const=wc)
    (3.0 * S).backward()  # non-unit upstream gradient exercises the conditional rescale kernel
    assert rel(rows, ref_rows) < FP32_TOL
    assert rel(S, refS) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL
    if per_row:
        assert rel(wcu.grad, wo.grad) < FP32_TOL


def test_loglik_fused_unit_grad_is_untouched(ops):
    g = torch.Generator().manual_seed(13)
    x = _mk("bce", 8, (3, 8, 8), g)
    t = torch.rand(8, 3, 8, 8, generator=g)
    xo = x.clone().requires_grad_(True)
    (-refmath.lpx_rows("bce", xo, t, 1.0, 1).sum()).backward()
    xc = x.cuda().requires_grad_(True)
    S, _ = ops.loglik_weighted_sum(xc, t.cuda(), "bce", w_const=-1.0)
    S.backward()
    assert rel(xc.grad, xo.grad) < FP32_TOL


def test_bce_clamps(ops):
    """log clamp at -100 and backward denominator clamp 1e-12 (F.binary_cross_entropy semantics)."""
    x = torch.tensor([[0.0, 1.0, 0.5, 1e-30, 1 - 1e-7, 0.25, 0.75, 0.5]])
    t = torch.tensor([[1.0, 0.0, 0.5, 1.0, 0.0, 0.0, 1.0, 1.0]])
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows("bce", xo, t, 1.0, 1)
    ref.sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc, t.cuda(), "bce")
    out.sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL


def test_lprob_nan_to_zero(ops):
    x = torch.tensor([[0.1, float("nan"), 0.3, 0.4]])
    t = torch.tensor([[0.0, 0.5, float("nan"), 1.0]])
    for lik in ("normal", "laplace"):
        xo = x.clone().requires_grad_(True)
        ref = refmath.lpx_rows("lprob", xo, t, 1.0, 1, lik)
        ref.sum().backward()
        xc = x.cuda().requires_grad_(True)
        out = ops.loglik_rows(xc, t.cuda(), "lprob", lik)
        out.sum().backward()
        assert rel(out, ref) < FP32_TOL
        assert torch.equal(torch.isnan(xc.grad.cpu()), torch.isnan(xo.grad))
        m = ~torch.isnan(xo.grad)
        assert rel(xc.grad.cpu()[m], xo.grad[m]) < FP32_TOL


@pytest.mark.parametrize("ltype,lik", [("bce", "normal"), ("lprob", "laplace"), ("mse", "normal")])
@pytest.mark.parametrize("tdtype", [torch.float32, torch.bfloat16])
def test_loglik_bf16(ops, ltype, lik, tdtype):
    g = torch.Generator().manual_seed(14)
    K, B, shape = 2, 4, (3, 16, 16)
    x = _mk(ltype, K * B, shape, g).to(torch.bfloat16)
    t = torch.rand(B, *shape, generator=g).to(tdtype)
    w = torch.randn(K * B, generator=g)
    xo = x.float().requires_grad_(True)
    ref = refmath.lpx_rows(ltype, xo, t.float(), 1.0, K, lik)
    (ref * w.to(ref.dtype)).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc, t.cuda(), ltype, lik)
    (out * w.cuda()).sum().backward()
    assert xc.grad.dtype == torch.bfloat16
    assert rel(out, ref) < 1e-4  # bf16 inputs are exact in fp32; only the accumulation order differs
    assert rel(xc.grad, xo.grad) < BF16_TOL
    xc2 = x.cuda().requires_grad_(True)
    S, _ = ops.loglik_weighted_sum(xc2, t.cuda(), ltype, lik, w_rows=w.cuda())
    S.backward()
    assert rel(xc2.grad, xo.grad) < BF16_TOL


def test_loglik_strided_rows(ops):
    """A mask crop loc[:, :T] (objectives.py:43-45) keeps the row stride: no copy, same numbers."""
    g = torch.Generator().manual_seed(15)
    full = torch.sigmoid(torch.randn(6, 40, generator=g))
    t = torch.rand(6, 24, generator=g)
    xo = full.clone().requires_grad_(True)
    ref = refmath.lpx_rows("bce", xo[:, :24], t, 1.0, 1)
    ref.sum().backward()
    xc = full.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc[:, :24], t.cuda(), "bce")
    out.sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL


@pytest.mark.parametrize("K,B,shape", [(1, 6, (5, 27)), (3, 4, (45, 27)), (2, 5, (9,)), (1, 7, (4, 6)), (1, 3, (246, 27))])
def test_catce(ops, K, B, shape):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(K * B, *shape, generator=g)
    idx = torch.randint(shape[-1], (B, *shape[:-1]), generator=g)
    t = torch.nn.functional.one_hot(idx, shape[-1]).float()
    t[0] = torch.rand(shape, generator=g)  # soft targets too
    w = torch.randn(K * B, generator=g)
    lam = 0.6
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows("category_ce", xo, t, lam, K)
    (ref * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.catce_rows(xc, t.cuda(), lam)
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL
    xc2 = x.cuda().requires_grad_(True)
    S, rows = ops.catce_weighted_sum(xc2, t.cuda(), lam, w_rows=w.cuda())
    (2.0 * S).backward()
    assert rel(S, (ref * w).sum()) < FP32_TOL
    assert rel(xc2.grad, 2.0 * xo.grad) < FP32_TOL


def test_catce_mask_crop(ops):
    g = torch.Generator().manual_seed(22)
    x = torch.randn(5, 12, 27, generator=g)
    t = torch.nn.functional.one_hot(torch.randint(27, (5, 8), generator=g), 27).float()
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows("category_ce", xo, t, 1.0, 1, mask_len=8)
    ref.sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.catce_rows(xc[:, :8], t.cuda())
    out.sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL


@pytest.mark.parametrize("K,B,shape", [(1, 6, (3, 4, 4)), (3, 5, (400,)), (1, 64, (3, 64, 64))])
def test_optimal_sigma(ops, K, B, shape):
    g = torch.Generator().manual_seed(31)
    x = torch.randn(K * B, *shape, generator=g)
    t = torch.rand(B, *shape, generator=g)
    w = torch.randn(K * B, generator=g)
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows("optimal_sigma", xo, t, 0.9, K)
    (ref * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.osigma_rows(xc, t.cuda(), 0.9)
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < 5e-5  # gradient = scalar (sum_r w_r, cancellation prone) x (x - t)


def _prior(D, g):
    logits = (torch.randn(1, D, generator=g) * 0.3)
    return logits


@pytest.mark.parametrize("M,B,D", [(2, 7, 16), (3, 33, 10), (3, 5, 64), (4, 3, 100)])
def test_latent_draws_poe_subsets(ops, M, B, D):
    """All 2^M-1 PoE subsets with the prior expert, K=1 sample each, KL vs the learnable prior (MVAE)."""
    from mmvae_b200.ops import Draw
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(41)
    post = [syn.make_posterior(g, B, D) for _ in range(M)]
    logits = _prior(D, g)
    subsets = refmath.poe_subsets(range(M))
    eps = [torch.randn(1, B, D, generator=g) for _ in subsets]
    wz = [torch.randn(1, B, D, generator=g) for _ in subsets]
    wk = [torch.randn(B, generator=g) for _ in subsets]

    def run(dev):
        mus = [p[0].detach().clone().to(dev).requires_grad_(True) for p in post]
        ss = [p[1].detach().clone().to(dev).requires_grad_(True) for p in post]
        lg = logits.detach().clone().to(dev).requires_grad_(True)
        mu0 = torch.zeros_like(lg)
        s0 = torch.softmax(lg, 1) * D
        tot = 0
        outs = []
        if dev == "cpu":
            mods = [dict(mu=m, s=s) for m, s in zip(mus, ss)]
            for a, sub in enumerate(subsets):
                loc, var = refmath.poe_mixing(mods, set(sub), B, D)
                z = refmath.normal_rsample(loc, var, eps[a])
                kl = refmath.kl_normal_normal(loc, var, mu0, s0).sum(-1)
                outs.append((loc, var, z, kl))
        else:
            draws = [Draw(mods=sub, prior=True, kl_mode=1, col0=0, width=D, K=1, want_params=True) for sub in subsets]
            res = ops.latent_draws(torch.stack(mus), torch.stack(ss), mu0, s0,
                                   torch.cat([e.reshape(-1) for e in eps]).to(dev), draws)
            outs = [(r["loc"], r["scale"], r["z"], r["kl"]) for r in res]
        for a, (loc, var, z, kl) in enumerate(outs):
            tot = tot + (z * wz[a].to(dev)).sum() + (kl * wk[a].to(dev)).sum() + (loc * 0.3).sum() + (var * 0.7).sum()
        tot.backward()
        return outs, mus, ss, lg

    o_ref, mu_r, s_r, lg_r = run("cpu")
    o_gpu, mu_g, s_g, lg_g = run("cuda")
    for a in range(len(subsets)):
        for x, y in zip(o_gpu[a], o_ref[a]):
            assert rel(x, y) < FP32_TOL
    for m in range(M):
        assert rel(mu_g[m].grad, mu_r[m].grad) < FP32_TOL
        assert rel(s_g[m].grad, s_r[m].grad) < FP32_TOL
    assert rel(lg_g.grad, lg_r.grad) < FP32_TOL


def test_latent_draws_direct_and_private(ops):
    """DMVAE style: direct draws on shared / private column slices, Laplace direct draws, KL vs N(0,1)."""
    from mmvae_b200.ops import Draw
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(42)
    M, B, D, Pv, K = 2, 9, 16, 10, 3
    post = [syn.make_posterior(g, B, D + Pv) for _ in range(M)]
    logits = _prior(D, g)
    eps_sh = torch.randn(K, B, D, generator=g)
    eps_pr = torch.randn(K, B, Pv, generator=g)
    u_lap = syn.make_noise(g, "laplace", (K, B, D))
    eps_j = torch.randn(K, B, D, generator=g)
    wts = [torch.randn(K, B, D, generator=g), torch.randn(K, B, Pv, generator=g), torch.randn(K, B, D, generator=g),
           torch.randn(K, B, D, generator=g)]
    wk = [torch.randn(B, generator=g) for _ in range(4)]

    def run(dev):
        mus = [p[0].detach().clone().to(dev).requires_grad_(True) for p in post]
        ss = [p[1].detach().clone().to(dev).requires_grad_(True) for p in post]
        lg = logits.detach().clone().to(dev).requires_grad_(True)
        mu0, s0 = torch.zeros_like(lg), torch.softmax(lg, 1) * D
        if dev == "cpu":
            sh, pr = (mus[0][:, :D], ss[0][:, :D]), (mus[1][:, D:], ss[1][:, D:])
            z0 = refmath.normal_rsample(*sh, eps_sh)
            k0 = refmath.kl_normal_normal(*sh, mu0, s0).sum(-1)
            z1 = refmath.normal_rsample(*pr, eps_pr)
            k1 = refmath.kl_normal_normal(*pr, torch.zeros(1, Pv), torch.ones(1, Pv)).sum(-1)
            z2 = refmath.laplace_rsample(mus[1][:, :D], ss[1][:, :D], u_lap)
            k2 = refmath.kl_laplace_normal(mus[1][:, :D], ss[1][:, :D], torch.zeros(1, D), torch.ones(1, D)).sum(-1)
            lj, vj = refmath.product_of_experts(torch.stack([m[:, :D] for m in mus]), torch.stack([s[:, :D] for s in ss]))
            z3 = refmath.normal_rsample(lj, vj, eps_j)
            k3 = refmath.kl_normal_normal(lj, vj, mu0, s0).sum(-1)
            outs = [(z0, k0), (z1, k1), (z2, k2), (z3, k3)]
        else:
            draws = [Draw(mods=(0,), direct=True, kl_mode=1, col0=0, width=D, K=K),
                     Draw(mods=(1,), direct=True, kl_mode=2, col0=D, width=Pv, K=K),
                     Draw(mods=(1,), direct=True, laplace=True, kl_mode=2, col0=0, width=D, K=K),
                     Draw(mods=(0, 1), prior=False, kl_mode=1, col0=0, width=D, K=K)]
            eps = torch.cat([e.reshape(-1) for e in (eps_sh, eps_pr, u_lap, eps_j)]).to(dev)
            res = ops.latent_draws(torch.stack(mus), torch.stack(ss), mu0, s0, eps, draws)
            outs = [(r["z"], r["kl"]) for r in res]
        tot = 0
        for i, (z, k) in enumerate(outs):
            tot = tot + (z * wts[i].to(dev)).sum() + (k * wk[i].to(dev)).sum()
        tot.backward()
        return outs, mus, ss, lg

    o_ref, mu_r, s_r, lg_r = run("cpu")
    o_gpu, mu_g, s_g, lg_g = run("cuda")
    for (za, ka), (zb, kb) in zip(o_gpu, o_ref):
        assert rel(za, zb) < FP32_TOL and rel(ka, kb) < FP32_TOL
    for m in range(M):
        assert rel(mu_g[m].grad, mu_r[m].grad) < FP32_TOL
        assert rel(s_g[m].grad, s_r[m].grad) < FP32_TOL
    assert rel(lg_g.grad, lg_r.grad) < FP32_TOL


@pytest.mark.parametrize("S,B", [(3, 8), (7, 16), (7, 100), (15, 33)])
def test_latent_draws_rowmask_selection(ops, S, B):
    """mixture_component_selection as a row -> subset map (function-level contract of mmvae_models.py:396-410):
    the map is bit exact and the selected rows equal the oracle's torch.cat of chunk slices."""
    from mmvae_b200.ops import Draw
    from mmvae_b200.mmvae_models import mopoe_row_subset_map, subset_bitmasks
    import mmvae_b200.synthetic as syn
    M = int(math.log2(S + 1))
    D = 12
    g = torch.Generator().manual_seed(43)
    post = [syn.make_posterior(g, B, D) for _ in range(M)]
    subs = refmath.mopoe_subsets(range(M))
    row_map = mopoe_row_subset_map(S, B)
    assert torch.equal(row_map, refmath.mopoe_row_to_subset(S, B))
    mods = [dict(mu=p[0], s=p[1]) for p in post]
    mus, lvs = [], []
    for sub in subs:
        mu = torch.stack([mods[i]["mu"] for i in sub])
        lv = torch.stack([mods[i]["s"] for i in sub])
        if len(sub) == M:
            mu = torch.cat((mu, torch.zeros(1, B, D)), 0)
            lv = torch.cat((lv, torch.zeros(1, B, D)), 0)
        a, b = refmath.product_of_experts(mu, lv)
        mus.append(a)
        lvs.append(b)
    w = torch.ones(S) / S
    mu_sel, var_sel = refmath.mixture_component_selection(torch.stack(mus), torch.stack(lvs), w)
    masks = subset_bitmasks(subs, M)  # bit 31 = prior for the full subset
    row_masks = masks[row_map.long()].cuda()
    res = ops.latent_draws(torch.stack([p[0] for p in post]).cuda(), torch.stack([p[1] for p in post]).cuda(), None,
                           None, None, [Draw(rowmask=True, width=D, want_params=True)], row_masks)
    assert rel(res[0]["loc"], mu_sel) < FP32_TOL
    assert rel(res[0]["scale"], var_sel) < FP32_TOL


@pytest.mark.parametrize("M,B,D,K,dists", [(2, 6, 16, 3, ("normal", "normal")), (2, 5, 64, 4, ("laplace", "laplace")),
                                           (3, 4, 10, 2, ("normal", "laplace", "normal")), (2, 3, 100, 9, ("normal", "normal"))])
@pytest.mark.parametrize("through_z", [True, False])
def test_moe_logdens(ops, M, B, D, K, dists, through_z):
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(51)
    post = [syn.make_posterior(g, B, D) for _ in range(M)]
    logits = _prior(D, g)
    noise = [syn.make_noise(g, dists[m], (K, B, D)) for m in range(M)]
    w_z = torch.randn(M, K, B, D, generator=g)
    w_lq = torch.randn(M, M, K, B, generator=g)
    w_lp = torch.randn(M, K, B, generator=g)

    def run(dev):
        mus = [p[0].detach().clone().to(dev).requires_grad_(True) for p in post]
        ss = [p[1].detach().clone().to(dev).requires_grad_(True) for p in post]
        lg = logits.detach().clone().to(dev).requires_grad_(True)
        mu0, s0 = torch.zeros_like(lg), torch.softmax(lg, 1) * D
        if dev == "cpu":
            z = torch.stack([refmath.rsample(dists[m], mus[m], ss[m], noise[m]) for m in range(M)])
            zin = z if through_z else z.detach()
            lq = torch.stack([torch.stack([refmath.log_prob(dists[j], zin[r], mus[j], ss[j]).sum(-1) for j in range(M)])
                              for r in range(M)])
            lpz = torch.stack([refmath.normal_log_prob(zin[r], mu0, s0).sum(-1) for r in range(M)])
        else:
            code = [1 if d == "laplace" else 0 for d in dists]
            z, lq, lpz = ops.moe_logdens(torch.stack(mus), torch.stack(ss), mu0, s0, torch.stack(noise).to(dev), code,
                                         through_z)
        tot = (z * w_z.to(dev)).sum() + (lq * w_lq.to(dev)).sum() + (lpz * w_lp.to(dev)).sum()
        tot.backward()
        return (z, lq, lpz), mus, ss, lg

    o_ref, mu_r, s_r, lg_r = run("cpu")
    o_gpu, mu_g, s_g, lg_g = run("cuda")
    for x, y in zip(o_gpu, o_ref):
        assert rel(x, y) < FP32_TOL
    for m in range(M):
        assert rel(mu_g[m].grad, mu_r[m].grad) < 2e-5
        assert rel(s_g[m].grad, s_r[m].grad) < 2e-5
    assert rel(lg_g.grad, lg_r.grad) < 2e-5


@pytest.mark.parametrize("M,L,K,B", [(2, 2, 30, 37), (3, 2, 5, 64), (2, 1, 1, 5)])
def test_iwae_combine(ops, M, L, K, B):
    g = torch.Generator().manual_seed(61)
    lpz = torch.randn(M, K, B, generator=g) * 3
    lq = torch.randn(M, M, K, B, generator=g) * 3
    lpx = torch.randn(M, L, K, B, generator=g) * 30
    beta = 1.3

    def run(dev):
        a, b, c = (t.detach().clone().to(dev).requires_grad_(True) for t in (lpz, lq, lpx))
        if dev == "cpu":
            lws = [a[r] + c[r].sum(0) - beta * refmath.log_mean_exp(b[r]) for r in range(M)]
            loss = -refmath.log_mean_exp(torch.cat(lws)).sum()
        else:
            loss, _ = ops.iwae_combine(a, b, c, beta)
        (2.0 * loss).backward()
        return loss, a.grad, b.grad, c.grad

    r, g_ = run("cpu"), run("cuda")
    for x, y in zip(g_, r):
        assert rel(x, y) < FP32_TOL


@pytest.mark.parametrize("M,L,K,B", [(2, 2, 50, 33), (3, 2, 4, 5000), (2, 2, 7, 1)])
def test_dreg_combine(ops, M, L, K, B):
    g = torch.Generator().manual_seed(62)
    lpz = torch.randn(M, K, B, generator=g)
    lq = torch.randn(M, M, K, B, generator=g)
    lpx = torch.randn(M, L, K, B, generator=g)

    def run(dev):
        a, b, c = (t.detach().clone().to(dev).requires_grad_(True) for t in (lpz, lq, lpx))
        if dev == "cpu":
            lw = torch.stack([a[r].sum(-1) + c[r].sum(0).sum(-1) - refmath.log_mean_exp(b[r]).sum(-1) for r in range(M)])
            with torch.no_grad():
                wt = (lw - torch.logsumexp(lw, 1, keepdim=True)).exp()
            loss = -(wt * lw).mean(0).sum()
        else:
            loss, _ = ops.dreg_combine(a, b, c)
        loss.backward()
        return loss, a.grad, b.grad, c.grad

    r, g_ = run("cpu"), run("cuda")
    for x, y in zip(g_, r):
        assert rel(x, y) < 5e-5  # batch sums of O(B) terms in fp32 feed a softmax


@pytest.mark.parametrize("lik", ["normal", "laplace"])
def test_lprob_selfscale(ops, lik):
    """lprob with padding masks: the reference overwrites the likelihood scale with the cropped loc (objectives.py:43-45),
    i.e. dist(loc = x, scale = x).log_prob(t); a negative mean gives log(negative) = NaN -> value 0 (:423) and, through
    autograd's zeroed upstream gradient, gradient 0."""
    import torch.distributions as dist
    g = torch.Generator().manual_seed(19)
    K, B, shape = 2, 5, (6, 8)
    x = torch.randn(K * B, *shape, generator=g)
    x = torch.where(x.abs() < 0.05, torch.full_like(x, 0.3), x)  # keep 1/x^2 gradients well conditioned
    t = torch.rand(B, *shape, generator=g)
    w = torch.randn(K * B, generator=g)
    xo = x.double().clone().requires_grad_(True)
    d = (dist.Laplace if lik == "laplace" else dist.Normal)(xo, torch.tensor(0.75, dtype=torch.float64), validate_args=False)
    d.scale = xo
    out = d.log_prob(t.double().repeat(K, 1, 1)).view(K * B, -1)
    assert torch.isnan(out).any()
    out = out.clone()
    out[torch.isnan(out)] = 0
    ref = 0.37 * out.sum(-1)
    (ref * w.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    rows = ops.loglik_rows(xc, t.cuda(), "lprob_selfscale", lik, 0.37)
    (rows * w.cuda()).sum().backward()
    assert rel(rows, ref) < FP32_TOL
    assert not torch.isnan(xc.grad).any()
    assert rel(xc.grad, xo.grad) < FP32_TOL
    # fused (ELBO) pass gives the same rows and gradient
    xf = x.cuda().requires_grad_(True)
    S, rows_f = ops.loglik_weighted_sum(xf, t.cuda(), "lprob_selfscale", lik, 0.37, w_rows=w.cuda())
    S.backward()
    assert rel(rows_f, ref) < FP32_TOL and rel(xf.grad, xo.grad) < FP32_TOL


def test_ops_refuse_cpu_tensors(ops):
    with pytest.raises(RuntimeError):
        ops.loglik_rows(torch.rand(4, 8), torch.rand(4, 8), "bce")


def test_loglik_l2_tiled_row_order(ops):
    """Targets larger than the L2 tile switch the launch to a (b-tile, k, b) row order: same numbers, any K."""
    g = torch.Generator().manual_seed(16)
    K, B, P = 3, 5000, 1024  # B*P*4 = 20 MB > 16 MB tile -> tiled order with a short last tile
    x = torch.sigmoid(torch.randn(K * B, P, generator=g))
    t = torch.rand(B, P, generator=g)
    w = torch.randn(K * B, generator=g)
    xo = x.clone().requires_grad_(True)
    ref = refmath.lpx_rows("bce", xo, t, 1.0, K)
    (ref * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.loglik_rows(xc, t.cuda(), "bce")
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < FP32_TOL
    assert rel(xc.grad, xo.grad) < FP32_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_bce_logits_fused_decoder_tail(ops, dtype, tol):
    """bce_logits == bce(clamp(sigmoid(y), 1e-6, 1-1e-6)) including the clamp's gradient mask (decoders.py:96-97)."""
    g = torch.Generator().manual_seed(17)
    K, B, shape = 2, 5, (3, 16, 16)
    y = (torch.randn(K * B, *shape, generator=g) * 4)
    y.view(-1)[:8] = torch.tensor([-30.0, 30.0, -13.9, 13.9, -13.7, 13.7, 0.0, -0.0])  # both sides of the clamp
    y = y.to(dtype)
    t = torch.rand(B, *shape, generator=g)
    w = torch.randn(K * B, generator=g)
    # The reference formulation log(1 - sigmoid(y)) cancels catastrophically in fp32 for saturated logits (up to 1e-4
    # absolute per element); the fused kernel evaluates log(1-s) directly.  So the check is against the oracle in
    # fp64, and the fp32 oracle must itself be within its own error band of the kernel.
    yo = y.double().requires_grad_(True)
    ref = refmath.lpx_rows("bce_logits", yo, t.double(), 0.8, K)
    (ref * w.double()).sum().backward()
    ref32 = refmath.lpx_rows("bce_logits", y.float(), t, 0.8, K)
    yc = y.cuda().requires_grad_(True)
    out = ops.loglik_rows(yc, t.cuda(), "bce_logits", "normal", 0.8)
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < (1e-5 if dtype == torch.float32 else 1e-4)
    assert rel(out, ref32) < 1e-4 and rel(out, ref) <= rel(ref32, ref) + 1e-6  # at least as close to the truth
    assert rel(yc.grad, yo.grad) < tol
    yc2 = y.cuda().requires_grad_(True)
    S, _ = ops.loglik_weighted_sum(yc2, t.cuda(), "bce_logits", "normal", 0.8, w_rows=w.cuda())
    S.backward()
    assert rel(yc2.grad, yo.grad) < tol


def test_prior_scale_and_iwae_rows(ops):
    """The folded glue kernels: softmax*D prior scale (fwd/bwd) and the pointer-table IWAE combine + one-launch bwd."""
    g = torch.Generator().manual_seed(63)
    lg = torch.randn(1, 37, generator=g)
    a = lg.clone().requires_grad_(True)
    ref = torch.softmax(a, 1) * 37
    w = torch.randn(1, 37, generator=g)
    (ref * w).sum().backward()
    b = lg.cuda().requires_grad_(True)
    out = ops.prior_scale(b)
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < FP32_TOL and rel(b.grad, a.grad) < FP32_TOL
    M, L, K, B = 2, 2, 7, 45
    lpz = torch.randn(M, K, B, generator=g) * 3
    lq = torch.randn(M, M, K, B, generator=g) * 3
    rows = [torch.randn(K * B, generator=g) * 30 for _ in range(M * L)]
    x = [t.cuda().requires_grad_(True) for t in [lpz, lq] + rows]
    loss, _ = ops.iwae_combine(x[0], x[1], torch.stack(x[2:]).view(M, L, K, B), 0.7)
    (1.5 * loss).backward()
    y = [t.cuda().requires_grad_(True) for t in [lpz, lq] + rows]
    loss2, _ = ops.iwae_combine_rows(y[0], y[1], y[2:], L, 0.7)
    (1.5 * loss2).backward()
    assert torch.equal(loss, loss2)
    for p, q in zip(x, y):
        assert rel(q.grad, p.grad) < 1e-6


@pytest.mark.parametrize("laplace", [False, True])
def test_kl_elementwise_and_objective_api(ops, laplace):
    """calc_kld / weighted_group_kld of the objective plugin (reference objectives.py:148-201) on the element-wise KL
    kernel, against torch.distributions.kl (what utils.kl_divergence dispatches to)."""
    import torch.distributions as dist
    import mmvae_b200
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(71)
    B, D = 37, 16
    mu, s = syn.make_posterior(g, B, D)
    lg = torch.randn(1, D, generator=g) * 0.3
    w = torch.randn(B, D, generator=g)
    Q = dist.Laplace if laplace else dist.Normal

    def run(dev):
        a, b, c = mu.clone().to(dev).requires_grad_(True), s.clone().to(dev).requires_grad_(True), lg.clone().to(dev).requires_grad_(True)
        prior = dist.Normal(torch.zeros(1, D, device=dev), torch.softmax(c, 1) * D)
        if dev == "cpu":
            kl = dist.kl_divergence(Q(a, b), prior)
        else:
            kl = mmvae_b200.MultimodalObjective("elbo").calc_kld(Q(a, b), prior)
        (kl * w.to(dev)).sum().backward()
        return kl, a.grad, b.grad, c.grad
    r, o = run("cpu"), run("cuda")
    for x, y in zip(o, r):
        assert rel(x, y) < FP32_TOL


@pytest.mark.parametrize("B,unit", [(5, False), (1000, True), (4099, False)])
def test_elbo_combine(ops, B, unit):
    """mmvae_objective_elbo against the eager spelling of reference objectives.py:54-67 / mmvae_models.py:181-187:
    sum of (deferred and already reduced) likelihood terms + weighted KL row sums, the logged kld, every gradient."""
    g = torch.Generator().manual_seed(5)
    dev = "cuda"
    x = [torch.sigmoid(torch.randn(B, 3, 8, 8, generator=g)).clamp(1e-6, 1 - 1e-6) for _ in range(2)]
    c = torch.randn(B, 7, 27, generator=g)
    t_img, t_txt = torch.rand(B, 3, 8, 8, generator=g), torch.softmax(torch.randn(B, 7, 27, generator=g), 1)
    kl = (torch.rand(3 * B, generator=g) * 4).to(dev).requires_grad_(True)
    extra = torch.randn((), generator=g).to(dev).requires_grad_(True)  # an already reduced term (coefficient 1)
    kc, kg = [0.7, 1.3, 2.0], [1 / 3.0, 0.0, 0.5]
    xs = [a.to(dev).requires_grad_(True) for a in x]
    cs = c.to(dev).requires_grad_(True)
    terms = [ops.loglik_weighted_sum(xs[0], t_img.to(dev), "bce", "normal", 0.4, w_const=-1.0, defer=True)[0],
             ops.loglik_weighted_sum(xs[1], t_img.to(dev), "bce", "normal", 1.0, w_const=-0.25, defer=True)[0],
             ops.catce_weighted_sum(cs, t_txt.to(dev), 2.0, w_const=-1.0, defer=True)[0], extra]
    loss, kld = ops.elbo_combine(terms, kl, kc, kg)
    up = ops.mark_unit_grad(torch.ones((), device=dev)) if unit else torch.tensor(0.37, device=dev)
    loss.backward(up)
    # eager reference with the oracle's row functions
    xo = [a.clone().requires_grad_(True) for a in x]
    co = c.clone().requires_grad_(True)
    klo = kl.detach().cpu().clone().requires_grad_(True)
    eo = extra.detach().cpu().clone().requires_grad_(True)
    ref = (-refmath.lpx_rows("bce", xo[0], t_img, 0.4, 1, "normal").sum()
           - 0.25 * refmath.lpx_rows("bce", xo[1], t_img, 1.0, 1, "normal").sum()
           - refmath.lpx_rows("category_ce", co, t_txt, 2.0, 1, "normal").sum() + eo
           + sum(kc[j] * klo[j * B:(j + 1) * B].sum() for j in range(3)))
    ref_kld = sum(kg[j] * klo[j * B:(j + 1) * B].sum() for j in range(3))
    (ref * float(up)).backward()
    assert rel(loss, ref) < FP32_TOL and rel(kld, ref_kld) < FP32_TOL
    assert rel(kl.grad, klo.grad) < FP32_TOL and rel(extra.grad, eo.grad) < FP32_TOL
    for a, b in zip(xs + [cs], xo + [co]):
        assert rel(a.grad, b.grad) < FP32_TOL
    if unit:
        ops.unmark_unit_grad(up)


def test_elbo_combine_limits(ops):
    z = torch.zeros((), device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError):
        ops.elbo_combine([z] * 49)
    loss, _ = ops.elbo_combine([z + 1.5, z + 2.0])  # no KL segments at all
    assert abs(float(loss) - 3.5) < 1e-6


@pytest.mark.parametrize("M,B,D,K,dists", [(2, 37, 16, 3, ("normal", "normal")), (2, 9, 64, 4, ("laplace", "laplace")),
                                           (3, 5, 32, 2, ("normal", "laplace", "normal"))])
def test_moe_logdens_fused_encoder_tail(ops, M, B, D, K, dists):
    """mmvae_moe_logdens_{fwd,bwd}_tail: raw second-head logits in, s = softmax(raw, -1) + 1e-6 (reference
    encoders.py:49-54) evaluated inside the kernel, gradient with respect to the raw logits -- against the same kernels
    fed with the torch tail (which test_moe_logdens pins to the oracle)."""
    g = torch.Generator().manual_seed(21)
    dev = "cuda"
    mu = torch.randn(M, B, D, generator=g)
    raw = torch.randn(M, B, D, generator=g) * 1.5
    mu0, s0 = torch.randn(1, D, generator=g) * 0.1, torch.rand(1, D, generator=g) + 0.5
    eps = torch.stack([torch.randn(K, B, D, generator=g) if d == "normal" else torch.rand(K, B, D, generator=g) * 1.98 - 0.99
                       for d in dists])
    codes = [1 if d == "laplace" else 0 for d in dists]
    wz, wq, wp = torch.randn(M, K, B, D, generator=g), torch.randn(M, M, K, B, generator=g), torch.randn(M, K, B, generator=g)

    def run(fused):
        m_, r_ = mu.to(dev).requires_grad_(True), raw.to(dev).requires_grad_(True)
        p0, p1 = mu0.to(dev).requires_grad_(True), s0.to(dev).requires_grad_(True)
        if fused:
            z, lq, lpz, sc = ops.moe_logdens_tail(m_, r_, p0, p1, eps.to(dev), codes)
        else:
            sc = torch.softmax(r_, -1) + 1e-6
            z, lq, lpz = ops.moe_logdens(m_, sc, p0, p1, eps.to(dev), codes)
        ((z * wz.to(dev)).sum() + (lq * wq.to(dev)).sum() + (lpz * wp.to(dev)).sum()).backward()
        return [z, lq, lpz, sc.detach(), m_.grad, r_.grad, p0.grad, p1.grad]

    for a, b in zip(run(True), run(False)):
        assert rel(a, b) < FP32_TOL


@pytest.mark.parametrize("M,B,D,pv", [(2, 33, 16, 10), (3, 7, 10, 0)])
def test_latent_draws_fused_encoder_tail(ops, M, B, D, pv):
    """mmvae_latent_draws_{fwd,bwd}_tail (PoE subsets with the prior expert, direct shared / private draws, KL rows)
    against the same kernels fed with the torch encoder tail."""
    g = torch.Generator().manual_seed(22)
    dev = "cuda"
    Dt = D + pv
    mu, raw = torch.randn(M, B, Dt, generator=g), torch.randn(M, B, Dt, generator=g) * 1.5
    mu0, s0 = torch.zeros(1, D), torch.rand(1, D, generator=g) + 0.5
    draws = [ops.Draw(mods=tuple(range(M)), prior=True, kl_mode=1, width=D, K=2, want_params=True),
             ops.Draw(mods=(0,), direct=True, kl_mode=1, width=D, K=1),
             ops.Draw(mods=(M - 1,), prior=True, kl_mode=1, width=D, K=1)]
    if pv:
        draws.append(ops.Draw(mods=(1,), direct=True, kl_mode=2, col0=D, width=pv, K=1))
    eps = torch.cat([torch.randn(d.K * B * d.width, generator=g) for d in draws])
    wz = torch.randn(eps.numel(), generator=g)

    def run(fused):
        m_, r_ = mu.to(dev).requires_grad_(True), raw.to(dev).requires_grad_(True)
        p1 = s0.to(dev).requires_grad_(True)
        s_in = r_ if fused else torch.softmax(r_, -1) + 1e-6
        res = ops.latent_draws(m_, s_in, mu0.to(dev), p1, eps.to(dev), draws, s_raw=fused)
        zs = torch.cat([r["z"].reshape(-1) for r in res])
        loss = (zs * wz.to(dev)).sum() + 0.7 * res.kl_packed.sum() + (res[0]["loc"] * 0.3).sum() + (res[0]["scale"] * 1.1).sum()
        loss.backward()
        sc = res.scales if fused else s_in.detach()
        return [zs, res.kl_packed, res[0]["loc"], res[0]["scale"], sc, m_.grad, r_.grad, p1.grad]

    for a, b in zip(run(True), run(False)):
        assert rel(a, b) < FP32_TOL


@pytest.mark.parametrize("K,B,shape", [(1, 8, (246, 27)), (2, 4, (128, 27)), (1, 8, (64, 6)), (3, 4, (100, 40)),
                                       (1, 4, (250, 31)), (1, 4, (512, 27)), (1, 4, (600, 27))])
def test_catce_bf16_long_class_axes(ops, K, B, shape):
    """bf16 reconstructions AND targets on long class axes: the register-resident kernel (<= 16 periods per warp,
    d = 27 compile-time and generic), the chunked pair kernel behind it (C = 600: 300 periods do not fit 16 warps) --
    forward rows + cached-statistics backward, and the fused value + gradient pass, against the oracle on the same
    bf16-rounded inputs."""
    g = torch.Generator().manual_seed(23)
    x = torch.randn(K * B, *shape, generator=g).bfloat16()
    idx = torch.randint(shape[-1], (B, *shape[:-1]), generator=g)
    t = torch.nn.functional.one_hot(idx, shape[-1]).float()
    t[:, shape[0] // 2:] = 0  # padded positions
    t[0] = torch.rand(shape, generator=g)
    t = t.bfloat16()
    w = torch.randn(K * B, generator=g)
    lam = 0.6
    xo = x.float().clone().requires_grad_(True)
    ref = refmath.lpx_rows("category_ce", xo, t.float(), lam, K)
    (ref * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.catce_rows(xc, t.cuda(), lam)
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < 1e-5  # fp32 accumulation over bf16 inputs: the row values are fp32 quantities
    assert rel(xc.grad, xo.grad) < BF16_TOL
    xc2 = x.cuda().requires_grad_(True)
    S, rows = ops.catce_weighted_sum(xc2, t.cuda(), lam, w_rows=w.cuda())
    (2.0 * S).backward()
    assert rel(S, (ref * w).sum()) < 1e-5 and rel(rows, ref) < 1e-5
    assert rel(xc2.grad, 2.0 * xo.grad) < BF16_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("K,B,shape", [(1, 6, (45, 27)), (2, 5, (7, 27)), (1, 4, (12, 40))])
def test_catce_fused_text_decoder_mask(ops, dtype, tol, K, B, shape):
    """mmvae_catce_rows_masked: the text decoder's "zero for padded area" multiply (reference decoders.py:722) inside the
    category_ce kernel -- forward rows, backward, fused value + gradient -- against the oracle fed with output * mask."""
    g = torch.Generator().manual_seed(31)
    T, d = shape
    x = torch.randn(K * B, T, d, generator=g).to(dtype)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    mask = torch.arange(T)[None, :] < lens[:, None]  # (B, T) bool
    t = torch.nn.functional.one_hot(torch.randint(d, (B, T), generator=g), d).float() * mask[..., None]
    w = torch.randn(K * B, generator=g)
    lam = 0.8
    xo = x.float().clone().requires_grad_(True)
    xm = xo * mask.repeat(K, 1)[..., None].float()
    ref = refmath.lpx_rows("category_ce", xm, t, lam, K)
    (ref * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = ops.catce_rows(xc, t.cuda(), lam, mask=mask.cuda())
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) < 1e-5
    assert rel(xc.grad, xo.grad) < tol
    assert float(xc.grad[~mask.repeat(K, 1).cuda()].abs().max()) == 0.0  # nothing flows into padded positions
    xc2 = x.cuda().requires_grad_(True)
    S, rows = ops.catce_weighted_sum(xc2, t.cuda(), lam, w_rows=w.cuda(), mask=mask.cuda())
    S.backward()
    assert rel(S, (ref * w).sum()) < 1e-5 and rel(xc2.grad, xo.grad) < tol
