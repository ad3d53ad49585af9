"""TEST INFRASTRUCTURE (build container only): pin oracle/refmath.py to the UNMODIFIED reference.

Runs every case of oracle/cases.py through the reference classes imported in place from /root/reference
(models.moe / poe / mopoe / dmvae -> objective(batch), with stand-in VAEs and injected noise) and through the
restatement, and asserts agreement of loss, kld, reconstruction terms and every gradient.  Also checks the
function-level pieces (product_of_experts, ReconLoss.*, mixture_component_selection bounds, kl formulas).

    PYTHONHASHSEED=0 python -m oracle.validate_against_reference
"""
import itertools
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from oracle import cases, ref_inplace, refmath  # noqa: E402

RTOL = 2e-6  # two fp32 CPU evaluations of the same formulas; differences are summation order only


def reference_poe_subsets(utils, batch):
    """The order subsample_input_modalities (utils.py:86-112) actually produced in this process."""
    subs = []
    for mi in utils.subsample_input_modalities(batch):
        subs.append(tuple(i for i, k in enumerate(batch.keys()) if mi[k]["data"] is not None))
    return subs


def run_reference(case):
    models, objectives, utils = ref_inplace.load()
    vaes = cases.build_vaes(case)
    cls = getattr(models, case["model"])
    model = cls(vaes, case["D"], {"obj": case["obj"], "beta": case["beta"], "K": case["K"]}, None)
    with torch.no_grad():
        model._pz_params[1].copy_(case["pz_logits"])
    if case["obj"] == "iwae":  # N1 shim: tuple with a .cuda() method, arithmetic untouched
        base_prop = cls.pz_params
        patched = type(cls.__name__ + "N1", (cls,), {
            "pz_params": property(lambda self: ref_inplace.PzParams(base_prop.fget(self)))})
        model.__class__ = patched
    batch = cases.build_batch(case)
    subsets = reference_poe_subsets(utils, batch) if case["model"] == "poe" else None
    with ref_inplace.NoiseInjector([n.clone() for n in case["noise"]]) as inj:
        out = model.objective(batch)
        assert not inj.noises, "reference consumed %d fewer noise tensors than planned" % len(inj.noises)
    out["loss"].backward()
    leaves = cases.named_leaves(vaes, model._pz_params[1])
    res = cases.collect(out, leaves)
    return res, subsets


def compare(a, b, name, rtol=RTOL):
    worst = 0.0
    for k in sorted(set(a) | set(b)):
        if k.startswith("_"):
            continue
        if k not in a or k not in b:
            if k == "reconstruction_loss" or k == "kld":
                continue
            raise AssertionError("%s: key %s missing on one side" % (name, k))
        x, y = a[k], b[k]
        if x is None or y is None:
            zx = x is None or float(x.abs().max()) == 0.0
            zy = y is None or float(y.abs().max()) == 0.0
            assert zx and zy, "%s: %s is None on one side but non-zero on the other" % (name, k)
            continue
        assert x.shape == y.shape, "%s: %s shape %s vs %s" % (name, k, tuple(x.shape), tuple(y.shape))
        denom = max(float(x.abs().max()), 1e-12)
        err = float((x - y).abs().max()) / denom
        worst = max(worst, err)
        assert err <= rtol, "%s: %s differs rel %.3e" % (name, k, err)
    return worst


def check_functions():
    models, objectives, utils = ref_inplace.load()
    g = torch.Generator().manual_seed(7)
    # a1 product_of_experts
    mu, lv = torch.randn(4, 9, 5, generator=g), torch.rand(4, 9, 5, generator=g)
    ra = models.mmvae_base.TorchMMVAE.product_of_experts(mu, lv)
    rb = refmath.product_of_experts(mu, lv)
    assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])
    # a13 chunk bounds, bit exact, through the reference method on a bare object
    mop = models.mopoe.__new__(models.mopoe)
    n = 0
    for S in (1, 2, 3, 5, 7, 15, 31, 63):
        for B in list(range(1, 70)) + [100, 127, 128, 255, 256, 1000, 1024, 4096, 65536, 99991]:
            mus = torch.arange(B, dtype=torch.float32).reshape(1, B, 1).repeat(S, 1, 1) + \
                1e6 * torch.arange(S, dtype=torch.float32).reshape(S, 1, 1)
            w = (1 / float(S)) * torch.ones(S)
            sel, _ = mop.moe_fusion(mus, mus.clone(), w)
            ref_map = (sel.reshape(-1) // 1e6).to(torch.int32)
            assert torch.equal(ref_map, refmath.mopoe_row_to_subset(S, B)), (S, B)
            n += 1
    # a19-a22 ReconLoss through recon_loss_fn
    obj = objectives.MultimodalObjective("elbo", 1.0)
    import torch.distributions as dist
    for ltype, lik, shape in [("bce", "normal", (6, 3, 4, 4)), ("mse", "normal", (6, 7)), ("l1", "normal", (6, 7)),
                              ("category_ce", "normal", (6, 5, 27)), ("category_ce", "normal", (6, 9)),
                              ("lprob", "normal", (6, 11)), ("lprob", "laplace", (6, 11)),
                              ("optimal_sigma", "normal", (6, 3, 4))]:
        for K in (1, 3):
            x = torch.randn(K * shape[0], *shape[1:], generator=g)
            if ltype == "bce":
                x = torch.sigmoid(x)
                x[0].fill_(0.0)  # exercises the log clamp at -100
                x[1].fill_(1.0)
            t = torch.rand(shape, generator=g)
            x1 = x.clone().requires_grad_(True)
            x2 = x.clone().requires_grad_(True)
            obj.set_ltype(ltype)
            D = dist.Laplace if lik == "laplace" else dist.Normal
            ra = obj.recon_loss_fn(D(x1, torch.tensor(0.75)), {"data": t, "masks": None}, K=K)
            rb = refmath.recon_logp(ltype, x2, t, K, lik)
            assert ra.shape == rb.shape and ra.dtype == rb.dtype, (ltype, ra.shape, rb.shape)
            assert torch.allclose(ra, rb, rtol=1e-6, atol=0), ltype
            wgt = torch.randn(ra.shape, generator=g).to(ra.dtype)
            (ra * wgt).sum().backward()
            (rb * wgt).sum().backward()
            assert torch.allclose(x1.grad, x2.grad, rtol=1e-6, atol=1e-9), ltype
            n += 1
    # a23 kl closed forms
    for (d1, f) in [(dist.Normal, refmath.kl_normal_normal), (dist.Laplace, refmath.kl_laplace_normal)]:
        l, s = torch.randn(5, 4, generator=g), torch.rand(5, 4, generator=g) + 0.1
        l0, s0 = torch.randn(1, 4, generator=g), torch.rand(1, 4, generator=g) + 0.5
        assert torch.allclose(utils.kl_divergence(d1(l, s), dist.Normal(l0, s0)), f(l, s, l0, s0), rtol=1e-6)
        n += 1
    l, s = torch.randn(5, 4, generator=g), torch.rand(5, 4, generator=g) + 0.1
    assert torch.allclose(utils.kl_divergence(dist.Laplace(l, s), dist.Laplace(l0, s0)),
                          refmath.kl_laplace_laplace(l, s, l0, s0), rtol=1e-6)
    # a25
    v = torch.randn(6, 5, generator=g)
    assert torch.equal(utils.log_mean_exp(v), refmath.log_mean_exp(v))
    # a11 subsets order
    for M in (2, 3, 4):
        names = ["mod_%d" % (i + 1) for i in range(M)]
        mop.vaes = {k: k for k in names}
        keys = list(models.mopoe.set_subsets(mop).keys())
        assert keys == ["_".join(names[i] for i in s) for s in refmath.mopoe_subsets(range(M))], keys
    return n


def unimodal_cases():
    g = torch.Generator().manual_seed(77)
    out = []
    for name, ltype, distn, shape, K, beta in [("uni_elbo_bce", "bce", "normal", (3, 8, 8), 1, 1.0),
                                               ("uni_elbo_laplace", "lprob", "laplace", (1, 7, 7), 1, 2.0),
                                               ("uni_elbo_ce", "category_ce", "normal", (5, 27), 1, 0.5)]:
        B, D = 6, 5
        import mmvae_b200.synthetic as syn
        mu, s = syn.make_posterior(g, B, D)
        P = 1
        for x in shape:
            P *= x
        W = torch.randn(P, D, generator=g) * 0.3
        b = torch.randn(P, generator=g) * 0.1
        target = syn.make_target(g, "onehot" if ltype == "category_ce" else "uniform", B, shape)
        noise = syn.make_noise(g, distn, (K, B, D))
        out.append(dict(name=name, ltype=ltype, dist=distn, shape=shape, K=K, beta=beta, mu=mu, s=s, W=W, b=b,
                        target=target, noise=noise))
    return out


def run_reference_unimodal(c):
    """Reference UnimodalObjective.elbo through calculate_loss with the distributions VAE.forward builds."""
    import torch.distributions as dist
    models, objectives, utils = ref_inplace.load()
    Dcls = dist.Laplace if c["dist"] == "laplace" else dist.Normal
    mu, s, W, b = (c[k].clone().requires_grad_(True) for k in ("mu", "s", "W", "b"))
    obj = objectives.UnimodalObjective("elbo", c["beta"])
    obj.set_ltype(c["ltype"])
    with ref_inplace.NoiseInjector([c["noise"].clone()]):
        qz_x = Dcls(mu, s)
        zs = qz_x.rsample(torch.Size([c["K"]]))
    lin = zs.reshape(1, -1, zs.shape[-1]) @ W.t() + b
    if c["ltype"] == "bce":
        lin = torch.sigmoid(lin).clamp(1e-6, 1 - 1e-6)
    loc = lin.reshape(-1, *c["shape"])
    px_z = Dcls(loc, torch.tensor(0.75))
    pz_params = (torch.zeros(1, mu.shape[1]), torch.ones(1, mu.shape[1]))
    out = obj.calculate_loss(px_z, {"data": c["target"], "masks": None}, qz_x, dist.Normal, pz_params, zs, K=c["K"])
    out["loss"].backward()
    return {"loss": out["loss"].detach().double(), "kld": out["kld"].detach().double(),
            "grad.mu": mu.grad.double(), "grad.s": s.grad.double(), "grad.W": W.grad.double(), "grad.b": b.grad.double()}


def run_oracle_unimodal(c, dtype=torch.float32):
    mu, s, W, b = (c[k].to(dtype).clone().requires_grad_(True) for k in ("mu", "s", "W", "b"))

    def dec(z):
        lin = z @ W.t() + b
        if c["ltype"] == "bce":
            lin = torch.sigmoid(lin).clamp(1e-6, 1 - 1e-6)
        return lin.reshape(-1, *c["shape"])
    out = refmath.unimodal_elbo(mu, s, c["dist"], c["noise"].to(dtype), dec, c["target"].to(dtype), c["ltype"], c["beta"], c["K"])
    out["loss"].backward()
    return {"loss": out["loss"].detach().double(), "kld": out["kld"].detach().double(),
            "grad.mu": mu.grad.double(), "grad.s": s.grad.double(), "grad.W": W.grad.double(), "grad.b": b.grad.double()}


def kl_df_cases():
    """Inputs of the analysis hook utils.make_kl_df (utils.py:130-162): M posteriors (n, D) + the Normal prior row."""
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(91)
    out = []
    for name, fam, M, n, D in (("kl_df_normal_m2", "normal", 2, 7, 6), ("kl_df_laplace_m3", "laplace", 3, 5, 4),
                               ("kl_df_normal_m1", "normal", 1, 4, 8)):
        post = [syn.make_posterior(g, n, D) for _ in range(M)]
        out.append(dict(name=name, family=fam, locs=[p[0] for p in post], scales=[p[1] for p in post],
                        loc0=torch.randn(1, D, generator=g) * 0.2, scale0=torch.softmax(torch.randn(1, D, generator=g), 1) * D))
    return out


def run_reference_kl_df(c):
    """The reference's make_kl_df on torch.distributions objects; returns the DataFrame as plain python / tensors."""
    import torch.distributions as dist
    models, objectives, utils = ref_inplace.load()
    cls = dist.Laplace if c["family"] == "laplace" else dist.Normal
    qs = [cls(l.clone(), s.clone()) for l, s in zip(c["locs"], c["scales"])]
    df = utils.make_kl_df(qs, dist.Normal(c["loc0"].clone(), c["scale0"].clone()))
    return {"columns": list(df.columns), "keys": [str(k) for k in df[df.columns[0]].tolist()],
            "dims": torch.tensor(df[df.columns[1]].to_numpy().astype("int64")),
            "values": torch.tensor(df[df.columns[2]].to_numpy().astype("float64"))}


def oracle_kl_df_values(c, dtype=torch.float32):
    """refmath.kl_table laid out like the DataFrame's value column (table-major, then dimension, then sample)."""
    t = refmath.kl_table(c["family"], [l.to(dtype) for l in c["locs"]], [s.to(dtype) for s in c["scales"]],
                         c["loc0"].to(dtype), c["scale0"].to(dtype))
    return t.permute(0, 2, 1).reshape(-1).double()


def main():
    assert ref_inplace.available(), "needs /root/reference"
    for c in kl_df_cases():
        ref = run_reference_kl_df(c)
        mine = oracle_kl_df_values(c)
        worst = float((mine - ref["values"]).abs().max() / ref["values"].abs().max())
        assert worst < 1e-6, (c["name"], worst)
        print("%-22s make_kl_df worst rel diff vs reference %.2e (%d rows)" % (c["name"], worst, mine.numel()))
    for c in unimodal_cases():
        worst = compare(run_reference_unimodal(c), run_oracle_unimodal(c), c["name"])
        print("%-22s unimodal worst rel diff vs reference %.2e" % (c["name"], worst))
    n = check_functions()
    print("function-level checks ok (%d comparisons)" % n)
    for case in cases.case_list():
        ref, subsets = run_reference(case)
        if subsets is not None:
            case["poe_subsets"] = subsets
        orc = cases.run_oracle(case)
        worst = compare(ref, orc, case["name"])
        print("%-22s loss=% .6f  worst rel diff vs reference %.2e  (%d tensors)" % (
            case["name"], float(ref["loss"]), worst, len(ref)))
    print("ORACLE PINNED: restatement == reference on all cases")


if __name__ == "__main__":
    main()
