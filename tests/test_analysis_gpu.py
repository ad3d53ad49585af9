"""GPU: the analysis / evaluation consumers of the hot-path math -- utils.make_kl_df (per-dimension KL
and symmetric J-divergence tables, reference utils.py:130-162), eval_forward and the numerical part of analyse_data
(reference trainer.py:242-279) -- against the frozen outputs of the reference's own make_kl_df and the oracle."""
import pytest
import torch
import torch.distributions as dist

pytestmark = pytest.mark.gpu

from oracle import cases, refmath  # noqa: E402

TOL = 1e-5


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


def test_make_kl_df_matches_reference_dataframe(golden):
    import mmvae_b200.utils as U
    for entry in golden["kl_df"]:
        c, ref = entry["case"], entry["reference"]
        cls = dist.Laplace if c["family"] == "laplace" else dist.Normal
        qs = [cls(l.cuda(), s.cuda()) for l, s in zip(c["locs"], c["scales"])]
        df = U.make_kl_df(qs, dist.Normal(c["loc0"].cuda(), c["scale0"].cuda()))
        assert list(df.columns) == ref["columns"], c["name"]
        assert [str(k) for k in df[df.columns[0]].tolist()] == ref["keys"], c["name"]  # row order and key strings
        assert torch.equal(torch.tensor(df[df.columns[1]].to_numpy().astype("int64")), ref["dims"])
        vals = torch.tensor(df[df.columns[2]].to_numpy().astype("float64"))
        assert _rel(vals, ref["values"]) < TOL, c["name"]


@pytest.mark.parametrize("family,M,n,D", [("normal", 2, 1000, 16), ("laplace", 4, 33, 64), ("normal", 3, 7, 10)])
def test_kl_table_kernel_matches_oracle(family, M, n, D):
    import mmvae_b200.ops as ops
    import mmvae_b200.synthetic as syn
    g = torch.Generator().manual_seed(3)
    post = [syn.make_posterior(g, n, D) for _ in range(M)]
    loc0 = torch.randn(1, D, generator=g) * 0.3
    s0 = torch.softmax(torch.randn(1, D, generator=g), 1) * D
    ref = refmath.kl_table(family, [p[0].double() for p in post], [p[1].double() for p in post], loc0.double(), s0.double())
    got = ops.kl_table(torch.stack([p[0] for p in post]).cuda(), torch.stack([p[1] for p in post]).cuda(), loc0.cuda(),
                       s0.cuda(), laplace=family == "laplace")
    assert got.shape == ref.shape == (M + M * (M - 1) // 2, n, D)
    assert _rel(got, ref) < TOL


@pytest.mark.parametrize("name", ["poe_elbo_m2", "moe_elbo_m2", "mopoe_elbo_m3", "dmvae_elbo_m2"])
def test_eval_forward_and_analyse_data(name):
    import mmvae_b200
    import mmvae_b200.utils as U
    case = next(c for c in cases.case_list() if c["name"] == name)
    vaes = cases.build_vaes(case, "cuda")
    model = mmvae_b200.MODEL_REGISTRY[case["model"]](vaes, case["D"], {"obj": "elbo", "beta": 1.0, "K": 1}, None).cuda()
    batch = cases.build_batch(case, "cpu")  # host batch: eval_forward moves it (reference data_to_device)
    out = U.eval_forward(model, batch)
    M = len(case["mods"])
    assert len(out["encoder_dist"]) == M and len(out["decoder_dist"]) == M and len(out["latent_samples"]) == M
    assert all(d.loc.is_cuda and not d.loc.requires_grad for d in out["decoder_dist"])
    res = U.analyse_data(model, batch, num_samples=17)
    zss = res["latent_samples"]
    assert zss[0].shape == (17, case["D"]) and len(zss) == M + 1
    # the table equals the oracle's closed forms on the encoder distributions the model returned
    qs = res["output"]["encoder_dist"]
    mu0, s0 = model.pz_params
    ref = refmath.kl_table("normal", [q.loc.double().cpu() for q in qs], [q.scale.double().cpu() for q in qs],
                           mu0.double().cpu(), s0.double().cpu())
    df = res["kl_df"]
    vals = torch.tensor(df[df.columns[2]].to_numpy().astype("float64"))
    assert _rel(vals, ref.permute(0, 2, 1).reshape(-1)) < TOL
    assert df[df.columns[0]].nunique() == M + M * (M - 1) // 2
