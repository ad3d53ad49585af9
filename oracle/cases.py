"""TEST INFRASTRUCTURE: seeded parity cases shared by the validation script, the golden generator and tests/.

A *case* is a plain dict of python scalars + CPU tensors fully describing one objective evaluation:
model kind, objective, K, beta, latent sizes, per-modality (data_dim, ltype, dist, lam), encoder outputs
(mu, s), linear stand-in decoder weights, targets, prior logits and the noise tensors in the reference's
rsample order.  ``run_oracle`` evaluates it with oracle/refmath.py; ``oracle/validate_against_reference.py``
evaluates the same case with the unmodified reference classes.
"""
import math
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import mmvae_b200.synthetic as syn  # noqa: E402  (torch-only helpers: stub VAEs + generators)

from . import refmath  # noqa: E402


def noise_plan(model, M, K, B, D, privs, dists):
    """Shapes + kinds of the noise tensors in the reference's rsample order."""
    if model == "poe":
        return [("normal", (1, B, D)) for _ in range(2 ** M - 1)]
    if model == "moe":
        return [(dists[m], (K, B, D)) for m in range(M)]
    if model == "mopoe":
        return [("normal", (K, B, D)) for _ in range(M)]
    if model == "dmvae":
        plan = [("normal", (K, B, D))]
        for m in range(M):
            plan.append(("normal", (K, B, D)))
            plan.append(("normal", (K, B, privs[m])))
            plan.extend(("normal", (1, B, D)) for _ in range(M - 1))
        return plan
    raise ValueError(model)


def make_case(name, seed, model, obj, B, D, mods, K=1, beta=1.0, private=None, pz_logits_std=0.3, keep_k=False):
    g = syn.gen(seed)
    M = len(mods)
    case = dict(name=name, seed=seed, model=model, obj=obj, B=B, D=D, K=K, beta=beta, private=private,
                keep_k=keep_k, mods=[])
    privs = []
    for spec in mods:
        pv = (private or 0)
        privs.append(pv)
        dz = D + pv
        mu, s = syn.make_posterior(g, B, dz)
        # dec_dim: what the decoder emits; with padding masks the target (and the mask) is shorter along dim 1 and
        # recon_loss_fn crops the decoder output (objectives.py:43-45)
        dec_dim = tuple(spec.get("dec_dim", spec["data_dim"]))
        P = int(math.prod(dec_dim))
        squash = spec.get("squash", spec["ltype"] in ("bce",))
        W = torch.randn(P, dz, generator=g) * (0.5 / math.sqrt(dz))
        b = torch.randn(P, generator=g) * 0.1
        target = syn.make_target(g, spec.get("target", "uniform"), B, spec["data_dim"])
        mask_len = spec["data_dim"][0] if "dec_dim" in spec else None
        case["mods"].append(dict(data_dim=dec_dim, ltype=spec["ltype"], dist=spec.get("dist", "normal"),
                                 lam=float(spec.get("lam", 1.0)), mu=mu, s=s, W=W, b=b, target=target,
                                 squash=squash, mask_len=mask_len))
    case["pz_logits"] = torch.randn(1, D, generator=g) * pz_logits_std
    plan = noise_plan(model, M, K, B, D, privs, [m["dist"] for m in case["mods"]])
    case["noise"] = [syn.make_noise(g, kind, shape) for kind, shape in plan]
    return case


def build_vaes(case, device="cpu", dtype=torch.float32):
    """StubVAEs with leaf encoders + linear decoders, keyed mod_1.. like the reference (trainer.py:99-107)."""
    vaes = {}
    for i, m in enumerate(case["mods"]):
        enc = syn.LeafEncoder(m["data_dim"], m["mu"].to(device=device, dtype=dtype),
                              m["s"].to(device=device, dtype=dtype))
        dec = syn.LinearDecoder(m["W"].shape[1], m["data_dim"], m["squash"])
        with torch.no_grad():
            dec.lin.weight.copy_(m["W"])
            dec.lin.bias.copy_(m["b"])
        dec = dec.to(device=device, dtype=dtype)
        if case.get("keep_k"):
            dec = KeepK(dec)
        vae = syn.StubVAE(enc, dec, case["D"], m["ltype"], private_latents=case.get("private"),
                          llik_scaling=m["lam"], prior_dist=m["dist"], id_name="mod_%d" % (i + 1))
        vaes["mod_%d" % (i + 1)] = vae.to(device)
    return vaes


class KeepK(torch.nn.Module):
    """Decoder wrapper that keeps the (K, B, ...) axes like reference Dec_MNIST / Dec_SVHN
    (decoders.py:145-147, :268-270) -- the only layout reference DReG accepts."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.data_dim = inner.data_dim

    def forward(self, z):
        K, B = z["latents"].shape[:2]
        mean, sc = self.inner(z)
        return mean.reshape(K, B, *self.data_dim), sc


def build_batch(case, device="cpu"):
    """Batch dict format of reference dataloader.py:85-120."""
    return {"mod_%d" % (i + 1): {"data": m["target"].to(device),
                                 "masks": None if m.get("mask_len") is None else
                                 torch.ones(m["target"].shape[0], m["mask_len"], dtype=torch.bool, device=device),
                                 "categorical": False}
            for i, m in enumerate(case["mods"])}


def named_leaves(vaes, pz_logits_param):
    out = {}
    for k, v in vaes.items():
        out[k + ".mu"] = v.enc.mu
        out[k + ".s"] = v.enc.s
        lin = v.dec.inner.lin if isinstance(v.dec, KeepK) else v.dec.lin
        out[k + ".W"] = lin.weight
        out[k + ".b"] = lin.bias
    out["pz_logits"] = pz_logits_param
    return out


def collect(outputs, leaves):
    res = {}
    for k in ("loss", "kld"):
        if k in outputs and outputs[k] is not None and torch.is_tensor(outputs[k]):
            res[k] = outputs[k].detach().double().cpu().clone()
    rl = outputs.get("reconstruction_loss")
    if isinstance(rl, (list, tuple)) and len(rl) and torch.is_tensor(rl[0]) and all(x.shape == rl[0].shape for x in rl):
        res["reconstruction_loss"] = torch.stack([x.detach().double().cpu() for x in rl])
    for n, p in leaves.items():
        res["grad." + n] = None if p.grad is None else p.grad.detach().double().cpu().clone()
    return res


def run_oracle(case, device="cpu", dtype=torch.float32, extra=False):
    """Evaluate ``case`` with the restatement (oracle/refmath.py); returns loss/kld/grads (fp64 copies)."""
    vaes = build_vaes(case, device, dtype)
    pz = torch.nn.Parameter(case["pz_logits"].to(device=device, dtype=dtype).clone())
    mods = []
    for i, m in enumerate(case["mods"]):
        v = vaes["mod_%d" % (i + 1)]
        inner = v.dec.inner if isinstance(v.dec, KeepK) else v.dec
        mods.append(dict(mu=v.enc.mu, s=v.enc.s, dist=m["dist"], ltype=m["ltype"], lam=m["lam"],
                         target=m["target"].to(device), mask_len=m.get("mask_len"),
                         dec=(lambda z, d=inner: d({"latents": z, "masks": None})[0])))
    noise = [n.to(device=device, dtype=dtype) for n in case["noise"]]
    fn = {"poe": refmath.poe_objective, "mopoe": refmath.mopoe_objective, "dmvae": refmath.dmvae_objective}
    if case["model"] == "moe":
        out = refmath.moe_objective(mods, pz, noise, obj=case["obj"], beta=case["beta"], K=case["K"])
    elif case["model"] == "poe":
        out = fn["poe"](mods, pz, noise, beta=case["beta"], subsets=case.get("poe_subsets"))
    else:
        out = fn[case["model"]](mods, pz, noise, beta=case["beta"], K=case["K"])
    out["loss"].backward()
    res = collect(out, named_leaves(vaes, pz))
    if extra:
        res["_raw"] = out
    return res


# ----------------------------------------------------------------------------------------------------------
# The frozen case list (small enough for tests/golden/*.pt; covers every model x objective x likelihood family)
# ----------------------------------------------------------------------------------------------------------
IMG = dict(data_dim=(3, 8, 8), ltype="bce", target="uniform")
TXT = dict(data_dim=(5, 27), ltype="category_ce", target="onehot")
ACT = dict(data_dim=(9,), ltype="category_ce", target="onehot")
ATT = dict(data_dim=(4, 6), ltype="category_ce", target="onehot")
MSE = dict(data_dim=(6, 5), ltype="mse", target="uniform")
L1 = dict(data_dim=(11,), ltype="l1", target="uniform")
OSG = dict(data_dim=(3, 4, 4), ltype="optimal_sigma", target="uniform")
LAP_A = dict(data_dim=(1, 7, 7), ltype="lprob", target="uniform", dist="laplace", lam=1.0)
LAP_B = dict(data_dim=(3, 6, 6), ltype="lprob", target="uniform", dist="laplace", lam=49.0 / 108.0)
NRM_A = dict(data_dim=(1, 7, 7), ltype="lprob", target="uniform", dist="normal", lam=1.0)
TXT_MASKED = dict(data_dim=(5, 27), dec_dim=(8, 27), ltype="category_ce", target="onehot")  # decoder pads to T=8
# lprob + padding masks: recon_loss_fn overwrites the likelihood scale with the cropped loc (objectives.py:43-45);
# sigmoid decoders keep loc (= scale) positive and away from 0 (an unsquashed decoder gives log(negative) = NaN -> 0
# entries and, for 0 < loc << 1, gradients ~ 1/loc^2 whose decoder-weight contraction is ill conditioned: the NaN path is
# covered at kernel level, tests/test_ops_gpu.py::test_lprob_selfscale)
# (only with mask length == decoder length: a real crop makes the reference raise in torch's _validate_sample, the
# distribution keeps the batch_shape it was built with)
LPM_NRM = dict(data_dim=(5, 6), dec_dim=(5, 6), ltype="lprob", target="uniform", dist="normal", squash=True)
LPM_LAP = dict(data_dim=(4, 7), dec_dim=(4, 7), ltype="lprob", target="uniform", dist="laplace", squash=True, lam=0.5)


def case_list():
    c = []
    c.append(make_case("poe_elbo_m2", 101, "poe", "elbo", B=6, D=4, mods=[IMG, TXT]))
    c.append(make_case("poe_elbo_m3", 102, "poe", "elbo", B=5, D=10, mods=[IMG, ACT, ATT], beta=2.5))
    c.append(make_case("poe_elbo_misc", 103, "poe", "elbo", B=7, D=3, mods=[MSE, OSG]))
    c.append(make_case("moe_elbo_m2", 201, "moe", "elbo", B=6, D=4, mods=[IMG, TXT]))
    c.append(make_case("moe_elbo_laplace", 202, "moe", "elbo", B=6, D=8, mods=[LAP_A, LAP_B], beta=0.5))
    c.append(make_case("moe_iwae_m2", 203, "moe", "iwae", B=6, D=4, K=3, mods=[IMG, TXT]))
    c.append(make_case("moe_iwae_laplace", 204, "moe", "iwae", B=4, D=8, K=5, mods=[LAP_A, LAP_B], beta=1.5))
    c.append(make_case("moe_dreg_laplace", 205, "moe", "dreg", B=4, D=8, K=5, mods=[LAP_A, LAP_B], keep_k=True))
    c.append(make_case("moe_dreg_normal", 206, "moe", "dreg", B=3, D=6, K=4, mods=[NRM_A, MSE], keep_k=True))
    c.append(make_case("mopoe_elbo_m2", 301, "mopoe", "elbo", B=8, D=4, mods=[IMG, TXT]))
    c.append(make_case("mopoe_elbo_m3", 302, "mopoe", "elbo", B=16, D=10, mods=[IMG, ACT, ATT], beta=0.7))
    c.append(make_case("mopoe_elbo_osigma", 303, "mopoe", "elbo", B=9, D=5, mods=[OSG, MSE, L1]))
    c.append(make_case("dmvae_elbo_m2", 401, "dmvae", "elbo", B=6, D=4, private=3, mods=[IMG, TXT]))
    c.append(make_case("dmvae_elbo_m3", 402, "dmvae", "elbo", B=5, D=4, private=2, mods=[IMG, ACT, ATT], beta=1.3))
    c.append(make_case("poe_elbo_masks", 501, "poe", "elbo", B=6, D=4, mods=[IMG, TXT_MASKED]))
    c.append(make_case("moe_iwae_masks", 502, "moe", "iwae", B=4, D=4, K=3, mods=[IMG, TXT_MASKED]))
    c.append(make_case("mopoe_elbo_masks", 503, "mopoe", "elbo", B=7, D=6, mods=[TXT_MASKED, IMG, ACT]))
    c.append(make_case("poe_elbo_lprob_masks", 504, "poe", "elbo", B=6, D=4, mods=[IMG, LPM_NRM]))
    c.append(make_case("mopoe_elbo_lprob_masks", 505, "mopoe", "elbo", B=5, D=4, mods=[LPM_LAP, LPM_NRM]))
    return c
