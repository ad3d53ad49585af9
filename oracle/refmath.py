"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A compact, device-agnostic, pure-torch restatement of the latent + objective hot path of
gabinsane/multimodal-vae-comparison.  Every function cites the reference
file:line it follows (paths relative to /root/reference/multimodal_compare).  It exists so that

  * tests/ can check the CUDA kernels against it on the GPU box (where /root/reference is absent),
  * bench.py's ``cpu_baseline`` / ``--impl reference`` leg can time the reference algorithm on host cores.

Only ``tests/``, ``__graft_entry__.smoke()`` and those bench legs may import this module.  The product
package never does (tests/test_no_oracle_in_product.py enforces it).

PARITY PIN: the reference ships no golden vector / known-answer test for this path, so
the pin is the reference ITSELF: ``oracle/validate_against_reference.py`` (run in the build container, where
/root/reference is mounted) executes the unmodified reference classes (MOE/POE/MoPOE/DMVAE, ReconLoss,
MultimodalObjective) in place on seeded inputs and asserts that this restatement reproduces loss, KL,
reconstruction terms and all gradients; ``oracle/gen_golden.py`` freezes those reference outputs under
tests/golden/*.pt and tests/test_oracle_golden.py re-checks the restatement against them on every run.

Semantics notes N1-N5 of SURVEY.md section 8a are applied exactly; reference quirks are reproduced, not fixed:
the PoE *variance* is used as the Normal scale, encoder "logvar" is softmax+1e-6 and gets exponentiated,
category_ce soft-maxes over dim 1, MoE-ELBO counts the total KL 2M times, DReG weights are a softmax over K of
batch-summed log-weights, optimal_sigma detaches its squared term, lprob accumulates in fp64.
"""
import itertools
import math

import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


# ----------------------------------------------------------------------------------------------------------
# torch.distributions formulas the reference relies on (third party, torch 2.11; formulas unchanged since 1.12)
# ----------------------------------------------------------------------------------------------------------
def normal_log_prob(z, loc, scale):
    """torch/distributions/normal.py:88-101."""
    var = scale ** 2
    return -((z - loc) ** 2) / (2 * var) - scale.log() - LOG_SQRT_2PI


def laplace_log_prob(z, loc, scale):
    """torch/distributions/laplace.py:86-90."""
    return -torch.log(2 * scale) - torch.abs(z - loc) / scale


def normal_rsample(loc, scale, eps):
    """torch/distributions/normal.py:82-86: loc + eps * scale, eps ~ N(0,1) of shape (K, *loc.shape)."""
    return loc + eps * scale


def laplace_rsample(loc, scale, u):
    """torch/distributions/laplace.py:73-84: u ~ U(finfo.eps - 1, 1); loc - scale*sign(u)*log1p(-|u|)."""
    return loc - scale * u.sign() * torch.log1p(-u.abs())


def kl_normal_normal(loc_p, scale_p, loc_q, scale_q):
    """torch/distributions/kl.py _kl_normal_normal."""
    var_ratio = (scale_p / scale_q).pow(2)
    t1 = ((loc_p - loc_q) / scale_q).pow(2)
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def kl_laplace_laplace(loc_p, scale_p, loc_q, scale_q):
    """torch/distributions/kl.py _kl_laplace_laplace."""
    scale_ratio = scale_p / scale_q
    loc_abs_diff = (loc_p - loc_q).abs()
    t1 = -scale_ratio.log()
    t2 = loc_abs_diff / scale_q
    t3 = scale_ratio * torch.exp(-loc_abs_diff / scale_p)
    return t1 + t2 + t3 - 1


def kl_laplace_normal(loc_p, scale_p, loc_q, scale_q):
    """torch/distributions/kl.py _kl_laplace_normal."""
    var_normal = scale_q.pow(2)
    scale_sqr_var_ratio = scale_p.pow(2) / var_normal
    t1 = 0.5 * torch.log(2 * scale_sqr_var_ratio / math.pi)
    t2 = 0.5 * loc_p.pow(2) / var_normal
    t3 = loc_p * loc_q / var_normal
    t4 = 0.5 * loc_q.pow(2) / var_normal
    return -t1 + scale_sqr_var_ratio + t2 - t3 + t4 - 1


def rsample(dist_name, loc, scale, noise):
    return laplace_rsample(loc, scale, noise) if dist_name == "laplace" else normal_rsample(loc, scale, noise)


def log_prob(dist_name, z, loc, scale):
    return laplace_log_prob(z, loc, scale) if dist_name == "laplace" else normal_log_prob(z, loc, scale)


def kl_to_normal(dist_name, loc, scale, loc0, scale0):
    """utils.py:399-405 kl_divergence -> registered closed forms; the model prior is always Normal
    (mmvae_base.py:31)."""
    if dist_name == "laplace":
        return kl_laplace_normal(loc, scale, loc0, scale0)
    return kl_normal_normal(loc, scale, loc0, scale0)


# ----------------------------------------------------------------------------------------------------------
# helpers: utils.py
# ----------------------------------------------------------------------------------------------------------
def log_mean_exp(value, dim=0, keepdim=False):
    """utils.py:395-396."""
    return torch.logsumexp(value, dim, keepdim=keepdim) - math.log(value.size(dim))


def softclip(t, mn):
    """utils.py:66-69 (the reference casts to float; kept in the input dtype so fp64 evaluation stays fp64)."""
    return mn + F.softplus(t - mn)


def prior_params(pz_logits_mu, pz_logits):
    """mmvae_models.py:28-30 (and :155-157, :275-277, :432-434): (mu0, softmax(_pz_params[1], 1) * D)."""
    return pz_logits_mu, F.softmax(pz_logits, dim=1) * pz_logits.size(-1)


def poe_subsets(mod_names):
    """utils.py:86-112 subsample_input_modalities: all non-empty subsets, grouped by size.  The reference
    iterates ``list(set(combinations))`` so the order INSIDE one size is PYTHONHASHSEED dependent; the loss is
    a plain sum over subsets so only the eps<->subset pairing depends on it.  This restatement (and the CUDA
    drop-in) uses itertools.combinations order; golden files record the order the reference run used."""
    out = []
    for n in range(1, len(mod_names) + 1):
        out.extend(itertools.combinations(list(mod_names), n))
    return out


def mopoe_subsets(mod_names):
    """mmvae_models.py:279-294 set_subsets: powerset without the empty set, by size then combinations order,
    keys '_'.join(sorted(names))."""
    xs = list(mod_names)
    out = []
    for n in range(1, len(xs) + 1):
        for c in itertools.combinations(xs, n):
            out.append(tuple(sorted(c)))
    return out


def mopoe_chunk_bounds(num_components, num_samples):
    """mmvae_models.py:339 + :377-410: weights = (1/float(S))*ones(S); reweight w/w.sum(); chunk k covers rows
    [start_k, start_k + int(floor(B * w_k))), the last chunk is extended to B.  Evaluated with the same fp32
    torch ops as the reference so that the integer bounds are bit-exact."""
    w = (1 / float(num_components)) * torch.ones(num_components)
    w = w / w.sum()
    idx_start, idx_end = [], []
    for k in range(num_components):
        i_start = 0 if k == 0 else int(idx_end[k - 1])
        if k == num_components - 1:
            i_end = num_samples
        else:
            i_end = i_start + int(torch.floor(num_samples * w[k]))
        idx_start.append(i_start)
        idx_end.append(i_end)
    idx_end[-1] = num_samples
    return idx_start, idx_end


def mopoe_row_to_subset(num_components, num_samples):
    """Row -> selected mixture component, implied by mixture_component_selection's torch.cat of slices."""
    s, e = mopoe_chunk_bounds(num_components, num_samples)
    idx = torch.empty(num_samples, dtype=torch.int32)
    for k in range(num_components):
        idx[s[k]:e[k]] = k
    return idx


# ----------------------------------------------------------------------------------------------------------
# a1: product of experts
# ----------------------------------------------------------------------------------------------------------
def product_of_experts(mu, logvar):
    """mmvae_base.py:203-222.  mu, logvar: (E, B, D).  Returns (pd_mu, pd_var); the caller uses pd_var as the
    Normal *scale* (mmvae_models.py:200, :365, :480)."""
    eps = 1e-8
    var = torch.exp(logvar) + eps
    T = 1.0 / var
    pd_mu = torch.sum(mu * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    return pd_mu, pd_var


# ----------------------------------------------------------------------------------------------------------
# a18-a22: reconstruction terms.  recon_logp returns  -loss  of shape (rows, -1)  (objectives.py:30-52)
# ----------------------------------------------------------------------------------------------------------
def reshape_target(loc, target, K):
    """objectives.py:103-125 reshape_for_loss: target.repeat(K,1,...).reshape(loc.shape) (k-major rows)."""
    return target.repeat(K, *([1] * (target.dim() - 1))).reshape(*loc.shape)


def recon_logp(ltype, loc, target, K=1, likelihood="normal", scale=0.75, mask_len=None):
    """BaseObjective.recon_loss_fn objectives.py:30-52 + ReconLoss.* objectives.py:389-509.
    loc: decoder mean, rows = K*B (k-major).  Returns -loss reshaped (rows, -1).
    mask_len: objectives.py:43-45 crops loc[:, :masks.shape[1]]."""
    if mask_len is not None:
        loc = loc[:, :mask_len]
    # reference: target.float(); the cast follows loc so that the restatement can also be evaluated in fp64
    target = reshape_target(loc, target.to(loc.dtype), K)
    bs = target.shape[0]
    if ltype == "bce_logits":  # decoder tail decoders.py:96-97 (sigmoid + clamp(eta, 1-eta)) followed by bce
        # clamp bounds as the fp32 reference sees them (fp32(1e-6), fp32(1-1e-6)), also when evaluated in fp64
        lo = float(torch.tensor(1e-6, dtype=torch.float32))
        hi = float(torch.tensor(1 - 1e-6, dtype=torch.float32))
        xs = torch.sigmoid(loc).clamp(lo, hi)
        loss = F.binary_cross_entropy(xs, target.detach(), reduction="none").reshape(bs, -1)
    elif ltype == "bce":  # objectives.py:391-406
        loss = F.binary_cross_entropy(loc, target.detach(), reduction="none").reshape(bs, -1)
    elif ltype == "lprob":  # objectives.py:408-424 (fp64 accumulate, NaN -> 0)
        # objectives.py:43-45: with padding masks the likelihood's scale is overwritten by its (cropped) loc
        sc = loc if mask_len is not None else torch.as_tensor(scale, dtype=loc.dtype, device=loc.device)
        out = log_prob(likelihood, target, loc, sc).view(bs, -1).double().reshape(bs, -1)
        out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
        loss = -out
    elif ltype == "l1":  # objectives.py:426-441
        loss = F.l1_loss(loc, target.detach(), reduction="none").reshape(bs, -1)
    elif ltype == "mse":  # objectives.py:443-458
        loss = F.mse_loss(loc, target.detach(), reduction="none").reshape(bs, -1)
    elif ltype == "category_ce":  # objectives.py:485-500 -- class axis is dim 1
        loss = F.cross_entropy(loc, target.detach(), reduction="none").reshape(bs, -1)
    elif ltype == "optimal_sigma":  # objectives.py:502-509
        t = target.detach()
        log_sigma = ((t - loc) ** 2).mean(list(range(loc.dim())), keepdim=True).sqrt().log()
        log_sigma = log_sigma.reshape(())
        log_sigma = softclip(log_sigma, -6)
        loss = (torch.pow((t - loc) / log_sigma.exp(), 2).clone().detach() + log_sigma
                + 0.5 * math.log(2 * math.pi)).reshape(bs, -1)
    else:
        raise ValueError(ltype)
    return -loss


def lpx_rows(ltype, loc, target, lam, K=1, likelihood="normal", mask_len=None):
    """The recurring pattern ``(recon_loss_fn(px_z, x) * llik_scaling).sum(-1)`` (mmvae_models.py:48-49, :177,
    :312-313, :448) -> one value per decoder row."""
    return (recon_logp(ltype, loc, target, K, likelihood, mask_len=mask_len) * lam).sum(-1)


# ----------------------------------------------------------------------------------------------------------
# a24: ELBO
# ----------------------------------------------------------------------------------------------------------
def elbo(lpx_z, kld, beta):
    """objectives.py:54-67: -(lpx_z.sum(-1) - beta*kld.sum()).sum()  (kld.sum() is a scalar broadcast against
    every leading row of lpx_z)."""
    return -(lpx_z.sum(-1) - beta * kld.sum()).sum()


# ----------------------------------------------------------------------------------------------------------
# Model-level restatements.  Common calling convention:
#   mods      : list of modality specs, dicts with keys
#                 mu, s        (B, Dtot) encoder outputs (None when the modality is absent)
#                 dist         "normal" | "laplace"   (posterior == likelihood family, vae.py:142-147)
#                 ltype, lam   likelihood name and llik_scaling
#                 target       (B, ...) data tensor; mask_len optional int
#                 dec          callable: latents (K,B,Dz) -> decoder mean with rows K*B (CNN/FNN convention)
#                 n_private    private latent width (DMVAE) or 0
#   pz_logits : (1, D) learnable _pz_params[1]      (mmvae_base.py:35-38); _pz_params[0] == 0
#   noise     : list of noise tensors consumed in the reference's rsample order
# ----------------------------------------------------------------------------------------------------------
def poe_mixing(mods, present, B, D):
    """POE.modality_mixing + prior_expert mmvae_models.py:210-250: prior expert (0, log 1 = 0) FIRST, then the
    present modalities in vaes order; product_of_experts over the stack."""
    ref = next(m for i, m in enumerate(mods) if i in present)
    mu = torch.zeros(1, B, D, dtype=ref["mu"].dtype, device=ref["mu"].device)
    lv = torch.zeros(1, B, D, dtype=ref["mu"].dtype, device=ref["mu"].device)
    for i, m in enumerate(mods):
        if i in present:
            mu = torch.cat((mu, m["mu"].unsqueeze(0)), 0)
            lv = torch.cat((lv, m["s"].unsqueeze(0)), 0)
    return product_of_experts(mu, lv)


def poe_objective(mods, pz_logits, noise, beta=1.0, subsets=None):
    """POE.objective mmvae_models.py:159-187 with POE.forward :189-208 (K is fixed to 1).
    Returns dict(loss, kld, reconstruction_loss[list], fused[list of (mu,var)], z[list])."""
    M = len(mods)
    B, D = mods[0]["mu"].shape
    names = list(range(M))
    subsets = poe_subsets(names) if subsets is None else subsets
    mu0 = torch.zeros_like(pz_logits)
    _, s0 = prior_params(mu0, pz_logits)
    noise = list(noise)
    losses, klds, fused, zs = [], [], [], []
    lpx_log = [[] for _ in range(M)]
    for a, subset in enumerate(subsets):
        mu, var = poe_mixing(mods, set(subset), B, D)
        z = normal_rsample(mu, var, noise.pop(0))  # (1,B,D)   mmvae_models.py:200-201
        fused.append((mu, var))
        zs.append(z)
        kld = kl_normal_normal(mu, var, mu0, s0)  # (B,D)     :173
        klds.append(kld.sum(-1))
        loc_lpx = []
        for i, m in enumerate(mods):
            # decoders get the subset's masks (None for absent modalities) but the loss always sees the
            # full target dict (mods[mod], :177)
            loc = m["dec"](z)
            lp = lpx_rows(m["ltype"], loc, m["target"], m["lam"], 1, m["dist"], m.get("mask_len"))
            loc_lpx.append(lp)
            if i == a:  # quirk: "mod == 'mod_{m+1}'" with m the SUBSET index (:179-180)
                lpx_log[a].append(lp)
        losses.append(elbo(torch.stack(loc_lpx).sum(0), kld.sum(-1), beta))  # :181-182
    ind = [-torch.stack(l).sum() / mods[i]["lam"] for i, l in enumerate(lpx_log)]
    return {"loss": torch.stack(losses).sum(), "reconstruction_loss": ind,
            "kld": torch.stack(klds).mean(0).sum(), "fused": fused, "z": zs}


def moe_forward(mods, noise, K):
    """MOE.forward mmvae_models.py:80-117: z_m = q_m.rsample([K]) per modality in vaes order."""
    noise = list(noise)
    return [rsample(m["dist"], m["mu"], m["s"], noise.pop(0)) for m in mods]


def moe_cross_source(M, r):
    """mmvae_models.py:112-116: cross_px_zs[target] is overwritten by every source != target, so the LAST
    modality different from the target survives (for M=2: the other one)."""
    return [s for s in range(M) if s != r][-1]


def moe_objective(mods, pz_logits, noise, obj="elbo", beta=1.0, K=1):
    """MOE.objective mmvae_models.py:32-78 + MultimodalObjective.{elbo,iwae,dreg} objectives.py:316-387.
    The fixed VAE-level prior N(0,1) is used for the ELBO KL (:45, vae.py:159-162); iwae/dreg use the learnable
    model prior (pz_params)."""
    M = len(mods)
    zs = moe_forward(mods, noise, K)
    B = mods[0]["mu"].shape[0]
    mu0 = torch.zeros_like(pz_logits)
    _, s0 = prior_params(mu0, pz_logits)
    out = {"z": zs}
    klds, lpx_zs = [], []
    for r, m in enumerate(mods):
        one = torch.ones_like(m["mu"][:1])
        kld = kl_to_normal(m["dist"], m["mu"], m["s"], torch.zeros_like(one), one)  # (B,D)
        klds.append(kld.sum(-1))
        # quirk: MOE.forward wraps the SELF reconstruction in dist.Normal regardless of the VAE's likelihood
        # family (mmvae_models.py:105-107) while cross reconstructions use vae.px_z (:116)
        lp_self = lpx_rows(m["ltype"], m["dec"](zs[r]), m["target"], m["lam"], K, "normal", m.get("mask_len"))
        s = moe_cross_source(M, r)
        lp_cross = lpx_rows(m["ltype"], m["dec"](zs[s]), m["target"], m["lam"], K, m["dist"], m.get("mask_len"))
        if obj == "elbo":
            zd = zs[s].detach()
            lwt = (log_prob(m["dist"], zd, m["mu"], m["s"])
                   - log_prob(mods[s]["dist"], zd, mods[s]["mu"], mods[s]["s"]).detach()).sum(-1).reshape(-1)
            lpx_zs.append(lp_self)  # exp(0) * lpx  (:50, :60)
            lpx_zs.append(lwt.exp() * lp_cross)  # :61
        else:
            lpx_zs.append([lp_self, lp_cross])  # :63-72
    if obj == "elbo":
        # :73 rows whose sum is exactly 0 are dropped (data dependent: exp(lwt) underflows for sharp posteriors),
        # which also lowers the number of times beta*kld.sum() is subtracted (objectives.py:67 broadcast)
        lpx = torch.stack([lp for lp in lpx_zs if lp.sum() != 0])
        loss = (1 / M) * elbo(lpx, torch.stack(klds), beta)  # objectives.py:330-340, mmvae_models.py:76-77
        out.update(loss=loss, kld=torch.stack(klds), reconstruction_loss=lpx)
        return out
    lq = lambda r: log_mean_exp(torch.stack(
        [log_prob(mj["dist"], zs[r], mj["mu"], mj["s"]).sum(-1) for mj in mods]))  # (K,B)
    lpz = lambda r: normal_log_prob(zs[r], mu0, s0).sum(-1)  # (K,B)
    if obj == "iwae":  # objectives.py:342-359 with the N1 shim
        lws = []
        for r in range(M):
            lpx_z = torch.stack(lpx_zs[r]).sum(0)
            lp = lpz(r)
            lws.append(lp + lpx_z.reshape(*lp.shape) - beta * lq(r))
        loss = -log_mean_exp(torch.cat(lws)).sum()
        out.update(loss=loss, lw=torch.cat(lws))
        return out
    if obj == "dreg":  # objectives.py:361-387; needs (K,)-shaped lpx (decoders keeping the K axis): the
        # restatement sums the (K*B,) rows over b, which is what view(K,-1).sum(-1) does for those decoders
        lws = []
        for r in range(M):
            lpx_z = torch.stack([x.reshape(K, -1).sum(-1) for x in lpx_zs[r]]).sum(0)  # (K,)
            lws.append(lpz(r).sum(-1) + lpx_z - lq(r).sum(-1))
        lw = torch.stack(lws)  # (M,K)
        with torch.no_grad():
            grad_wt = (lw - torch.logsumexp(lw, 1, keepdim=True)).exp()
        loss = -(grad_wt * lw).mean(0).sum()
        out.update(loss=loss, lw=lw)
        return out
    raise ValueError(obj)


def mixture_component_selection(mus, logvars, w_modalities):
    """mmvae_models.py:396-410, literal: num_samples = mus.shape[1]; chunk k = rows [start_k, end_k) of component k
    along dim 1, concatenated."""
    num_components, num_samples = mus.shape[0], mus.shape[1]
    st, en = mopoe_chunk_bounds(num_components, num_samples)
    mu_sel = torch.cat([mus[k, st[k]:en[k]] for k in range(w_modalities.shape[0])])
    lv_sel = torch.cat([logvars[k, st[k]:en[k]] for k in range(w_modalities.shape[0])])
    return mu_sel, lv_sel


def mopoe_mixing(mods, D):
    """MoPOE.modality_mixing / poe_fusion / moe_fusion mmvae_models.py:322-410.
    Prior expert (0,0) is appended LAST and only for the full subset (:386-389).
    QUIRK (reproduced): poe_fusion returns (1,B,D) tensors (:391-393) and modality_mixing stacks them with another
    unsqueeze(0) (:336-337) -> (S,1,B,D); mixture_component_selection therefore sees num_samples == 1 (:397), every
    chunk but the last is empty and the joint posterior of EVERY row is the last available subset (the full
    product incl. the prior expert when no modality is missing).  Called on a (S,B,D) stack the same function
    does produce contiguous batch chunks -- see mopoe_row_to_subset for that function-level contract."""
    M = len(mods)
    present = [i for i, m in enumerate(mods) if m["mu"] is not None]
    B = mods[present[0]]["mu"].shape[0]
    mus, lvs, kept = [], [], []
    for sub in mopoe_subsets(range(M)):
        if not all(i in present for i in sub):
            continue
        mu = torch.stack([mods[i]["mu"] for i in sub])
        lv = torch.stack([mods[i]["s"] for i in sub])
        if len(sub) == M:
            mu = torch.cat((mu, torch.zeros(1, B, D, dtype=mu.dtype, device=mu.device)), 0)
            lv = torch.cat((lv, torch.zeros(1, B, D, dtype=mu.dtype, device=mu.device)), 0)
        pm, pv = product_of_experts(mu, lv)
        mus.append(pm.unsqueeze(0))  # (1,B,D)  :391-393
        lvs.append(pv.unsqueeze(0))
        kept.append(sub)
    S = len(mus)
    mus_t, lvs_t = torch.stack(mus), torch.stack(lvs)  # (S,1,B,D)
    w = (1 / float(S)) * torch.ones(S)
    w = w / w.sum()
    mu_sel, var_sel = mixture_component_selection(mus_t, lvs_t, w)  # (1,B,D)
    row_map = torch.full((B,), S - 1, dtype=torch.int32)
    return mu_sel.squeeze(0), var_sel.squeeze(0), list(zip(kept, mus, lvs)), row_map


def mopoe_objective(mods, pz_logits, noise, beta=1.0, K=1):
    """MoPOE.objective mmvae_models.py:296-320 + forward :351-370 + weighted_group_kld objectives.py:184-201."""
    M = len(mods)
    D = pz_logits.shape[-1]
    mu0 = torch.zeros_like(pz_logits)
    _, s0 = prior_params(mu0, pz_logits)
    mu_j, var_j, subsets, row_map = mopoe_mixing(mods, D)
    noise = list(noise)
    zs, lpx_zs = [], []
    for m in mods:
        z = normal_rsample(mu_j, var_j, noise.pop(0))  # fresh eps per modality (:366)
        zs.append(z)
        lpx_zs.append(lpx_rows(m["ltype"], m["dec"](z), m["target"], m["lam"], K, m["dist"], m.get("mask_len")))
    dists = [(m["mu"], m["s"]) for m in mods] + [(mu_j, var_j)]
    klds = [kl_normal_normal(l, s, mu0, s0) for l, s in dists]
    w = (1 / len(dists)) * torch.ones(len(dists), dtype=mu_j.dtype, device=mu_j.device)
    gkl = (torch.stack(klds).sum(-1).mean(1) * w).sum()
    lpx = torch.stack(lpx_zs).sum(0).mean()
    loss = elbo(lpx, gkl, beta)
    ind = [-l / mods[i]["lam"] for i, l in enumerate(lpx_zs)]
    return {"loss": loss, "kld": gkl, "reconstruction_loss": ind, "joint": (mu_j, var_j), "z": zs,
            "row_map": row_map, "subsets": subsets}


def dmvae_objective(mods, pz_logits, noise, beta=1.0, K=1):
    """DMVAE.objective mmvae_models.py:436-465 + forward :467-503.  Noise order: joint, then per
    modality (shared, private, one fresh shared sample per other modality)."""
    M = len(mods)
    D = pz_logits.shape[-1]
    mu0 = torch.zeros_like(pz_logits)
    _, s0 = prior_params(mu0, pz_logits)
    sh = [(m["mu"][:, :D], m["s"][:, :D]) for m in mods]  # mmvae_base.py:155-156
    pr = [(m["mu"][:, D:], m["s"][:, D:]) for m in mods]
    mu_j, var_j = product_of_experts(torch.stack([a for a, _ in sh]), torch.stack([b for _, b in sh]))  # no prior
    noise = list(noise)
    z_joint = normal_rsample(mu_j, var_j, noise.pop(0))
    losses, ind, klds, zs = [], [], [], {"joint": z_joint, "shared": [], "private": [], "cross": []}
    for i, m in enumerate(mods):
        z_sh = normal_rsample(sh[i][0], sh[i][1], noise.pop(0))
        z_pr = normal_rsample(pr[i][0], pr[i][1], noise.pop(0))
        zs["shared"].append(z_sh)
        zs["private"].append(z_pr)
        lp = lambda loc: lpx_rows(m["ltype"], loc, m["target"], m["lam"], 1, m["dist"], m.get("mask_len"))
        lpx_z = lp(m["dec"](torch.cat([z_sh, z_pr], -1)))
        lpx_poe = lp(m["dec"](torch.cat([z_joint, z_pr], -1)))
        cross, kpriv = [], []
        for j in range(M):
            if j == i:
                continue
            z_c = normal_rsample(sh[j][0], sh[j][1], noise.pop(0))  # fresh, 1 sample (:499)
            zs["cross"].append(z_c)
            cross.append(lp(m["dec"](torch.cat([z_c, z_pr], -1))))
            one = torch.ones_like(pr[i][0][:1])
            kpriv.append(kl_normal_normal(pr[i][0], pr[i][1], torch.zeros_like(one), one))  # vae.py:190-196
        kld = kl_normal_normal(sh[i][0], sh[i][1], mu0, s0)
        kld_poe = kl_normal_normal(mu_j, var_j, mu0, s0)
        loss = elbo(lpx_z, kld.sum(-1), beta) + elbo(lpx_poe, kld_poe, beta) \
            + elbo(torch.stack(cross).sum(), torch.stack(kpriv).sum(-1), beta)
        losses.append(loss)
        ind.append(lpx_z)
        klds.append(kld)
    ind_r = [-(l).sum() / mods[i]["lam"] for i, l in enumerate(ind)]
    return {"loss": torch.stack(losses).sum(), "reconstruction_loss": ind_r,
            "kld": torch.stack(klds).mean(0).sum(), "joint": (mu_j, var_j), "z": zs}


# ----------------------------------------------------------------------------------------------------------
# Unimodal VAE: VAE.forward vae.py:99-119 + VAE.objective :264-281 + UnimodalObjective.elbo
# objectives.py:233-247.  The fixed VAE-level prior is N(0, 1) (_pz_params = zeros, ones, vae.py:159-162).
# ----------------------------------------------------------------------------------------------------------
def unimodal_elbo(mu, s, dist_name, noise, dec, target, ltype, beta=1.0, K=1, mask_len=None):
    """loss = -(lpx_z.sum(-1) - beta * kld.sum()).sum(): the scalar KL total is broadcast against EVERY decoder row
    (objectives.py:67), i.e. it is counted rows = K*B times -- reproduced.  Note VAE.objective calls
    calculate_loss without K (defaults to 1) while forward() used K=1 as well."""
    z = rsample(dist_name, mu, s, noise)  # (K,B,D)
    loc = dec(z.reshape(1, -1, z.shape[-1]))  # vae.py:112
    lpx_z = recon_logp(ltype, loc, target, K, dist_name, mask_len=mask_len)
    kld = kl_to_normal(dist_name, mu, s, torch.zeros_like(mu[:1]), torch.ones_like(mu[:1]))
    loss = elbo(lpx_z, kld, beta)
    return {"loss": loss, "kld": kld, "reconstruction_loss": lpx_z, "z": z}


# ----------------------------------------------------------------------------------------------------------
# f3: per-dimension KL tables of the analysis hook (utils.py:130-162 make_kl_df)
# ----------------------------------------------------------------------------------------------------------
def kl_table(family, locs, scales, loc0, scale0):
    """[KL(q_i || p) for i] + [0.5 (KL(q_i||q_j) + KL(q_j||q_i)) for i < j] through torch.distributions.kl_divergence
    (utils.py:399-405 -> torch kl.py closed forms), stacked (T, n, D)."""
    import itertools
    import torch.distributions as dist
    cls = dist.Laplace if family == "laplace" else dist.Normal
    qs = [cls(l, s) for l, s in zip(locs, scales)]
    pz = dist.Normal(loc0, scale0)
    rows = [dist.kl_divergence(q, pz) for q in qs]
    rows += [0.5 * (dist.kl_divergence(p, q) + dist.kl_divergence(q, p)) for p, q in itertools.combinations(qs, 2)]
    return torch.stack(rows)
