"""Drop-in model plugins MOE (MMVAE), POE (MVAE), MoPOE, DMVAE -- constructor, ``objective(batch) -> dict``,
``forward(inputs, K) -> VAEOutput``, ``modality_mixing``, ``pz_params`` as in reference models/mmvae_models.py, with
everything between the encoder outputs and the scalar loss running in the sm_100a kernels (ops.py).

Reference quirks are reproduced on purpose: the PoE variance is the Normal scale; MoE-ELBO drops rows
whose sum is exactly 0 and counts the total KL once per remaining row; MOE wraps its self-reconstruction in a Normal
whatever the likelihood family; the in-model MoPoE joint is the last available subset for every row; DReG weights are
a softmax over K of batch-summed log-weights.
"""
import functools
import itertools

import torch
import torch.distributions as dist

from . import ops
from .mmvae_base import TorchMMVAE, dist_code
from .ops import Draw


# ----------------------------------------------------------------------------------------------------------
# subset bookkeeping (host side, integer work)
# ----------------------------------------------------------------------------------------------------------
def poe_subsets(names):
    """All non-empty subsets grouped by size (reference utils.py:86-112 subsample_input_modalities; the reference
    order inside one size depends on PYTHONHASHSEED, the loss -- a plain sum over subsets -- does not)."""
    names = list(names)
    return [c for n in range(1, len(names) + 1) for c in itertools.combinations(names, n)]


def mopoe_subsets(names):
    """Reference mmvae_models.py:279-294 set_subsets: powerset without the empty set, by size then name."""
    names = list(names)
    return [tuple(sorted(c)) for n in range(1, len(names) + 1) for c in itertools.combinations(names, n)]


def subset_bitmasks(subsets, M):
    """uint32 expert masks; bit 31 flags the prior expert, which MoPoE appends only to the full subset
    (mmvae_models.py:386-389)."""
    vals = []
    for sub in subsets:
        m = 0
        for i in sub:
            m |= 1 << int(i)
        if len(sub) == M:
            m |= 1 << 31
        vals.append(m - (1 << 32) if m >= (1 << 31) else m)  # uint32 bit pattern in int32 storage
    return torch.tensor(vals, dtype=torch.int32)


@functools.lru_cache(maxsize=256)
def mopoe_chunk_bounds(num_components, num_samples):
    """Chunk bounds of MoPOE.mixture_component_selection (mmvae_models.py:396-410) with the weights of :339 and
    reweight_weights :377.  Index work must be bit exact, and floor(B * w_k) depends on how the S fp32 weights
    were summed (S = 31 flips), so the weights are evaluated with the same torch fp32 host ops the reference uses."""
    w = (1 / float(num_components)) * torch.ones(num_components)
    w = w / w.sum()
    starts, ends = [], []
    for k in range(num_components):
        i0 = 0 if k == 0 else ends[k - 1]
        i1 = num_samples if k == num_components - 1 else i0 + int(torch.floor(num_samples * w[k]))
        starts.append(i0)
        ends.append(i1)
    ends[-1] = num_samples
    return tuple(starts), tuple(ends)


def mopoe_row_subset_map(num_components, num_samples):
    """Row -> mixture component of mixture_component_selection applied to a (S, B, D) stack."""
    st, en = mopoe_chunk_bounds(num_components, num_samples)
    idx = torch.empty(num_samples, dtype=torch.int32)
    for k in range(num_components):
        idx[st[k]:en[k]] = k
    return idx


def mopoe_inmodel_component(num_components):
    """Component that MoPOE.modality_mixing actually selects for EVERY row: it stacks (1,B,D) tensors into
    (S,1,B,D), so mixture_component_selection sees num_samples == 1 (mmvae_models.py:336-337, :397)."""
    _, en = mopoe_chunk_bounds(num_components, 1)
    return next(k for k in range(num_components) if en[k] == 1)


def _family(vae):
    return "laplace" if getattr(vae, "px_z", dist.Normal) is dist.Laplace else "normal"


def _first(t):
    return t[0] if isinstance(t, (tuple, list)) else t


def _unmasked(vae):
    """A text decoder that declares ``returns_unmasked = True`` hands over its output BEFORE the reference's
    "zero for padded area" multiply (decoders.py:722); category_ce then applies the padding mask inside its kernel."""
    return bool(getattr(vae.dec, "returns_unmasked", False))


def _fold_ok(vae, masks):
    """Term folding: a decoder that declares ``folds_K = True`` decodes the rows of (K, B, Dz) latents independently and
    returns them k-major -- what the reference's CNN / FNN decoders do (decoders.py:96-98, :400; NOT its transformer text
    decoder, which drops K).  The latents of all likelihood terms of that modality (the 2^M-1 subset draws of
    MVAE, the shared / joint / cross draws of DMVAE) are then stacked along K: ONE decoder call and ONE likelihood launch
    over (terms * B) rows instead of one per term -- same values (row r reads target row r % B), larger grids, one
    ramp-up and tail per modality.  Only without padding masks (they differ between the terms of MVAE), and not for
    ``optimal_sigma``, whose sigma is the RMS over all elements of ONE reconstruction (objectives.py:502-509)."""
    return bool(getattr(vae.dec, "folds_K", False)) and masks is None and vae.ltype != "optimal_sigma"


def _ltype(vae):
    """Likelihood the kernels evaluate for this VAE.  A decoder that declares ``returns_logits = True`` hands over
    pre-sigmoid logits; its ``bce`` is then evaluated by the fused ``bce_logits`` kernel, which folds the reference
    decoder tail sigmoid(.).clamp(1e-6, 1-1e-6) (decoders.py:96-97) into the row reduction."""
    if vae.ltype == "bce" and getattr(vae.dec, "returns_logits", False):
        return "bce_logits"
    return vae.ltype


# ----------------------------------------------------------------------------------------------------------
class MOE(TorchMMVAE):
    """MMVAE, mixture of experts (reference mmvae_models.py:10-131)."""

    def __init__(self, vaes, n_latents: int, obj_config: dict, model_config=None):
        super().__init__(vaes, n_latents, **obj_config)
        self.model_config = model_config
        self.modelName = "moe"

    @staticmethod
    def _cross_source(M, r):
        """mmvae_models.py:112-116: cross_px_zs[target] is overwritten by every source != target -> the last one."""
        others = [s for s in range(M) if s != r]
        return others[-1] if others else None

    def _sample(self, names, enc, K, prior, through_z):
        """Returns (mu, s, s_raw, codes, z, lq, lpz); s_raw: `s` holds raw encoder logits (tail fused into the kernels)."""
        mu, s, raw = self._stack_raw(enc, names, "shared")
        M, B, D = mu.shape
        if raw and not ops.moe_rk_supported(M, D):  # the fused tail lives in the flat MoE kernels only
            s, raw = self._enc_tail(s), False
        codes = [dist_code(self.vaes[n]) for n in names]
        eps = torch.stack([self._noise("laplace" if c else "normal", (K, B, D), mu.device) for c in codes])
        if raw:
            z, lq, lpz, _ = ops.moe_logdens_tail(mu, s, prior[0], prior[1], eps, codes, through_z)
        else:
            z, lq, lpz = ops.moe_logdens(mu, s, prior[0], prior[1], eps, codes, through_z)
        return mu, s, raw, codes, z, lq, lpz

    def objective(self, data):
        self._require_all(data)
        names = list(self.vaes.keys())
        M, K, beta = len(names), self.K, self.obj_fn.beta
        if M > 2:
            # reference MOE.forward keeps ONE cross reconstruction per target (cross_px_zs[target] is overwritten by every
            # source, mmvae_models.py:112-116) and its objective mis-shapes for M = 3: only M <= 2 is defined
            raise ValueError("MOE.objective is only defined for at most two modalities (the reference's cross-modal "
                             "bookkeeping, mmvae_models.py:112-116, breaks for M >= 3); got %d" % M)
        enc = self.encode(data, raw_ok=True)
        obj = self.obj_fn.obj_name
        if obj == "elbo":
            if K != 1:
                raise ValueError("MOE elbo needs K == 1 (the reference breaks for K > 1, mmvae_models.py:62)")
            # the ELBO branch never touches the learnable prior (fixed N(0,1) VAE prior, :45): keep its grad None
            prior = tuple(p.detach() for p in self.pz_params)
            mu, s, raw, codes, z, lq, _ = self._sample(names, enc, 1, prior, through_z=False)
            kls = ops.latent_draws(mu, s, None, None, None,
                                   [Draw(mods=(m,), direct=True, laplace=bool(codes[m]), kl_mode=2, width=mu.shape[-1])
                                    for m in range(M)], s_raw=raw)
            total, rows_log, n_keep = 0.0, [], 0.0
            for r, name in enumerate(names):
                vae = self.vaes[name]
                self.obj_fn.set_ltype(vae.ltype)
                loc = _first(vae.dec({"latents": z[r], "masks": data[name]["masks"]}))
                # self reconstruction: always a Normal likelihood (dist.Normal(*px_z), :105-107)
                S_self, rows_self = self.obj_fn.lpx_weighted_sum(loc, data[name], vae.llik_scaling, w_const=-1.0 / M,
                                                                 ltype=_ltype(vae), unmasked=_unmasked(vae), family="normal")
                total = total + S_self
                n_keep = n_keep + (S_self != 0).float()
                rows_log.append(rows_self)
                src = self._cross_source(M, r)
                if src is None:
                    continue
                loc = _first(vae.dec({"latents": z[src], "masks": data[name]["masks"]}))
                lwt = lq[src, r, 0] - lq[src, src, 0].detach()  # sum_d log q_r(z_s) - log q_s(z_s)   (:56-59)
                iw = lwt.exp()
                S_cross, rows_cross = self.obj_fn.lpx_weighted_sum(loc, data[name], vae.llik_scaling,
                                                                   w_rows=-iw / M, ltype=_ltype(vae), unmasked=_unmasked(vae), family=_family(vae))
                total = total + S_cross
                n_keep = n_keep + (S_cross != 0).float()  # rows summing to exactly 0 are dropped (:73)
                rows_log.append(iw.detach() * rows_cross)
            kld = torch.stack([k["kl"] for k in kls])  # (M,B)
            loss = total + (beta / M) * n_keep * kld.sum()  # total KL once per kept row (objectives.py:67)
            return {"loss": loss, "reconstruction_loss": torch.stack(rows_log), "kld": kld}
        mu, s, _, codes, z, lq, lpz = self._sample(names, enc, K, self._prior(), through_z=True)
        B = mu.shape[1]
        L = 1 if M == 1 else 2
        # the row kernels write straight into slices of one (M, L, K*B) buffer: the list of row vectors IS the stacked
        # tensor the reference builds with torch.stack (no concatenation copies)
        buf = torch.empty((M, L, K * B), dtype=torch.float32, device=mu.device)
        rows = []
        for r, name in enumerate(names):
            vae = self.vaes[name]
            self.obj_fn.set_ltype(vae.ltype)
            rows.append(self.obj_fn.lpx_rows(_first(vae.dec({"latents": z[r], "masks": data[name]["masks"]})), data[name],
                                             vae.llik_scaling, ltype=_ltype(vae), unmasked=_unmasked(vae), family="normal", out=buf[r, 0]))
            src = self._cross_source(M, r)
            if src is not None:
                rows.append(self.obj_fn.lpx_rows(_first(vae.dec({"latents": z[src], "masks": data[name]["masks"]})),
                                                 data[name], vae.llik_scaling, ltype=_ltype(vae), unmasked=_unmasked(vae), family=_family(vae),
                                                 out=buf[r, 1]))
        # both combines read the row vectors through a pointer table; "lpx_z" is the same memory, for logging
        d = {"lpz": lpz, "lq": lq, "lpx_z": buf.view(M, L, K, B), "lpx_rows": rows}
        return self.obj_fn.calculate_loss(d)

    def modality_mixing(self, mods):
        return mods

    def forward(self, x, K=1):
        """mmvae_models.py:80-117."""
        missing, filled = self.get_missing_modalities(x)
        assert len(filled) > 0, "at least one modality must be present for forward call"
        enc = self.encode(x)
        prior = tuple(p.detach() for p in self.pz_params)
        mu, s, _, codes, z, _, _ = self._sample(filled, enc, K, prior, through_z=False)
        zs, qzs, px_zs, cross = {}, {}, {}, {}
        for i, name in enumerate(filled):
            qzs[name] = self.vaes[name].qz_x(mu[i], s[i])
            zs[name] = {"latents": z[i], "masks": x[name]["masks"]}
        for name in missing:
            qzs[name] = None
        for name in filled:
            px_zs[name] = dist.Normal(*self.vaes[name].dec(zs[name]))
        for name in missing:
            zs[name] = {"latents": zs[filled[0]]["latents"], "masks": x[name]["masks"]}
            px_zs[name] = dist.Normal(*self.vaes[name].dec(zs[name]))
        for modality, zd in zs.items():
            for mod_vae, vae in self.vaes.items():
                if mod_vae != modality:
                    cross[mod_vae] = {modality: vae.px_z(*vae.dec({"latents": zd["latents"], "masks": x[mod_vae]["masks"]}))}
        return self.make_output_dict(qzs, px_zs, zs, cross_decoder_dist=cross)


# ----------------------------------------------------------------------------------------------------------
class POE(TorchMMVAE):
    """MVAE, product of experts (reference mmvae_models.py:134-250)."""

    def __init__(self, vaes, n_latents: int, obj_config: dict, model_config=None):
        super().__init__(vaes, n_latents, **obj_config)
        self.model_config = model_config
        self.modelName = "poe"
        for vae in self.vaes.values():
            assert vae.prior_str in ["normal", "gaussian"], "POE only works with gaussian priors! Adjust the config"

    def objective(self, mods):
        """mmvae_models.py:159-187: sum over all 2^M-1 modality subsets of the ELBO of the PoE posterior.  One
        kernel fuses the experts of every subset (prior expert included), draws z and reduces the KL rows; each
        (subset, modality) likelihood is a single fused value+gradient pass."""
        self._require_all(mods)
        names = list(self.vaes.keys())
        M, beta, D = len(names), self.obj_fn.beta, self.n_latents
        enc = self.encode(mods, raw_ok=True)  # each encoder runs once; the reference re-runs it per subset, same outputs
        mu, s, raw = self._stack_raw(enc, names, "shared")
        B = mu.shape[1]
        subsets = poe_subsets(range(M))
        eps = torch.cat([self._noise("normal", (1, B, D), mu.device).reshape(-1) for _ in subsets])
        mu0, s0 = self._prior()
        res = ops.latent_draws(mu, s, mu0, s0, eps,
                               [Draw(mods=sub, prior=True, kl_mode=1, width=D, K=1) for sub in subsets], s_raw=raw)
        terms = []
        rec_log = [None] * M
        folded = [_fold_ok(self.vaes[n], mods[n]["masks"]) for n in names]
        for i, name in enumerate(names):
            if not folded[i]:
                continue
            vae = self.vaes[name]
            self.obj_fn.set_ltype(vae.ltype)
            # the packed z of the S subset draws IS the (S, B, D) latent stack: one decoder call, one likelihood launch
            z_all = torch.cat([res[a]["z"] for a in range(len(subsets))], 0)
            loc = _first(vae.dec({"latents": z_all, "masks": None}))
            S, rows = self.obj_fn.lpx_weighted_sum(loc, mods[name], vae.llik_scaling, w_const=-1.0, ltype=_ltype(vae),
                                                   family=_family(vae), defer=True)
            terms.append(S)
            if i < len(subsets):  # logging quirk: modality m is paired with subset m (:179-180)
                rec_log[i] = ops.reduce_sum(rows[i * B:(i + 1) * B], -1.0 / vae.llik_scaling)
        for a, sub in enumerate(subsets):
            z = res[a]["z"]
            for i, name in enumerate(names):
                if folded[i]:
                    continue
                vae = self.vaes[name]
                self.obj_fn.set_ltype(vae.ltype)
                masks = mods[name]["masks"] if i in sub else None
                loc = _first(vae.dec({"latents": z, "masks": masks}))
                S, rows = self.obj_fn.lpx_weighted_sum(loc, mods[name], vae.llik_scaling, w_const=-1.0,
                                                       ltype=_ltype(vae), family=_family(vae), defer=True,
                                                       # (the decoder of a modality outside the subset runs without a
                                                       # mask, :176 -- nothing to fuse there)
                                                       unmasked=_unmasked(vae) and masks is not None)
                terms.append(S)
                if i == a:  # logging quirk: "mod == 'mod_{m+1}'" with m the subset index (:179-180)
                    rec_log[i] = ops.reduce_sum(rows, -1.0 / vae.llik_scaling)
        # loss = sum_A -(sum_b sum_m lam_m lpx - beta sum_b KL_A), "kld" = sum_b mean_A KL_A[b]: one launch
        loss, kld = ops.elbo_combine(terms, res.kl_packed, [beta] * len(subsets), [1.0 / len(subsets)] * len(subsets))
        return {"loss": loss, "reconstruction_loss": [r for r in rec_log if r is not None], "kld": kld}

    def modality_mixing(self, x, K=None):
        """mmvae_models.py:210-232: (mu, var-as-scale, {mod: Normal(mu_m, s_m)}) of the present modalities + prior."""
        enc = self.encode(x)
        present = [n for n in self.vaes.keys() if n in enc and enc[n]["shared"] is not None]
        mu, s = self._stack(enc, present, "shared")
        res = ops.latent_draws(mu, s, None, None, None,
                               [Draw(mods=tuple(range(len(present))), prior=True, width=self.n_latents,
                                     want_params=True)])
        single = {n: dist.Normal(mu[i], s[i]) for i, n in enumerate(present)}
        return res[0]["loc"], res[0]["scale"], single

    def forward(self, inputs, K=1):
        """mmvae_models.py:189-208."""
        enc = self.encode(inputs)
        present = [n for n in self.vaes.keys() if n in enc and enc[n]["shared"] is not None]
        mu, s = self._stack(enc, present, "shared")
        B, D = mu.shape[1], self.n_latents
        eps = self._noise("normal", (K, B, D), mu.device).reshape(-1)
        res = ops.latent_draws(mu, s, None, None, eps,
                               [Draw(mods=tuple(range(len(present))), prior=True, width=D, K=K, want_params=True)])[0]
        qz_x = dist.Normal(res["loc"], res["scale"])
        z = res["z"]
        single = {n: dist.Normal(mu[i], s[i]) for i, n in enumerate(present)}
        px_d = {mod: vae.px_z(*vae.dec({"latents": z, "masks": inputs[mod]["masks"]})) for mod, vae in self.vaes.items()}
        qz_d = {key: qz_x for key in inputs.keys()}
        z_d = {key: {"latents": z, "masks": inputs[key]["masks"]} for key in inputs.keys()}
        return self.make_output_dict(single, px_d, z_d, joint_dist=qz_d)

    def prior_expert(self, size, use_cuda=False):
        """mmvae_models.py:235-250 (kept for API compatibility; the kernels add the prior expert themselves)."""
        dev = self._pz_params[0].device if use_cuda else "cpu"
        return torch.zeros(size, device=dev), torch.zeros(size, device=dev)


# ----------------------------------------------------------------------------------------------------------
class MoPOE(TorchMMVAE):
    """Mixture of products of experts, generalised multimodal ELBO (reference mmvae_models.py:253-410)."""

    def __init__(self, vaes, n_latents: int, obj_config: dict, model_config=None):
        super().__init__(vaes, n_latents, **obj_config)
        self.model_config = model_config
        self.modelName = "mopoe"
        self.subsets = self.set_subsets()
        self.weights = None
        self._mask_cache = {}

    def set_subsets(self):
        names = list(self.vaes.keys())
        return {"_".join(sub): [self.vaes[n] for n in sub] for sub in mopoe_subsets(names)}

    def _available(self, present_idx):
        """Subsets (as index tuples, reference order) all of whose modalities are present."""
        names = list(self.vaes.keys())
        idx = {n: i for i, n in enumerate(names)}
        subs = [tuple(idx[n] for n in sub) for sub in mopoe_subsets(names)]
        return [s for s in subs if all(i in present_idx for i in s)]

    def _row_masks(self, subs, B, device):
        """Expert bitmask of every row for the joint the model selects (see mopoe_inmodel_component)."""
        M = len(self.vaes)
        key = (tuple(subs), B, str(device))
        if key not in self._mask_cache:
            comp = mopoe_inmodel_component(len(subs))
            bits = 0
            for i in subs[comp]:
                bits |= 1 << i
            if len(subs[comp]) == M:
                bits |= 1 << 31
            if bits >= 1 << 31:
                bits -= 1 << 32  # two's complement for the int32 storage that backs the uint32 view
            self._mask_cache[key] = torch.full((B,), bits, dtype=torch.int32, device=device)
        return self._mask_cache[key]

    def objective(self, mods):
        """mmvae_models.py:296-320 + weighted_group_kld objectives.py:184-201."""
        self._require_all(mods)
        names = list(self.vaes.keys())
        M, beta, D = len(names), self.obj_fn.beta, self.n_latents
        enc = self.encode(mods, raw_ok=True)
        mu, s, raw = self._stack_raw(enc, names, "shared")
        B = mu.shape[1]
        Bt = self._batch_total(B)  # batch means are global under batch sharding
        K = 1  # objective() calls forward(mods) with the default K (:305)
        subs = self._available(set(range(M)))
        row_masks = self._row_masks(subs, B, mu.device)
        eps = torch.cat([self._noise("normal", (K, B, D), mu.device).reshape(-1) for _ in names])
        mu0, s0 = self._prior()
        draws = [Draw(rowmask=True, kl_mode=1 if i == 0 else 0, width=D, K=K) for i in range(M)]
        draws += [Draw(mods=(i,), direct=True, kl_mode=1, width=D) for i in range(M)]
        res = ops.latent_draws(mu, s, mu0, s0, eps, draws, row_masks, s_raw=raw)
        terms, ind = [], []
        for i, name in enumerate(names):
            vae = self.vaes[name]
            self.obj_fn.set_ltype(vae.ltype)
            loc = _first(vae.dec({"latents": res[i]["z"], "masks": mods[name]["masks"]}))
            S, rows = self.obj_fn.lpx_weighted_sum(loc, mods[name], vae.llik_scaling, w_const=-1.0 / Bt,
                                                   ltype=_ltype(vae), unmasked=_unmasked(vae), family=_family(vae),
                                                   defer=vae.ltype != "optimal_sigma")
            terms.append(S)
            ind.append(-rows / vae.llik_scaling)
        # packed KL rows: joint (draw 0) + the M unimodal posteriors; group KL = sum / ((M+1) * B) (objectives.py:184-201)
        c = 1.0 / ((M + 1) * Bt)
        loss, gkl = ops.elbo_combine(terms, res.kl_packed, [beta * c] * (M + 1), [c] * (M + 1))
        return {"loss": loss, "reconstruction_loss": ind, "kld": gkl}

    def modality_mixing(self, input_batch):
        """mmvae_models.py:322-349: {'modalities', 'joint': [mu, var], 'subsets': {key: [mu (1,B,D), var]}}."""
        enc = self.encode(input_batch)
        names = list(self.vaes.keys())
        present = [i for i, n in enumerate(names) if n in enc and enc[n]["shared"] is not None]
        pnames = [names[i] for i in present]
        mu, s = self._stack(enc, pnames, "shared")
        pos = {g: l for l, g in enumerate(present)}
        subs = self._available(set(present))
        M = len(names)
        draws = [Draw(mods=tuple(pos[i] for i in sub), prior=(len(sub) == M), width=self.n_latents, want_params=True)
                 for sub in subs]
        res = ops.latent_draws(mu, s, None, None, None, draws)
        comp = mopoe_inmodel_component(len(subs))
        self.weights = (1 / float(len(subs))) * torch.ones(len(subs), device=mu.device)
        latents = {"modalities": {n: {"shared": enc[n]["shared"], "private": enc[n]["private"]} for n in enc},
                   "joint": [res[comp]["loc"], res[comp]["scale"]],
                   "subsets": {"_".join(names[i] for i in sub): [res[k]["loc"].unsqueeze(0), res[k]["scale"].unsqueeze(0)]
                               for k, sub in enumerate(subs)}}
        return latents

    def forward(self, inputs, K=1):
        """mmvae_models.py:351-370."""
        latents = self.modality_mixing(inputs)
        mu_j, var_j = latents["joint"]
        qz_d, px_d, z_d, qz_joint = {}, {}, {}, {}
        for mod, vae in self.vaes.items():
            sh = latents["modalities"][mod]["shared"] if mod in latents["modalities"] else None
            qz_d[mod] = dist.Normal(*sh) if sh is not None else None
            qz_joint[mod] = dist.Normal(mu_j, var_j)
            z = mu_j + var_j * self._noise("normal", (K, *mu_j.shape), mu_j.device)
            z_d[mod] = {"latents": z, "masks": inputs[mod]["masks"]}
            px_d[mod] = vae.px_z(*vae.dec(z_d[mod]))
        return self.make_output_dict(qz_d, px_d, z_d, qz_joint)

    def reweight_weights(self, w):
        """mmvae_models.py:377-378."""
        return w / w.sum()

    def poe_fusion(self, mus, logvars):
        """mmvae_models.py:385-394 on a (m, B, D) stack of the subset's experts: the prior expert (0, 0) is appended
        only when the subset holds every modality; returns [(1,B,D) mu, (1,B,D) var-as-scale]."""
        m = mus.shape[0]
        res = ops.latent_draws(mus, logvars, None, None, None,
                               [Draw(mods=tuple(range(m)), prior=(m == len(self.vaes)), width=mus.shape[-1],
                                     want_params=True)])
        return [res[0]["loc"].unsqueeze(0), res[0]["scale"].unsqueeze(0)]

    def moe_fusion(self, mus, logvars, weights):
        """mmvae_models.py:380-383."""
        return self.mixture_component_selection(mus, logvars, self.reweight_weights(weights))

    def mixture_component_selection(self, mus, logvars, w_modalities=None):
        """mmvae_models.py:396-410 on a (S, n, ...) stack: contiguous chunks of dim 1, chunk k from component k."""
        S, n = mus.shape[0], mus.shape[1]
        st, en = mopoe_chunk_bounds(S, n)
        return [torch.cat([mus[k, st[k]:en[k]] for k in range(S)]), torch.cat([logvars[k, st[k]:en[k]] for k in range(S)])]


# ----------------------------------------------------------------------------------------------------------
class DMVAE(TorchMMVAE):
    """Private-shared disentangled multimodal VAE (reference mmvae_models.py:413-530)."""

    def __init__(self, vaes, n_latents: int, obj_config: dict, model_config=None):
        super().__init__(vaes, n_latents, **obj_config)
        self.model_config = model_config
        self.modelName = "dmvae"
        assert self.latent_factorization, "DMVAE requires private_latents in the config"

    def _draws(self, names, K):
        """Draw list in the reference's rsample order: joint, then per modality shared, private and one
        fresh single-sample shared draw from every OTHER modality."""
        M, D = len(names), self.n_latents
        draws = [Draw(mods=tuple(range(M)), prior=False, kl_mode=1, width=D, K=K, want_params=True)]
        index = {"joint": 0, "shared": [], "private": [], "cross": []}
        for i, n in enumerate(names):
            pv = self.vaes[n].private_latents
            index["shared"].append(len(draws))
            draws.append(Draw(mods=(i,), direct=True, kl_mode=1, col0=0, width=D, K=K))
            index["private"].append(len(draws))
            draws.append(Draw(mods=(i,), direct=True, kl_mode=2, col0=D, width=pv, K=K))
            cr = {}
            for j in range(M):
                if j != i:
                    cr[j] = len(draws)
                    draws.append(Draw(mods=(j,), direct=True, col0=0, width=D, K=1))
            index["cross"].append(cr)
        return draws, index

    def _run(self, mods, K, raw_ok=False):
        names = list(self.vaes.keys())
        enc = self.encode(mods, raw_ok=raw_ok)
        mu, s, raw = self._stack_raw(enc, names, "full")
        B = mu.shape[1]
        draws, index = self._draws(names, K)
        eps = torch.cat([self._noise("normal", (d.K, B, d.width), mu.device).reshape(-1) for d in draws])
        mu0, s0 = self._prior()
        res = ops.latent_draws(mu, s, mu0, s0, eps, draws, s_raw=raw)
        return names, enc, mu, s, res, index

    def objective(self, mods):
        """mmvae_models.py:436-465: per modality three ELBO terms (own shared, joint, cross)."""
        self._require_all(mods)
        beta = self.obj_fn.beta
        names, enc, mu, s, res, index = self._run(mods, 1, raw_ok=True)
        M = len(names)
        z_joint = res[0]["z"]
        terms, ind = [], []
        # packed KL rows: joint, then (shared_i, private_i) per modality -- their coefficients in the loss / in "kld"
        kc, kg = [M * beta], [0.0]
        for i, name in enumerate(names):
            vae = self.vaes[name]
            self.obj_fn.set_ltype(vae.ltype)
            z_sh, z_pr = res[index["shared"][i]]["z"], res[index["private"][i]]["z"]
            fam, lam, masks = _family(vae), vae.llik_scaling, mods[name]["masks"]

            def term(z_a):
                loc = _first(vae.dec({"latents": torch.cat([z_a, z_pr], -1), "masks": masks}))
                return self.obj_fn.lpx_weighted_sum(loc, mods[name], lam, w_const=-1.0, ltype=_ltype(vae), unmasked=_unmasked(vae), family=fam,
                                                    defer=True)

            if _fold_ok(vae, masks):  # own shared, joint and cross latents as ONE (terms, B, Dz) stack
                zs = [z_sh, z_joint] + [res[di]["z"] for di in index["cross"][i].values()]
                z_all = torch.cat([torch.cat([z_a, z_pr], -1) for z_a in zs], 0)
                loc = _first(vae.dec({"latents": z_all, "masks": None}))
                S_all, rows_all = self.obj_fn.lpx_weighted_sum(loc, mods[name], lam, w_const=-1.0, ltype=_ltype(vae),
                                                               family=fam, defer=True)
                terms.append(S_all)
                rows1 = rows_all[:z_sh.shape[1]]
            else:
                S1, rows1 = term(z_sh)
                terms += [S1, term(z_joint)[0]] + [term(res[di]["z"])[0] for di in index["cross"][i].values()]
            # -(lpx - beta KL_shared) - (lpx_joint - beta KL_joint) - sum_cross (lpx_cross - beta KL_private)  (:455-463)
            kc += [beta, beta * len(index["cross"][i])]
            kg += [1.0 / M, 0.0]  # "kld" = sum_b mean_i KL_shared_i[b]
            ind.append(ops.reduce_sum(rows1, -1.0 / lam))
        loss, kld = ops.elbo_combine(terms, res.kl_packed, kc, kg)
        return {"loss": loss, "reconstruction_loss": ind, "kld": kld}

    def modality_mixing(self, mods):
        return mods

    def forward(self, x, K=1):
        """mmvae_models.py:467-503 (all modalities present)."""
        self._require_all(x)
        names, enc, mu, s, res, index = self._run(x, K)
        D = self.n_latents
        joint_d = self.qz_x(res[0]["loc"], res[0]["scale"])
        joint_dist, qz_xs, qz_private, zss, px_zs, joint_px_zs, cross_px_zs = {}, {}, {}, {}, {}, {}, {}
        for i, mod in enumerate(names):
            vae = self.vaes[mod]
            joint_dist[mod] = joint_d
            qz_xs[mod] = self.qz_x(mu[i][:, :D], s[i][:, :D])
            qz_private[mod] = self.qz_x(mu[i][:, D:D + vae.private_latents], s[i][:, D:D + vae.private_latents])
            z_sh, z_pr = res[index["shared"][i]]["z"], res[index["private"][i]]["z"]
            zss[mod] = {"latents": z_sh, "masks": x[mod]["masks"]}
            dec = lambda za: vae.px_z(*vae.dec({"latents": torch.cat([za, z_pr], -1), "masks": x[mod]["masks"]}))
            px_zs[mod] = dec(z_sh)
            joint_px_zs[mod] = dec(res[0]["z"])
            cross_px_zs[mod] = {names[j]: dec(res[di]["z"]) for j, di in index["cross"][i].items()}
        return self.make_output_dict(qz_xs, px_zs, zss, joint_dist, qz_private, None, joint_px_zs, cross_px_zs)
