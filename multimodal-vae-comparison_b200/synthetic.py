"""Synthetic workloads and stand-in per-modality VAEs.

The hot path sits between the encoder outputs and the scalar loss, so benchmarks and parity tests replace
the reference's dense encoders/decoders by small stand-ins exposing the attribute surface the model plugins
read from ``models.vae.VAE`` (reference vae.py:121-196): ``enc``/``dec`` callables, ``qz_x``/``px_z``/``pz``
distribution classes, ``llik_scaling``, ``ltype``, ``n_latents``, ``private_latents``, ``prior_str``,
``modelName``, ``_pz_params`` and ``pz_params_private``.  Pure torch -- no kernels here.
"""
import math

import torch
import torch.distributions as dist
import torch.nn as nn
import torch.nn.functional as F

DIST_MAP = {"normal": dist.Normal, "gaussian": dist.Normal, "laplace": dist.Laplace}
ETA = 1e-6  # reference utils.py:254 Constants.eta


class LeafEncoder(nn.Module):
    """Returns fixed (mu, s) leaves -- the encoder contract of SURVEY 8: s = softmax(raw)+1e-6."""

    def __init__(self, data_dim, mu, s):
        super().__init__()
        self.data_dim = tuple(data_dim)
        self.mu = nn.Parameter(mu.clone())
        self.s = nn.Parameter(s.clone())

    def forward(self, x):
        return self.mu, self.s


class LinearEncoder(nn.Module):
    """flatten -> two Linear heads with the reference tail (encoders.py:49-54): mu, softmax(raw,-1)+1e-6."""

    def __init__(self, data_dim, out_dim, returns_raw_logvar=False):
        super().__init__()
        self.data_dim = tuple(data_dim)
        p = int(math.prod(self.data_dim))
        self.mu_layer = nn.Linear(p, out_dim)
        self.logvar_layer = nn.Linear(p, out_dim)
        # True: hand the raw output of the second head to the latent kernels, which apply the tail themselves
        self.returns_raw_logvar = bool(returns_raw_logvar)

    def forward(self, x):
        d = x["data"].float().reshape(x["data"].shape[0], -1)
        raw = self.logvar_layer(d)
        return self.mu_layer(d), raw if self.returns_raw_logvar else F.softmax(raw, dim=-1) + ETA


class LinearDecoder(nn.Module):
    """latents (K,B,Dz) -> Linear -> optional sigmoid+clamp (reference decoders.py:96-98) -> rows K*B
    (the CNN/FNN k-major convention).  Returns (mean, 0.75) like every reference decoder."""

    def __init__(self, in_dim, data_dim, squash, returns_logits=False):
        super().__init__()
        self.data_dim = tuple(data_dim)
        self.lin = nn.Linear(in_dim, int(math.prod(self.data_dim)))
        self.squash = squash
        # True: hand the pre-sigmoid logits to the likelihood kernel, which applies the tail itself (bce_logits)
        self.returns_logits = bool(returns_logits and squash)
        # rows of the (K, B, Dz) latents are decoded independently and come back k-major (the reference's CNN / FNN
        # decoders do the same, decoders.py:96-98, :400): the plugins may stack the latents of several likelihood terms
        # of this modality along K and decode them in ONE call (mmvae_models._fold_ok)
        self.folds_K = True
        # the likelihood scale every reference decoder returns (decoders.py:98 builds it from a host scalar on every
        # call -- a pageable H2D copy, which a CUDA-graph capture does not allow): a non-persistent buffer instead
        self.register_buffer("_scale", torch.tensor(0.75), persistent=False)

    def forward(self, z):
        z = z["latents"]
        d = self.lin(z)
        if self.squash and not self.returns_logits:
            d = torch.sigmoid(d).clamp(ETA, 1 - ETA)
        return d.reshape(-1, *self.data_dim), self._scale


class LeafDecoder(nn.Module):
    """Ignores z and hands out pre-made reconstruction leaves in call order (the measurement protocol of
    SURVEY 8d: decoders are replaced by leaf tensors so that no dense layer is timed)."""

    def __init__(self, data_dim, leaves):
        super().__init__()
        self.data_dim = tuple(data_dim)
        self.leaves = nn.ParameterList([nn.Parameter(l) for l in leaves])
        self._i = 0

    def reset(self):
        self._i = 0

    def forward(self, z):
        out = self.leaves[self._i % len(self.leaves)]
        self._i += 1
        return out, torch.tensor(0.75, device=out.device)


class StubVAE(nn.Module):
    """Attribute-compatible stand-in for reference models.vae.VAE (vae.py:121-196)."""

    def __init__(self, enc, dec, n_latents, ltype, private_latents=None, llik_scaling=1.0, prior_dist="normal",
                 id_name="mod_1"):
        super().__init__()
        self.enc, self.dec = enc, dec
        self.prior_str = prior_dist.lower()
        self.pz = self.px_z = self.qz_x = DIST_MAP[self.prior_str]
        self.prior_dist = self.post_dist = self.likelihood_dist = self.pz
        self.llik_scaling = llik_scaling
        self.data_dim = enc.data_dim
        self.private_latents = private_latents
        self.n_latents = n_latents
        self.total_latents = n_latents + (private_latents or 0)
        self._pz_params = nn.ParameterList([
            nn.Parameter(torch.zeros(1, self.total_latents), requires_grad=False),
            nn.Parameter(torch.ones(1, self.total_latents), requires_grad=False)])
        self._pz_params_private = None
        if private_latents is not None:
            self._pz_params_private = nn.ParameterList([
                nn.Parameter(torch.zeros(1, private_latents), requires_grad=False),
                nn.Parameter(torch.ones(1, private_latents), requires_grad=False)])
        self.modelName = id_name
        self.ltype = ltype

    @property
    def pz_params_private(self):
        return self._pz_params_private[0], \
            F.softmax(self._pz_params_private[1], dim=1) * self._pz_params_private[1].size(-1)

    @property
    def pz_params(self):
        return self._pz_params[0], F.softmax(self._pz_params[1], dim=1) * self._pz_params[1].size(-1)


# ------------------------------------------------------------------------------------------------------
# tensors of SURVEY 8d
# ------------------------------------------------------------------------------------------------------
def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_posterior(g, B, D):
    """mu ~ N(0,1); s = softmax(N(0,1), -1) + 1e-6."""
    mu = torch.randn(B, D, generator=g)
    s = F.softmax(torch.randn(B, D, generator=g), dim=-1) + ETA
    return mu, s


def make_noise(g, dist_name, shape):
    """eps ~ N(0,1); Laplace: u ~ U(-1+2^-23, 1) (torch laplace.py:75-79)."""
    if dist_name == "laplace":
        lo = torch.finfo(torch.float32).eps - 1
        return torch.rand(shape, generator=g) * (1 - lo) + lo
    return torch.randn(shape, generator=g)


def make_target(g, kind, B, data_dim):
    """image targets U(0,1); text targets one-hot over the last axis."""
    if kind == "onehot":
        idx = torch.randint(data_dim[-1], (B, *data_dim[:-1]), generator=g)
        return F.one_hot(idx, data_dim[-1]).float()
    return torch.rand(B, *data_dim, generator=g)


def make_recon(g, ltype, rows, data_dim):
    """recon = sigmoid(N(0,1)).clamp(1e-6, 1-1e-6) for bce/lprob, N(0,1) logits otherwise."""
    x = torch.randn(rows, *data_dim, generator=g)
    if ltype in ("bce", "lprob"):
        x = torch.sigmoid(x).clamp(ETA, 1 - ETA)
    return x


# name -> model kind, objective, K, latent D, private, modality specs (data_dim, ltype, target kind, dist, lam)
WORKLOADS = {
    # C1: MVAE (PoE) ELBO, CdSprites+ level 1
    "c1_poe_elbo_cdsprites_l1": dict(model="poe", obj="elbo", K=1, D=16, B=32, mods=[
        dict(data_dim=(3, 64, 64), ltype="bce", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(7, 27), ltype="category_ce", target="onehot", dist="normal", lam=1.0)]),
    # C2: MMVAE (MoE) IWAE K=30, CdSprites+ level 5  -- the configuration the metric is quoted on
    "c2_moe_iwae_cdsprites_l5": dict(model="moe", obj="iwae", K=30, D=16, B=256, mods=[
        dict(data_dim=(3, 64, 64), ltype="bce", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(45, 27), ltype="category_ce", target="onehot", dist="normal", lam=1.0)]),
    # C3: MoPoE ELBO trimodal SPRITES
    "c3_mopoe_elbo_sprites": dict(model="mopoe", obj="elbo", K=1, D=10, B=16, mods=[
        dict(data_dim=(8, 64, 64, 3), ltype="bce", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(9,), ltype="category_ce", target="onehot", dist="normal", lam=1.0),
        dict(data_dim=(4, 6), ltype="category_ce", target="onehot", dist="normal", lam=1.0)]),
    # C3 / configs[2], VILANRO shapes (reference configs/config_vilanro.yml): language (B,4,V) with V = 27, actions
    # (B,100,4,1), RGB (B,3,64,64), every likelihood optimal_sigma (objectives.py:502-509), D = 32, B = 64
    "c3_mopoe_elbo_vilanro": dict(model="mopoe", obj="elbo", K=1, D=32, B=64, mods=[
        dict(data_dim=(4, 27), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(100, 4, 1), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(3, 64, 64), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0)]),
    # the shipped config_vilanro.yml itself mixes with `poe` (MVAE: all 7 subsets, 21 likelihood terms)
    "c3_poe_elbo_vilanro": dict(model="poe", obj="elbo", K=1, D=32, B=64, mods=[
        dict(data_dim=(4, 27), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(100, 4, 1), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(3, 64, 64), ltype="optimal_sigma", target="uniform", dist="normal", lam=1.0)]),
    # C4: MMVAE DReG K=50, D=64, MNIST/SVHN-shaped Laplace likelihoods
    "c4_moe_dreg_mnistsvhn": dict(model="moe", obj="dreg", K=50, D=64, B=1024, mods=[
        dict(data_dim=(1, 28, 28), ltype="lprob", target="uniform", dist="laplace", lam=1.0),
        dict(data_dim=(3, 32, 32), ltype="lprob", target="uniform", dist="laplace", lam=784.0 / 3072.0)]),
    # C4 latent + combine only: same latent path, the four likelihood row vectors are synthetic (K*B,)
    # leaves, so that batches up to 64k fit one GPU; bytes/sample = 4*K*D*4 per drawn tensor + row scalars
    "c4_moe_dreg_latent_only": dict(model="moe", obj="dreg", K=50, D=64, B=16384, latent_only=True, mods=[
        dict(data_dim=(1, 28, 28), ltype="lprob", target="uniform", dist="laplace", lam=1.0),
        dict(data_dim=(3, 32, 32), ltype="lprob", target="uniform", dist="laplace", lam=784.0 / 3072.0)]),
    # C5: DMVAE ELBO, CUB shapes (bf16 in the bench)
    "c5_dmvae_elbo_cub": dict(model="dmvae", obj="elbo", K=1, D=16, private=10, B=256, mods=[
        dict(data_dim=(3, 64, 64), ltype="bce", target="uniform", dist="normal", lam=1.0),
        dict(data_dim=(246, 27), ltype="category_ce", target="onehot", dist="normal", lam=1.0)]),
}
