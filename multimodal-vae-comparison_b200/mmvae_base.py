"""Base class of the drop-in multimodal VAE plugins -- same constructor, attributes, state_dict keys and method
surface as reference models/mmvae_base.py:12-240 (``TorchMMVAE``); the math between the encoder outputs and the loss
runs in the sm_100a kernels."""
import abc

import numpy as np
import torch
import torch.distributions as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .objectives import MultimodalObjective
from .output_storage import VAEOutput

_FIELDS = ("encoder_dist", "decoder_dist", "latent_samples", "joint_dist", "enc_dist_private", "dec_dist_private",
           "joint_decoder_dist", "cross_decoder_dist")


def dist_code(vae) -> int:
    """0 normal / 1 laplace from the VAE's posterior class (reference vae.py:142-147)."""
    q = getattr(vae, "qz_x", dist.Normal)
    if q is dist.Laplace:
        return 1
    if q is dist.Normal:
        return 0
    raise NotImplementedError("posterior family %s is outside the accelerated path (normal / laplace only)" % q)


class TorchMMVAE(nn.Module):
    def __init__(self, vaes, n_latents: int, obj: str, beta=1, K=1):
        super().__init__()
        self.vaes = nn.ModuleDict(vaes)
        self.modelName = "TorchMMVAE"
        self.qz_x = dist.Normal
        self.px_z = dist.Normal
        self.pz = dist.Normal
        self.n_latents = n_latents
        self.K = K
        self.obj_fn = MultimodalObjective(obj, beta)
        # same parameter names / shapes as the reference (mmvae_base.py:35-38) so its checkpoints load unchanged
        self._pz_params = nn.ParameterList([
            nn.Parameter(torch.zeros(1, self.n_latents), requires_grad=False),
            nn.Parameter(torch.zeros(1, self.n_latents), requires_grad=True)])
        self.set_likelihood_scales()
        # test / reproducibility hook: callable(kind: "normal"|"laplace", shape) -> noise tensor.  None = draw on device
        self.noise_source = None
        # batch-sharded training (parallel.py): process group and the global batch size seen by batch-mean terms
        self.group = None
        self.global_batch = None

    # ---- reference surface ---------------------------------------------------------------------------------
    def set_likelihood_scales(self):
        """mmvae_base.py:41-47."""
        min_dim = min(int(np.prod(vae.enc.data_dim)) for vae in self.vaes.values())
        for vae in self.vaes.values():
            if vae.llik_scaling == "auto":
                vae.llik_scaling = min_dim / int(np.prod(vae.enc.data_dim))
            else:
                vae.llik_scaling = float(vae.llik_scaling)

    @property
    def pz_params(self):
        """(mu0, softmax(_pz_params[1], 1) * D) -- mmvae_models.py:28-30 and siblings."""
        return self._pz_params[0], F.softmax(self._pz_params[1], dim=1) * self._pz_params[1].size(-1)

    def _prior(self):
        """pz_params for the kernels: same values as the `pz_params` property, the softmax*D evaluated by one tiny
        kernel each way (ops.prior_scale) instead of four eager ones."""
        if self._pz_params[1].is_cuda:
            peer = self.group if hasattr(self.group, "bufs_dev") else None  # parallel.PeerGroup: fused gradient sync
            return self._pz_params[0], ops.prior_scale(self._pz_params[1], peer)
        return self.pz_params

    @property
    def latent_factorization(self):
        return any(v.private_latents is not None for v in self.vaes.values())

    def add_vaes(self, vae_dict):
        if not all(isinstance(key, str) for key in vae_dict.keys()):
            raise ValueError("Expected modality name as str, but got {}.".format(list(vae_dict.keys())))
        self.vaes.update(vae_dict)

    def make_output_dict(self, encoder_dist=None, decoder_dist=None, latent_samples=None, joint_dist=None,
                         enc_dist_private=None, dec_dist_private=None, joint_decoder_dist=None,
                         cross_decoder_dist=None):
        out = VAEOutput()
        vals = locals()
        for f in _FIELDS:
            out.set_with_dict(vals[f], f)
        return out

    def encode(self, inputs, raw_ok=False):
        """mmvae_base.py:139-159: {mod: {"shared": (mu, s), "private": (mu, s) | None}}.

        An encoder that declares ``returns_raw_logvar = True`` returns the RAW output of its second Linear head instead
        of applying the reference tail ``softmax(raw, -1) + 1e-6`` (encoders.py:49-54).  With raw_ok (the objectives)
        the raw tensor is passed on under "full" with "raw": True and the latent kernels evaluate the tail themselves
        (SURVEY 8f rank 1); otherwise the tail is applied here, so every other caller sees reference semantics."""
        qz_xs = {}
        for modality, vae in self.vaes.items():
            if modality in inputs and inputs[modality]["data"] is not None:
                mu, s = vae.enc(inputs[modality])
                raw = bool(getattr(vae.enc, "returns_raw_logvar", False))
                if raw and not raw_ok:
                    s, raw = self._enc_tail(s), False
                if raw:  # slices of raw logits mean nothing: only the full row goes to the kernels
                    qz_xs[modality] = {"shared": None, "private": None, "full": (mu, s), "raw": True}
                elif not self.latent_factorization:
                    qz_xs[modality] = {"shared": (mu, s), "private": None, "full": (mu, s), "raw": False}
                else:
                    n = vae.n_latents
                    qz_xs[modality] = {"shared": [mu[:, :n], s[:, :n]], "private": [mu[:, n:], s[:, n:]],
                                       "full": (mu, s), "raw": False}
            elif modality in inputs:
                qz_xs[modality] = {"shared": None, "private": None, "full": None, "raw": False}
        return qz_xs

    @staticmethod
    def _enc_tail(raw):
        """Reference encoders.py:52: F.softmax(logvar_layer(data), dim=-1) + Constants.eta."""
        return F.softmax(raw, dim=-1) + 1e-6

    def decode(self, samples):
        out = {}
        for modality, vae in self.vaes.items():
            if modality in samples and samples[modality]["latents"] is not None:
                out[modality] = vae.dec(samples[modality])
            elif modality in samples:
                out[modality] = None
        return out

    def get_missing_modalities(self, mods):
        missing = [m for m, v in mods.items() if v["data"] is None]
        present = [m for m, v in mods.items() if v["data"] is not None]
        return missing, present

    @staticmethod
    def product_of_experts(mu, logvar):
        """mmvae_base.py:203-222 on (E,B,D) stacks -> (pd_mu, pd_var); runs the fusion kernel."""
        E = mu.shape[0]
        res = ops.latent_draws(mu, logvar, None, None, None,
                               [ops.Draw(mods=tuple(range(E)), width=mu.shape[-1], want_params=True)])
        return res[0]["loc"], res[0]["scale"]

    @abc.abstractmethod
    def modality_mixing(self, mods):
        pass

    @abc.abstractmethod
    def objective(self, mods):
        pass

    def forward(self, inputs, K=1):
        raise NotImplementedError

    # ---- helpers shared by the plugins ---------------------------------------------------------------------
    def _noise(self, kind, shape, device):
        if self.noise_source is not None:
            return self.noise_source(kind, tuple(shape)).to(device=device, dtype=torch.float32)
        if kind == "laplace":  # torch laplace.py:75-79
            lo = torch.finfo(torch.float32).eps - 1
            return torch.empty(shape, device=device, dtype=torch.float32).uniform_(lo, 1)
        return torch.randn(shape, device=device, dtype=torch.float32)

    def _require_all(self, mods):
        missing = [m for m in self.vaes.keys() if m not in mods or mods[m]["data"] is None]
        if missing:
            raise ValueError("{}.objective needs every modality present (as the reference does); missing: {}".format(
                type(self).__name__, missing))

    def _stack(self, enc, names, part="full"):
        """(M,B,Dtot) fp32 stacks of the encoder outputs of `names` (padded to the widest modality)."""
        mus = [enc[n][part][0].float() for n in names]
        ss = [enc[n][part][1].float() for n in names]
        width = max(m.shape[-1] for m in mus)
        if any(m.shape[-1] != width for m in mus):
            mus = [F.pad(m, (0, width - m.shape[-1])) for m in mus]
            ss = [F.pad(s, (0, width - s.shape[-1]), value=1.0) for s in ss]
        return torch.stack(mus), torch.stack(ss)

    def _stack_raw(self, enc, names, part="full"):
        """(mu, s, s_raw): like _stack, for encoder outputs obtained with encode(raw_ok=True).  s_raw is True when `s`
        holds raw second-head logits for the kernels' fused encoder tail -- every encoder of `names` returned raw
        logits, all of the same width, and `part` is the whole encoder row.  Otherwise the tail is applied here."""
        whole = part == "full" or not self.latent_factorization
        raws = [bool(enc[n].get("raw")) for n in names]
        widths = {enc[n]["full"][0].shape[-1] for n in names}
        if all(raws) and whole and len(widths) == 1:
            return (torch.stack([enc[n]["full"][0].float() for n in names]),
                    torch.stack([enc[n]["full"][1].float() for n in names]), True)
        fixed = {}
        for n in names:
            e = enc[n]
            if e.get("raw"):
                mu, s = e["full"][0], self._enc_tail(e["full"][1])
                k = self.vaes[n].n_latents
                e = {"full": (mu, s), "shared": (mu, s) if not self.latent_factorization else [mu[:, :k], s[:, :k]],
                     "private": None if not self.latent_factorization else [mu[:, k:], s[:, k:]]}
            fixed[n] = e
        mu, s = self._stack(fixed, names, part)
        return mu, s, False

    def _batch_total(self, B):
        return self.global_batch if self.global_batch is not None else B
