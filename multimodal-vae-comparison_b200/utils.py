"""Hot-path helpers with the reference's names (reference utils.py: log_mean_exp :395-396, kl_divergence :399-405,
subsample_input_modalities :86-112, find_out_batch_size :72-78, softclip :66-69, Constants :253-259, combinatorial
:595-601) and the analysis consumers of the same math (make_kl_df :130-162, trainer.eval_forward / analyse_data
trainer.py:242-279), written for this package."""
import itertools
import math

import torch
import torch.distributions as dist
import torch.nn.functional as F

from . import ops


class Constants:
    eta = 1e-6
    eps = 1e-9
    log2 = math.log(2)
    log2pi = math.log(2 * math.pi)
    logceilc = 88
    logfloorc = -104


def log_mean_exp(value, dim=0, keepdim=False):
    """logsumexp(value, dim) - log(size) (reference utils.py:395-396).  The objectives never call this on the hot path
    (the IWAE / DReG kernels fuse it); it is kept for user code."""
    return torch.logsumexp(value, dim, keepdim=keepdim) - math.log(value.size(dim))


def softclip(tensor, min):
    return min + F.softplus(tensor - min)


def find_out_batch_size(inputs):
    for v in inputs.values():
        if v["data"] is not None:
            return v["data"].shape[0]
    return None


def combinatorial(lst):
    return [(a, b) for i, a in enumerate(lst) for b in lst[i + 1:]]


def kl_divergence(d1, d2, K=100):
    """Closed form KL for the pairs on the path through the element-wise KL kernel (Normal / Laplace posterior against
    a Normal prior on CUDA tensors); torch's registry otherwise; Monte-Carlo estimate as the reference's last resort."""
    if isinstance(d2, dist.Normal) and isinstance(d1, (dist.Normal, dist.Laplace)) and d1.loc.is_cuda \
            and d2.loc.numel() in (1, d1.loc.shape[-1]):
        return ops.kl_elementwise(d1.loc, d1.scale, d2.loc, d2.scale, isinstance(d1, dist.Laplace))
    if (type(d1), type(d2)) in torch.distributions.kl._KL_REGISTRY:
        return torch.distributions.kl_divergence(d1, d2)
    samples = d1.rsample(torch.Size([K]))
    return (d1.log_prob(samples) - d2.log_prob(samples)).mean(0)


def subsample_input_modalities(mods, forbidden=()):
    """All non-empty modality subsets of a batch dict as batch dicts whose absent modalities carry data = masks = None
    (reference utils.py:86-112).  Differences on purpose: the tensors are SHARED, not deep-copied (the reference
    deep-copies every image tensor 2^M - 1 times), and the order is itertools.combinations order instead of the
    PYTHONHASHSEED dependent set order."""
    names = list(mods.keys())
    out = []
    for n in range(1, len(names) + 1):
        for combo in itertools.combinations(names, n):
            if "+".join(combo) in forbidden:
                continue
            entry = {}
            for k in names:
                if k in combo:
                    entry[k] = mods[k]
                else:
                    entry[k] = {kk: (None if kk in ("data", "masks") else vv) for kk, vv in mods[k].items()}
            out.append(entry)
    return out


# ----------------------------------------------------------------------------------------------------------
# analysis / evaluation consumers
# ----------------------------------------------------------------------------------------------------------
KL_AX_NAMES = ["Dimensions", r"KL$(q\,||\,p)$"]


def _tensors_to_df(tensors, head, keys, ax_names):
    """Long-format table of reference visualization.py:106-122 (tensor_to_df / tensors_to_df): one row per
    (tensor, sample, dimension), columns [head, ax_names[0], ax_names[1]], sample-major inside a dimension."""
    import numpy as np
    import pandas as pd
    dfs = []
    for t in tensors:
        assert t.ndim == 2, "Can only currently convert 2D tensors to dataframes"
        df = pd.DataFrame(data=np.asarray(t), columns=np.arange(t.shape[1]))
        dfs.append(df.melt(value_vars=df.columns, var_name=ax_names[0], value_name=ax_names[1]))
    df = pd.concat(dfs, keys=keys)
    df.reset_index(level=0, inplace=True)
    df.rename(columns={"level_0": head}, inplace=True)
    return df


def kl_tables(qz_xs, pz):
    """(T, n, D) tensor of per-dimension divergences: KL(q_i || p) for every posterior, then the symmetric
    J(q_i, q_j) = 0.5 (KL(q_i||q_j) + KL(q_j||q_i)) of every pair i < j -- one kernel launch (ops.kl_table) for Normal
    or Laplace posteriors on the GPU; everything stays on the device."""
    if not isinstance(qz_xs, (list, tuple)):
        qz_xs = [qz_xs]
    fam = type(qz_xs[0])
    if fam not in (dist.Normal, dist.Laplace) or any(type(q) is not fam for q in qz_xs) or not isinstance(pz, dist.Normal):
        raise NotImplementedError("kl_tables: Normal or Laplace posteriors of one family against a Normal prior")
    D = qz_xs[0].loc.shape[-1]
    loc = torch.stack([q.loc.reshape(-1, D) for q in qz_xs])
    scale = torch.stack([q.scale.reshape(-1, D) for q in qz_xs])
    return ops.kl_table(loc, scale, pz.loc, pz.scale, laplace=fam is dist.Laplace)


def make_kl_df(qz_xs, pz):
    """Reference utils.py:130-162: the per-dimension KL table the analysis hook plots -- KL(q(z|x_i) || p(z)) for every
    modality and the symmetric J divergence of every pair, as the same long-format pandas DataFrame (same keys, column
    names and row order).  The divergences come from one kernel launch on the device; only the finished (T, n, D) table
    is copied to the host (the reference moves every distribution to the CPU first and calls kl_divergence M + 2 C(M,2)
    times there)."""
    if isinstance(qz_xs, (list, tuple)) and len(qz_xs) == 1:
        qz_xs = qz_xs[0]
    if isinstance(qz_xs, (list, tuple)):
        M = len(qz_xs)
        table = kl_tables(list(qz_xs), pz).cpu()
        keys = [r"KL$(q(z|x_{})\,||\,p(z))$".format(i) for i in range(M)] + \
               [r"J$(q(z|x_{})\,||\,q(z|x_{}))$".format(i, j) for i, j in itertools.combinations(range(M), 2)]
        return _tensors_to_df([table[i] for i in range(table.shape[0])], "KL", keys, KL_AX_NAMES)
    table = kl_tables([qz_xs], pz).cpu()
    return _tensors_to_df([table[0]], "KL", [r"KL$(q(z|x)\,||\,p(z))$"], KL_AX_NAMES)


def check_input_unpacked(mods):
    """reference utils.py:80-84."""
    if len(mods.keys()) == 1:
        mods = mods[list(mods.keys())[0]]
    return mods


def data_to_device(data, device):
    """reference utils.py:114-118 (returns a new dict instead of mutating the caller's)."""
    return {key: {k: v.to(device=device, non_blocking=True) if hasattr(v, "to") else v for k, v in entry.items()}
            for key, entry in data.items()}


def eval_forward(model, data):
    """reference trainer.py:274-279 (MultimodalVAE.eval_forward): forward pass outside training -> unpacked VAEOutput."""
    device = next(model.parameters()).device
    with torch.no_grad():
        output = model.forward(check_input_unpacked(data_to_device(data, device)))
    return output.unpack_values()


def analyse_data(model, data, num_samples=250):
    """The numerical part of reference trainer.py:242-272 (analyse_data): encode `data`, build the per-dimension KL table
    and collect the latent samples the T-SNE plot embeds (prior samples first, then one entry per modality).  Plotting
    (seaborn / matplotlib) is outside this path.  Returns {"kl_df", "latent_samples", "output"}."""
    out = eval_forward(model, data)
    pz = model.pz(*model.pz_params)
    with torch.no_grad():
        zss = [pz.sample(torch.Size([1, num_samples])).view(-1, pz.batch_shape[-1])] + \
              [zs["latents"].view(-1, zs["latents"].size(-1)) for zs in out["latent_samples"]]
        kl_df = make_kl_df([q for q in out["encoder_dist"] if q is not None], pz)
    return {"kl_df": kl_df, "latent_samples": zss, "output": out}
