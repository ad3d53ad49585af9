"""Hot-path helpers with the reference's names (reference utils.py: log_mean_exp :395-396, kl_divergence :399-405,
subsample_input_modalities :86-112, find_out_batch_size :72-78, softclip :66-69, Constants :253-259, combinatorial
:595-601), written for this package."""
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
