"""TEST INFRASTRUCTURE / CPU BASELINE: the oracle restatement (oracle/refmath.py) evaluated on the leaf-tensor
protocol of SURVEY.md 8d -- the same tensors mmvae_b200.workloads.LeafStep feeds to the CUDA path.  Decoders are
replaced by leaf reconstruction tensors handed out in the order the reference evaluates its likelihood terms."""
import torch

from . import refmath


def build(cfg, tensors, device="cpu", dtype=torch.float32):
    """Leaves (requires_grad) + the modality spec list refmath's model restatements take."""
    mu = tensors["mu"].to(device=device, dtype=dtype).clone().requires_grad_(True)
    s = tensors["s"].to(device=device, dtype=dtype).clone().requires_grad_(True)
    pz = tensors["pz_logits"].to(device=device, dtype=dtype).clone().requires_grad_(True)
    recon = [r.to(device=device).float().to(dtype).clone().requires_grad_(True) for r in tensors["recon"]]
    it = iter(recon)
    dec = lambda z: next(it)
    mods = [dict(mu=mu[i], s=s[i], dist=m["dist"], ltype=m["ltype"], lam=m["lam"],
                 target=tensors["targets"][i].to(device=device, dtype=dtype), dec=dec)
            for i, m in enumerate(cfg["mods"])]
    noise = [n.to(device=device, dtype=dtype) for n in tensors["noise"]]
    return dict(mu=mu, s=s, pz_logits=pz, recon=recon), mods, noise


def loss(cfg, mods, pz, noise, beta=1.0):
    model = cfg["model"]
    if model == "moe":
        return refmath.moe_objective(mods, pz, noise, obj=cfg["obj"], beta=beta, K=cfg["K"])["loss"]
    if model == "poe":
        return refmath.poe_objective(mods, pz, noise, beta=beta)["loss"]
    if model == "mopoe":
        return refmath.mopoe_objective(mods, pz, noise, beta=beta, K=1)["loss"]
    if model == "dmvae":
        return refmath.dmvae_objective(mods, pz, noise, beta=beta, K=1)["loss"]
    raise ValueError(model)


def run(cfg, tensors, beta=1.0, device="cpu", dtype=torch.float32):
    """One objective fwd+bwd.  Returns (loss, {name: grad}) with grads for mu, s, pz_logits and recon[i]."""
    leaves, mods, noise = build(cfg, tensors, device, dtype)
    l = loss(cfg, mods, leaves["pz_logits"], noise, beta)
    l.backward()
    grads = {"mu": leaves["mu"].grad, "s": leaves["s"].grad, "pz_logits": leaves["pz_logits"].grad}
    for i, r in enumerate(leaves["recon"]):
        grads["recon%d" % i] = r.grad
    return l.detach(), grads
