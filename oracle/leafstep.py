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


def loss_latent_only(cfg, mods, pz, noise, rows, beta=1.0, zs_out=None):
    """MoE IWAE / DReG with the likelihood row vectors given directly (SURVEY 8d "latent + combine only"):
    objectives.py:342-387 on top of MOE.forward's samples."""
    M, K = len(mods), cfg["K"]
    zs = refmath.moe_forward(mods, noise, K)
    if zs_out is not None:
        zs_out.extend(zs)
    mu0 = torch.zeros_like(pz)
    _, s0 = refmath.prior_params(mu0, pz)
    L = len(rows) // M
    lws = []
    for r in range(M):
        lpz = refmath.normal_log_prob(zs[r], mu0, s0).sum(-1)  # (K,B)
        lq = refmath.log_mean_exp(torch.stack(
            [refmath.log_prob(mj["dist"], zs[r], mj["mu"], mj["s"]).sum(-1) for mj in mods]))
        lpx = sum(rows[r * L + l].reshape(K, -1) for l in range(L))
        if cfg["obj"] == "iwae":
            lws.append(lpz + lpx - beta * lq)
        else:
            lws.append(lpz.sum(-1) + lpx.sum(-1) - lq.sum(-1))
    if cfg["obj"] == "iwae":
        return -refmath.log_mean_exp(torch.cat(lws)).sum()
    lw = torch.stack(lws)
    with torch.no_grad():
        wt = (lw - torch.logsumexp(lw, 1, keepdim=True)).exp()
    return -(wt * lw).mean(0).sum()


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
    if cfg.get("latent_only"):
        zs = []
        l = loss_latent_only(cfg, mods, leaves["pz_logits"], noise, leaves["recon"], beta, zs)
        if tensors.get("dz") is not None:  # gradient the (absent) decoders would send into z: a second backward root
            dz = tensors["dz"].to(device=device, dtype=dtype)
            (l + sum((z * dz[r]).sum() for r, z in enumerate(zs))).backward()
        else:
            l.backward()
    else:
        l = loss(cfg, mods, leaves["pz_logits"], noise, beta)
        l.backward()
    grads = {"mu": leaves["mu"].grad, "s": leaves["s"].grad, "pz_logits": leaves["pz_logits"].grad}
    for i, r in enumerate(leaves["recon"]):
        grads["recon%d" % i] = r.grad
    return l.detach(), grads
