"""CPU, world_size 2, gloo: the host-side logic of the batch-sharded path -- shard ranges, the flat-bucket gradient
all-reduce(SUM), and the three forward exchanges of SURVEY 8e (DReG batch sums, MoPoE global batch mean,
optimal_sigma global RMS) -- reproduce the single-process result of the oracle on the full batch."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(tensors, cfg, lo, hi, K):
    """Rows [lo,hi) of every per-sample tensor of a leaf set (reconstructions are k-major: row = k*B + b)."""
    B = cfg["B"]
    out = {"mu": tensors["mu"][:, lo:hi], "s": tensors["s"][:, lo:hi], "pz_logits": tensors["pz_logits"],
           "targets": [t[lo:hi] for t in tensors["targets"]]}
    out["recon"] = [r.view(r.shape[0] // B, B, *r.shape[1:])[:, lo:hi].reshape(-1, *r.shape[1:]) for r in tensors["recon"]]
    out["noise"] = [n[:, lo:hi] for n in tensors["noise"]]
    c = dict(cfg)
    c["B"] = hi - lo
    return c, out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mmvae_b200.parallel as par
        import mmvae_b200.workloads as W
        from oracle import leafstep, refmath
        res = {}
        # shard ranges tile the batch
        assert [par.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
        # ---- additive objectives: sum of shard losses == full loss; replicated grads all-reduce(SUM) --------
        for name in ("c1_poe_elbo_cdsprites_l1", "c2_moe_iwae_cdsprites_l5", "c5_dmvae_elbo_cub"):
            cfg, t = W.make_leaves(name, B=6, seed=5)
            t["pz_logits"] = torch.randn(1, cfg["D"], generator=torch.Generator().manual_seed(1)) * 0.3
            full_loss, full_g = leafstep.run(cfg, t, dtype=torch.float64)
            lo, hi = par.shard_range(6, rank, world)
            c2, t2 = _shard(t, cfg, lo, hi, cfg["K"])
            leaves, mods, noise = leafstep.build(c2, t2, dtype=torch.float64)
            loss = leafstep.loss(c2, mods, leaves["pz_logits"], noise)
            loss.backward()
            pz = torch.nn.Parameter(leaves["pz_logits"].detach().clone())
            pz.grad = leaves["pz_logits"].grad.clone()
            par.GradSync([pz])()
            tot = loss.detach().clone()
            dist.all_reduce(tot)
            res[name] = (float((tot - full_loss).abs() / full_loss.abs()),
                         float((pz.grad - full_g["pz_logits"]).abs().max() / full_g["pz_logits"].abs().max().clamp_min(1e-30)),
                         float((leaves["mu"].grad - full_g["mu"][:, lo:hi]).abs().max()))
        # ---- DReG parity mode: (M,K) batch sums all-reduced between the combine stages -----------------------
        cfg, t = W.make_leaves("c4_moe_dreg_mnistsvhn", B=6, seed=6)
        cfg["K"] = 5
        cfg, t = W.make_leaves("c4_moe_dreg_mnistsvhn", B=6, seed=6)
        full_loss, _ = leafstep.run(cfg, t, dtype=torch.float64)
        lo, hi = par.shard_range(6, rank, world)
        c2, t2 = _shard(t, cfg, lo, hi, cfg["K"])
        leaves, mods, noise = leafstep.build(c2, t2, dtype=torch.float64)
        lw_local = refmath.moe_objective(mods, leaves["pz_logits"], noise, obj="dreg", K=cfg["K"])["lw"].detach()
        dist.all_reduce(lw_local)  # stage-1 partial sums -> global
        wt = (lw_local - torch.logsumexp(lw_local, 1, keepdim=True)).exp()
        res["dreg"] = float((-(wt * lw_local).mean(0).sum() - full_loss).abs() / full_loss.abs())
        # ---- MoPoE: batch means are global -> local sums scaled by 1/B_global -------------------------------
        cfg, t = W.make_leaves("c3_mopoe_elbo_sprites", B=6, seed=7)
        full_loss, _ = leafstep.run(cfg, t, dtype=torch.float64)
        c2, t2 = _shard(t, cfg, lo, hi, 1)
        leaves, mods, noise = leafstep.build(c2, t2, dtype=torch.float64)
        part = leafstep.loss(c2, mods, leaves["pz_logits"], noise).detach() * (hi - lo) / 6.0
        dist.all_reduce(part)
        res["mopoe"] = float((part - full_loss).abs() / full_loss.abs())
        # ---- optimal_sigma: one scalar (sum of squares) all-reduced ----------------------------------------
        g = torch.Generator().manual_seed(9)
        x, tt = torch.randn(6, 40, generator=g, dtype=torch.float64), torch.rand(6, 40, generator=g, dtype=torch.float64)
        full = refmath.lpx_rows("optimal_sigma", x, tt, 1.0, 1)
        ss = ((tt[lo:hi] - x[lo:hi]) ** 2).sum().reshape(1)
        dist.all_reduce(ss)
        log_sigma = refmath.softclip((ss / x.numel()).sqrt().log(), -6)
        rows = -(((tt[lo:hi] - x[lo:hi]) / log_sigma.exp()) ** 2 + log_sigma + 0.5 * math.log(2 * math.pi)).sum(-1)
        res["osigma"] = float((rows - full[lo:hi]).abs().max() / full.abs().max())
        # ---- armed GradSync: hooks fire the flat-bucket all-reduce once the LAST gradient exists; wait() runs it when
        # a parameter got no gradient in this backward (every rank still issues exactly one collective per step)
        a, b, c = (torch.nn.Parameter(torch.full((3,), float(i + 1))) for i in range(3))
        sync = par.GradSync([a, b, c]).arm()
        ((a * (rank + 1)).sum() + (b * 2.0).sum() + (c * 0.5 * (rank + 1)).sum()).backward()
        fired_in_hook = sync._inflight
        sync.wait()
        ok = fired_in_hook and torch.allclose(a.grad, torch.full((3,), 3.0)) and torch.allclose(b.grad, torch.full((3,), 4.0)) \
            and torch.allclose(c.grad, torch.full((3,), 1.5))
        for p_ in (a, b, c):
            p_.grad = None
        ((a * (rank + 1)).sum() + (b * 2.0).sum()).backward()  # c unused: the bucket is completed in wait()
        late = not sync._inflight
        sync.wait()
        ok = ok and late and torch.allclose(a.grad, torch.full((3,), 3.0)) and torch.allclose(c.grad, torch.zeros(3))
        sync.disarm()
        res["gradsync_armed"] = 0.0 if ok else 1.0
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_equals_full_batch_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
        assert p.exitcode == 0
    res = q.get()
    for k, v in res.items():
        if isinstance(v, tuple):  # (loss, all-reduced replicated grad [fp32 flat bucket], local grads)
            assert v[0] < 1e-9 and v[1] < 1e-6 and v[2] < 1e-9, (k, v)
        else:
            assert v < 1e-9, (k, v)


def test_shard_batch_slices_reference_batch_dict():
    import mmvae_b200.parallel as par
    batch = {"mod_1": {"data": torch.arange(10).view(10, 1), "masks": None, "categorical": False},
             "mod_2": {"data": torch.arange(20).view(10, 2), "masks": torch.ones(10, 2, dtype=torch.bool), "categorical": True}}
    parts = [par.shard_batch(batch, r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p["mod_1"]["data"] for p in parts]), batch["mod_1"]["data"])
    assert [p["mod_2"]["masks"].shape[0] for p in parts] == [3, 3, 2, 2]
    assert parts[0]["mod_2"]["categorical"] is True
