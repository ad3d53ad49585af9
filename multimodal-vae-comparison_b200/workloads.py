"""Functional objective steps on leaf tensors -- the measurement protocol of SURVEY.md 8d.

"Objective fwd+bwd" = fusion / sampling / log-densities / KL (segment A), the decoder log-likelihood row reductions
and the ELBO / IWAE / DReG combination (segment B), and the backward of both, with the dense decoders replaced by
*leaf* reconstruction tensors so that no dense layer is timed.  Inputs: ``mu_m, s_m`` (B, Dtot), prior logits (1, D),
``recon_{t<-s}`` (K*B, *data_dim) (all requiring grad), targets (B, *data_dim), pre-generated noise.  Outputs: the
loss and the gradient w.r.t. every leaf.

Everything here is product code (it only calls ops.py -> the C ABI); the same leaf tensors are fed to the oracle by
tests/ and by bench.py's cpu_baseline leg.
"""
import math

import torch

from . import ops
from . import synthetic as syn
from .mmvae_models import MOE, mopoe_inmodel_component, mopoe_subsets, poe_subsets
from .ops import Draw


def term_plan(model, M):
    """(target modality, source tag) of every likelihood term, in evaluation order."""
    if model == "moe":
        plan = []
        for r in range(M):
            plan.append((r, "self"))
            if MOE._cross_source(M, r) is not None:
                plan.append((r, "cross"))
        return plan
    if model == "poe":
        return [(m, "subset%d" % a) for a in range(2 ** M - 1) for m in range(M)]
    if model == "mopoe":
        return [(m, "joint") for m in range(M)]
    if model == "dmvae":
        plan = []
        for m in range(M):
            plan += [(m, "shared"), (m, "joint")] + [(m, "cross%d" % j) for j in range(M) if j != m]
        return plan
    raise ValueError(model)


def _noise_plan(model, M, K, B, D, pv, dists):
    """Shapes of the noise tensors in the reference's rsample order."""
    if model == "poe":
        return [("normal", (1, B, D)) for _ in range(2 ** M - 1)]
    if model == "moe":
        return [(dists[m], (K, B, D)) for m in range(M)]
    if model == "mopoe":
        return [("normal", (1, B, D)) for _ in range(M)]
    plan = [("normal", (1, B, D))]
    for m in range(M):
        plan += [("normal", (1, B, D)), ("normal", (1, B, pv))] + [("normal", (1, B, D))] * (M - 1)
    return plan


def make_leaves(name, B=None, seed=1234, recon_dtype=torch.float32):
    """Synthetic leaf tensors of workload `name` (synthetic.WORKLOADS), generated on the host with a fixed seed.
    Returns (cfg, dict of CPU tensors): mu, s (M,B,Dtot), pz_logits (1,D), targets [..], recon [..], noise [..]."""
    cfg = dict(syn.WORKLOADS[name])
    B = B or cfg["B"]
    cfg["B"] = B
    g = syn.gen(seed)
    M, K, D, pv = len(cfg["mods"]), cfg["K"], cfg["D"], cfg.get("private") or 0
    post = [syn.make_posterior(g, B, D + pv) for _ in range(M)]
    t = {"mu": torch.stack([p[0] for p in post]), "s": torch.stack([p[1] for p in post]),
         "pz_logits": torch.zeros(1, D)}
    # bf16 configurations carry bf16 inputs AND outputs (T counted at 2 bytes): targets are rounded to
    # bf16 once, here; the oracle sees the same rounded values
    t["targets"] = [syn.make_target(g, m["target"], B, m["data_dim"]).to(recon_dtype) for m in cfg["mods"]]
    Kr = K if cfg["model"] == "moe" else 1
    if cfg.get("latent_only"):
        # synthetic likelihood ROWS (one value per decoder row) instead of reconstructions
        t["recon"] = [-(torch.rand(Kr * B, generator=g) * 50 + 500) for _ in term_plan(cfg["model"], M)]
    else:
        t["recon"] = [syn.make_recon(g, cfg["mods"][tm]["ltype"], Kr * B, cfg["mods"][tm]["data_dim"]).to(recon_dtype)
                      for tm, _ in term_plan(cfg["model"], M)]
    t["noise"] = [syn.make_noise(g, kind, shape)
                  for kind, shape in _noise_plan(cfg["model"], M, K, B, D, pv, [m["dist"] for m in cfg["mods"]])]
    if cfg.get("latent_only"):
        # the gradient the (absent) decoders would send back into z: a second root of the backward, so that the latent
        # backward reads eps AND dz like in a full step (SURVEY 8d counts 4*K*D*e bytes per drawn tensor)
        t["dz"] = torch.randn(M, K, B, D, generator=g) * 0.1
    return cfg, t


def algorithmic_bytes(cfg, recon_dtype=torch.float32):
    """Minimum HBM traffic of one objective fwd+bwd per SAMPLE: per likelihood term 2R+T when the
    row weights are known a priori (ELBO), 3R+2T otherwise (IWAE / DReG); latent segment 4*K*D*4 per drawn tensor
    plus the (M, Dtot) parameters read and their gradients written."""
    e = 2 if recon_dtype == torch.bfloat16 else 4
    M, K, D, pv = len(cfg["mods"]), cfg["K"], cfg["D"], cfg.get("private") or 0
    Kr = K if cfg["model"] == "moe" else 1
    tot = 0
    for tm, _ in term_plan(cfg["model"], M):
        if cfg.get("latent_only"):
            tot += 2 * Kr * 4  # the row value read, its gradient written
            continue
        P = int(math.prod(cfg["mods"][tm]["data_dim"]))
        R, T = Kr * P * e, P * e
        tot += (2 * R + T) if cfg["obj"] == "elbo" else (3 * R + 2 * T)
    n_noise = sum(int(math.prod(shape[2:])) * shape[0] for _, shape in
                  _noise_plan(cfg["model"], M, K, 1, D, pv, [m["dist"] for m in cfg["mods"]]))
    tot += 4 * n_noise * 4 + 2 * 2 * M * (D + pv) * 4
    return tot


class LeafStep:
    """One objective fwd+bwd on device-resident leaves through the CUDA path.  ``run()`` returns the loss tensor;
    gradients land in ``.grad`` of ``self.mu / self.s / self.pz_logits / self.recon[i]``."""

    def __init__(self, cfg, tensors, device="cuda", beta=1.0, group=None, global_batch=None, sync_grads=False, fold=False):
        self.cfg, self.beta, self.group = cfg, beta, group
        self.model, self.obj = cfg["model"], cfg["obj"]
        self.M, self.K, self.D, self.pv, self.B = len(cfg["mods"]), cfg["K"], cfg["D"], cfg.get("private") or 0, cfg["B"]
        self.Bt = global_batch or self.B
        dev = torch.device(device)
        self.mu = tensors["mu"].to(dev).requires_grad_(True)
        self.s = tensors["s"].to(dev).requires_grad_(True)
        self.pz_logits = tensors["pz_logits"].to(dev).requires_grad_(True)
        self.targets = [x.to(dev) for x in tensors["targets"]]
        self.recon = [x.to(dev).requires_grad_(True) for x in tensors["recon"]]
        self.noise = [x.to(dev) for x in tensors["noise"]]
        self.dz = tensors["dz"].to(dev) if tensors.get("dz") is not None else None
        self._z = None
        self.plan = term_plan(self.model, self.M)
        # fold: the likelihood terms of one target modality as ONE (terms * B, ...) leaf -- what a plugin gets from a
        # decoder that folds K when it stacks the latents of those terms (mmvae_models._fold_ok): one launch per modality,
        # same values (row r reads target row r % B).  ELBO models with constant row weights only.
        self.fold = bool(fold) and self.obj == "elbo" and self.model in ("poe", "dmvae", "mopoe") and \
            not cfg.get("latent_only") and all(m["ltype"] != "optimal_sigma" for m in cfg["mods"])
        if self.fold:
            by_mod = {}
            for i, (tm, _) in enumerate(self.plan):
                by_mod.setdefault(tm, []).append(i)
            self.recon = [torch.cat([tensors["recon"][i] for i in idx], 0).to(dev).requires_grad_(True)
                          for tm, idx in sorted(by_mod.items())]
            self.fold_index = {tm: idx for tm, idx in by_mod.items()}  # for tests: which terms a folded leaf holds
            self.plan = [(tm, "folded") for tm in sorted(by_mod)]
        self.codes = [1 if m["dist"] == "laplace" else 0 for m in cfg["mods"]]
        if self.model == "mopoe":
            subs = mopoe_subsets(range(self.M))
            comp = mopoe_inmodel_component(len(subs))
            bits = sum(1 << i for i in subs[comp]) | ((1 << 31) if len(subs[comp]) == self.M else 0)
            bits = bits - (1 << 32) if bits >= (1 << 31) else bits
            self.row_masks = torch.full((self.B,), bits, dtype=torch.int32, device=dev)
        if self.model != "moe":
            self.eps_packed = torch.cat([n.reshape(-1) for n in self.noise])
        else:
            self.eps_stacked = torch.stack(self.noise)  # (M,K,B,D), what mmvae_moe_logdens_* consume
        self._mu0 = torch.zeros(1, self.D, device=dev)
        self._ticket = ops.new_ticket(dev) if dev.type == "cuda" else None
        self._one = ops.mark_unit_grad(torch.ones((), device=dev)) if dev.type == "cuda" else torch.ones(())
        # Stream plan (measured r1 on C2, CUDA-graph replay, ms/step): single stream 0.447; latent kernels on a second
        # stream 0.439; likelihood terms alternating between two streams 0.430; both (the default, `streams = 3`)
        # 0.416.  Two streaming kernels resident at a time cover each other's ramp-up and tail, and the latency-bound
        # latent kernels fill SM slots beside them.  Rejected: long-row terms on one stream and short-row terms on the
        # other (0.506: the category_ce CTAs take SM slots from the streaming kernels for their whole lifetime).
        # Forks and joins are stream waits inside the captured graph; autograd replays every backward node on the
        # stream of its forward, so the backward overlaps the same way.
        self.streams = 3 if dev.type == "cuda" else 1
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None   # every other likelihood term
        self.side2 = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None  # latent kernels
        self._forked = []
        # sync_grads: all-reduce(SUM) the gradient of the replicated prior logits INSIDE the step, launched from a
        # post-accumulate-grad hook on a side stream (parallel.GradSync.arm) and joined at the end of run().  The
        # IWAE / DReG branch creates the latent nodes after the likelihood nodes, so autograd (highest sequence number
        # first) runs the latent backward -- which produces that gradient -- before the long likelihood backward
        # kernels, and the collective hides behind them.
        # With a parallel.PeerGroup as `group` the same all-reduce is FUSED into the backward kernel of the prior scale
        # (peer memory over NVLink, csrc/peer.cuh): no NCCL call, no side stream, nothing to join.
        self.sync = None
        self.peer = group if (sync_grads and hasattr(group, "bufs_dev")) else None
        if sync_grads and group is not None and self.peer is None:
            from .parallel import GradSync
            self.sync = GradSync([self.pz_logits], group).arm()

    def finish(self):
        if self.sync is not None:
            self.sync.wait()

    # -- stream plan ------------------------------------------------------------------------------------------
    def _terms(self, fn):
        """[fn(i) for every likelihood term]; with 3 streams the odd terms are issued on the second stream."""
        n = len(self.plan)
        if self.streams == 1 or n < 2:
            return [fn(i) for i in range(n)]
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        out = [None] * n
        with torch.cuda.stream(self.side):
            for i in range(1, n, 2):
                out[i] = fn(i)
        for i in range(0, n, 2):
            out[i] = fn(i)
        self._forked.append(self.side)
        return out

    def _latent(self, fn):
        """fn() (the latent kernels) on the third stream."""
        if self.streams == 1:
            return fn()
        self.side2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side2):
            out = fn()
        self._forked.append(self.side2)
        return out

    def _join(self, *tensors):
        """The current stream waits for the forked streams; `tensors` were produced there and are consumed here."""
        cur = torch.cuda.current_stream()
        for st in self._forked:
            cur.wait_stream(st)
        if self._forked:
            for t in tensors:
                if torch.is_tensor(t):
                    t.record_stream(cur)
                    d = getattr(t, "_mmvae_deferred", None)
                    if d is not None:  # the row vector ops.elbo_combine reads in place of the placeholder sum
                        d[0].record_stream(cur)
        self._forked = []

    def leaves(self):
        return [self.mu, self.s, self.pz_logits] + self.recon

    def zero_grad(self):
        for t in self.leaves():
            t.grad = None

    # -- likelihood helpers ---------------------------------------------------------------------------------
    def _family(self, tm, tag):
        # MOE wraps its self reconstruction in a Normal whatever the family (reference mmvae_models.py:105-107)
        return "normal" if (self.model == "moe" and tag == "self") else self.cfg["mods"][tm]["dist"]

    def _rows(self, i):
        if self.cfg.get("latent_only"):
            return self.recon[i]  # already a (K*B,) row vector leaf
        tm, tag = self.plan[i]
        m = self.cfg["mods"][tm]
        fam = self._family(tm, tag)
        if m["ltype"] == "category_ce":
            return ops.catce_rows(self.recon[i], self.targets[tm], m["lam"])
        if m["ltype"] == "optimal_sigma":
            return ops.osigma_rows(self.recon[i], self.targets[tm], m["lam"], self.group, i)
        return ops.loglik_rows(self.recon[i], self.targets[tm], m["ltype"], fam, m["lam"])

    def _wsum(self, i, w_const=1.0, w_rows=None, defer=False):
        """defer: leave the batch sum of the term to ops.elbo_combine (one launch for the whole loss)."""
        tm, tag = self.plan[i]
        m = self.cfg["mods"][tm]
        fam = self._family(tm, tag)
        if m["ltype"] == "category_ce":
            return ops.catce_weighted_sum(self.recon[i], self.targets[tm], m["lam"], w_rows=w_rows, w_const=w_const,
                                          defer=defer)
        if m["ltype"] == "optimal_sigma":
            rows = ops.osigma_rows(self.recon[i], self.targets[tm], m["lam"], self.group, i)
            return (torch.dot(rows, w_rows) if w_rows is not None else w_const * rows.sum()), rows.detach()
        return ops.loglik_weighted_sum(self.recon[i], self.targets[tm], m["ltype"], fam, m["lam"], w_rows=w_rows,
                                       w_const=w_const, defer=defer)

    def _osigma(self, i):
        return self.cfg["mods"][self.plan[i][0]]["ltype"] == "optimal_sigma"

    def _prior(self):
        return self._mu0, ops.prior_scale(self.pz_logits, self.peer)

    # -- objectives -----------------------------------------------------------------------------------------
    def loss(self):
        return getattr(self, "_" + self.model)()

    def backward(self, loss):
        """Backward from the static unit root (no ones_like fill kernel at its head).  Latent-only workloads have a
        second root: z with the synthetic decoder gradient `dz`."""
        if self.dz is not None and self._z is not None:
            torch.autograd.backward([loss, self._z], [self._one, self.dz])
        else:
            loss.backward(self._one)
        self._z = None

    def run(self):
        self.zero_grad()
        loss = self.loss()
        self.backward(loss)
        self.finish()
        return loss

    def _moe(self):
        M, K, B = self.M, self.K, self.B
        eps = self.eps_stacked
        if self.obj == "elbo":
            mu0, s0 = self._prior()
            z, lq, _ = ops.moe_logdens(self.mu, self.s, mu0.detach(), s0.detach(), eps, self.codes, False)
            kls = ops.latent_draws(self.mu, self.s, None, None, None,
                                   [Draw(mods=(m,), direct=True, laplace=bool(self.codes[m]), kl_mode=2, width=self.D)
                                    for m in range(M)])
            total, n_keep, i = 0.0, 0.0, 0
            for r in range(M):
                S, _ = self._wsum(i, w_const=-1.0 / M)
                i += 1
                total, n_keep = total + S, n_keep + (S != 0).float()
                src = MOE._cross_source(M, r)
                if src is None:
                    continue
                lwt = lq[src, r, 0] - lq[src, src, 0].detach()
                S, _ = self._wsum(i, w_rows=-lwt.exp() / M)
                i += 1
                total, n_keep = total + S, n_keep + (S != 0).float()
            kld = torch.stack([k["kl"] for k in kls])
            return total + (self.beta / M) * n_keep * kld.sum()
        # likelihood rows first, latent nodes last: the backward then starts with the latent kernels (see __init__)
        rows = self._terms(self._rows)

        def latent():
            mu0, s0 = self._prior()
            return ops.moe_logdens(self.mu, self.s, mu0, s0, eps, self.codes, True)

        z, lq, lpz = self._latent(latent)
        self._join(lq, lpz, z, *rows)
        self._z = z
        L = len(rows) // M
        if self.obj == "iwae":
            return ops.iwae_combine_rows(lpz, lq, rows, L, self.beta, self._ticket)[0]
        return ops.dreg_combine_rows(lpz, lq, rows, L, self.group)[0]

    def _poe(self):
        M, D = self.M, self.D
        subsets = poe_subsets(range(M))

        def latent():
            mu0, s0 = self._prior()
            return ops.latent_draws(self.mu, self.s, mu0, s0, self.eps_packed,
                                    [Draw(mods=sub, prior=True, kl_mode=1, width=D, K=1) for sub in subsets]).kl_packed

        kl = self._latent(latent)
        sums = self._terms(lambda i: self._wsum(i, w_const=-1.0, defer=True)[0])
        self._join(kl, *sums)
        # loss = -sum_A sum_m sum_b lpx + beta sum_A sum_b KL_A: one launch (mmvae_objective_elbo)
        return ops.elbo_combine(sums, kl, [self.beta] * len(subsets))[0]

    def _mopoe(self):
        M, D = self.M, self.D
        draws = [Draw(rowmask=True, kl_mode=1 if i == 0 else 0, width=D, K=1) for i in range(M)]
        draws += [Draw(mods=(i,), direct=True, kl_mode=1, width=D) for i in range(M)]

        def latent():
            mu0, s0 = self._prior()
            return ops.latent_draws(self.mu, self.s, mu0, s0, self.eps_packed, draws, self.row_masks).kl_packed

        kl = self._latent(latent)  # packed rows: joint (draw 0), then the M unimodal posteriors
        sums = self._terms(lambda i: self._wsum(i, w_const=-1.0 / self.Bt, defer=not self._osigma(i))[0])
        self._join(kl, *sums)
        return ops.elbo_combine(sums, kl, [self.beta / ((M + 1) * self.Bt)] * (M + 1))[0]

    def _dmvae(self):
        M, D, pv = self.M, self.D, self.pv
        draws = [Draw(mods=tuple(range(M)), kl_mode=1, width=D, K=1)]
        idx = []
        for i in range(M):
            idx.append(len(draws))
            draws.append(Draw(mods=(i,), direct=True, kl_mode=1, col0=0, width=D, K=1))
            draws.append(Draw(mods=(i,), direct=True, kl_mode=2, col0=D, width=pv, K=1))
            draws += [Draw(mods=(j,), direct=True, col0=0, width=D, K=1) for j in range(M) if j != i]

        def latent():
            mu0, s0 = self._prior()
            return ops.latent_draws(self.mu, self.s, mu0, s0, self.eps_packed, draws).kl_packed

        kl = self._latent(latent)
        sums = self._terms(lambda i: self._wsum(i, w_const=-1.0, defer=True)[0])
        self._join(kl, *sums)
        # packed KL rows: joint, then (shared_i, private_i) per modality; the joint KL enters once per modality, the
        # private KL once per cross term (reference mmvae_models.py:455-463)
        return ops.elbo_combine(sums, kl, [M * self.beta] + [self.beta, (M - 1) * self.beta] * M)[0]


class GraphedStep:
    """CUDA-graph capture of a LeafStep (fwd+bwd): the launch-bound chain of ~30 small kernels replays as one graph
    launch.  Gradients are written into static buffers (``step.grads``)."""

    def __init__(self, step: LeafStep, warmup=3):
        self.step = step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step.run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        step.zero_grad()
        with torch.cuda.graph(self.graph):
            self.loss = step.loss()
            step.backward(self.loss)
            step.finish()  # joins the side stream of an in-step gradient all-reduce into the capture
        self.grads = [t.grad for t in step.leaves()]

    def run(self):
        self.graph.replay()
        return self.loss

    def close(self):
        """Destroy the captured graph.  Required before ``dist.destroy_process_group()`` when the step carries an
        in-graph NCCL all-reduce: a communicator waits for every graph that captured it to be destroyed (measured r1:
        destroy_process_group() blocks forever otherwise)."""
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
