#!/usr/bin/env python
"""bench.py -- objective fwd+bwd samples/s of the latent + objective hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--batch B_PER_GPU | --global-batch B] [--dtype fp32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload at N=1: BASELINE.json configs[1] -- MMVAE (MoE) IWAE K=30 on CdSprites+ level-5 shapes, batch 256, fp32 (the
configuration the metric is quoted on; it fits one GPU).  N>1 is batch sharded: weak scaling by default (the workload's
batch per GPU), `--global-batch B` splits a fixed batch over the ranks (strong scaling, e.g. configs[1] "batch 256,
8xB200" = 32 per GPU).  Collectives: one NCCL all-reduce(SUM) of the replicated-parameter gradient per step plus the
forward exchanges exact parity needs (DReG (M,K) batch sums, optimal_sigma scalar) -- all captured in the step's CUDA
graph; no data-path collective.

A "step" is one objective fwd+bwd on synthetic leaf tensors (SURVEY.md 8d protocol).  `value`: inputs resident in HBM.
`e2e`: the call a user of the reference makes -- model.objective(batch) + backward through the drop-in plugin (stand-in
linear encoders / decoders), the batch coming from pinned HOST memory every step, loss read back, gradients of the
replicated parameters all-reduced at N>1.  `e2e_leaf`: the leaf-tensor step with EVERY input (reconstructions included)
shipped from pinned host memory.  Timed region: at least K steps and at least 0.3 s of back-to-back steps (the count is in
`timed_steps`).  Working sets smaller than 4x the 126 MB L2 are timed step by step with an L2 flush in between.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "c2_moe_iwae_cdsprites_l5"
METRIC = "objective fwd+bwd samples/s"
MIN_TIMED_S = 0.3
L2_BYTES = 126 << 20
# CPU arms: a step of the full workload batch unless its element count exceeds this budget (C2 at B=256: 0.64 G)
CPU_ELEMS_PER_STEP = 700e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the workload's batch)")
    ap.add_argument("--global-batch", type=int, default=None, help="fixed global batch split over the ranks (strong scaling)")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--streams", type=int, default=3, choices=[1, 3],
                    help="3: likelihood terms alternate between two streams + latent kernels on a third (default); 1: one stream")
    ap.add_argument("--eager-sync", action="store_true",
                    help="N>1: all-reduce after the step (eager NCCL call) instead of inside it (side stream, captured)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the sharded-vs-unsharded numerical check")
    ap.add_argument("--no-roofline-timer", action="store_true",
                    help="skip the per-kernel CUDA-event leg (profiling runs under ncu: fewer launches to wade through)")
    ap.add_argument("--fold-terms", action="store_true",
                    help="ELBO workloads: the likelihood terms of a modality as ONE (terms*B, ...) leaf and launch -- what "
                         "the plugins do for decoders that fold K (same bytes, same values, fewer and larger launches)")
    ap.add_argument("--sweep", default=None,
                    help="several configurations in one process group, one JSON line each: workload:global_batch[:w],... "
                         "(w: the number is the batch per GPU, weak scaling)")
    ap.add_argument("--nccl-only", action="store_true",
                    help="N>1: every collective through NCCL (default: the small ones fused into our kernels over peer memory)")
    ap.add_argument("--cpu-budget-s", type=float, default=10.0, help="wall-clock budget of the cpu_baseline leg (seconds)")
    ap.add_argument("--cpu-batch", type=int, default=None,
                    help="opt-in: samples per step of the CPU arms (default: the workload batch, bounded by an element budget)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "how": "nvidia-smi -lms 20 over the timed region (>= 0.3 s of back-to-back steps)"}


class KernelTimer:
    """CUDA events around selected C-ABI launches, on the launching (current) stream."""

    def __init__(self, names):
        import torch
        self.torch, self.names, self.ev = torch, set(names), {n: [] for n in names}

    def wrap(self, name, fn, args=()):
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        # Eager steps are CPU bound around the small kernels: if the stream is idle when `a` is recorded, the host's
        # launch latency (~5-10 us of Python + ctypes) lands between the two events.  A ~10 us device-side spin keeps
        # the stream busy while record / launch / record are enqueued, so the events bracket the kernel alone.
        try:
            self.torch.cuda._sleep(20000)
        except Exception:
            pass
        a.record()
        rc = fn()
        b.record()
        nb = None
        if name.startswith("mmvae_moe_logdens") and len(args) > 10:
            # (mu, s, M, B, D, K, dists, mu0, s0, eps, ...): fwd reads eps + writes z; bwd reads eps (+ dz_ext)
            M, B, D, K = args[2], args[3], args[4], args[5]
            big, rows = M * K * B * D * 4, (M * M + M) * K * B * 4
            if name.endswith("_fwd"):
                nb = 2 * big + rows + 2 * M * B * D * 4
            else:
                has_dz = bool(getattr(args[10], "value", args[10]))
                nb = big * (2 if has_dz else 1) + rows + 4 * M * B * D * 4
        if name.startswith("mmvae_loglik_rowreduce") and len(args) > 8:
            # (recon, ld, dtype, target, ld, dtype, rows, B, P, ...) -> algorithmic bytes
            ex, et, rows, B, P = (2 if args[2] else 4), (2 if args[5] else 4), args[6], args[7], args[8]
            R, T = rows * P * ex, B * P * et
            nb = (R + T if name.endswith("_fwd") else 2 * R + T) + rows * 4
        if name.startswith("mmvae_osigma") and len(args) > 8:
            ex, et, rows, B, P = (2 if args[2] else 4), (2 if args[5] else 4), args[6], args[7], args[8]
            R, T = rows * P * ex, B * P * et
            nb = (2 * R + T) if name.endswith("_bwd") else (R + T)
        self.ev[name].append((a, b, nb))
        return rc

    def biggest(self, name):
        """(times_ms, bytes) of the launches with the largest algorithmic byte count (the dominant term)."""
        if not self.ev.get(name):
            return [], 0
        top = max(nb or 0 for _, _, nb in self.ev[name])
        ms = [a.elapsed_time(b) for a, b, nb in self.ev[name] if (nb or 0) == top]
        return (ms[len(ms) // 5:] if len(ms) >= 5 else ms), top


def local_batch(args, cfg_default_B, world):
    """(samples per GPU, scaling mode)."""
    if args.global_batch is not None:
        if args.global_batch % world:
            raise SystemExit("--global-batch %d is not divisible by %d ranks" % (args.global_batch, world))
        return args.global_batch // world, "strong"
    return (args.batch or cfg_default_B), "weak"


def flush_needed(step_bytes):
    """Working set of a step vs the 126 MB L2: big steps stream far more than L2 holds (no flush needed); small ones are
    timed one by one with a 256 MB write in between so that no step starts with its inputs cached."""
    return step_bytes < 4 * L2_BYTES


def config_dict(args, cfg, B, world, scaling):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm of the same invocation
    (how the GPU arm ran -- launch mode, streams, gradient sync -- is in its `run` object)."""
    import mmvae_b200.workloads as W
    import torch
    step_bytes = W.algorithmic_bytes(dict(cfg, B=B), torch.bfloat16 if args.dtype == "bf16" else torch.float32) * B
    return {"workload": args.workload, "model": cfg["model"], "objective": cfg["obj"], "K": cfg["K"],
            "latent_dim": cfg["D"], "batch_per_gpu": B, "global_batch": B * world,
            "mods": [{"data_dim": list(m["data_dim"]), "ltype": m["ltype"]} for m in cfg["mods"]],
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "scaling": scaling,
            "parallelism": "batch-sharded x%d, NCCL all-reduce of replicated grads" % world,
            "l2": ("per-step working set %.0f MB: 256 MB L2 flush between steps, steps timed one by one" % (step_bytes / 1e6))
                  if flush_needed(step_bytes) else
                  ("per-step working set %.2f GB >> 126 MB L2, no flush" % (step_bytes / 1e9))}


def cpu_sample_batch(args, cfg, B):
    """Samples per step of the CPU arms: the workload's own batch (same configuration as the GPU arm) unless one step would
    exceed the element budget; --cpu-batch overrides (explicit opt-in)."""
    import mmvae_b200.workloads as W
    if args.cpu_batch:
        return args.cpu_batch
    per_sample = W.algorithmic_bytes(dict(cfg, B=1)) / 4.0
    return int(max(1, min(B, CPU_ELEMS_PER_STEP // per_sample)))


def cpu_steps(cfg, t, n_warm, n_steps=None, budget_s=None):
    from oracle import leafstep
    for _ in range(n_warm):
        leafstep.run(cfg, t)
    n, t0 = 0, time.perf_counter()
    while (n_steps is not None and n < n_steps) or \
            (n_steps is None and (n < 3 or (time.perf_counter() - t0 < budget_s and n < 200))):
        leafstep.run(cfg, t)
        n += 1
    return n, time.perf_counter() - t0


def reference_arm(args, rank, world):
    """The reference's algorithm for this path on the box's host cores: /root/reference (pure Python/torch, nothing to
    compile) is absent on the GPU box, so this is the oracle port (oracle/refmath.py, pinned to the in-place reference by
    oracle/validate_against_reference.py) with all host threads.  Same workload, same batch per step as the GPU arm
    (bounded by an element budget for the big sweep workloads; the sample is stated)."""
    if rank != 0:
        return
    import torch
    import mmvae_b200.synthetic as syn
    import mmvae_b200.workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, scaling = local_batch(args, syn.WORKLOADS[args.workload]["B"], world)
    Bc = cpu_sample_batch(args, syn.WORKLOADS[args.workload], B)
    cfg, t = W.make_leaves(args.workload, B=Bc, seed=1234)
    n, dt = cpu_steps(cfg, t, 1, n_steps=args.steps)
    val = Bc * n / dt
    sample = "%d of %d samples per step of %s (same shapes, K=%d), %d steps, %.1f s" % (Bc, B, args.workload, cfg["K"], n, dt)
    line = {"metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": n,
            "warmup": 1, "ms_per_step": 1e3 * dt / n, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config_dict(args, cfg, B, world, scaling),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline(args, cfg_full, B):
    import torch
    import mmvae_b200.workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bc = cpu_sample_batch(args, cfg_full, B)
    cfg, t = W.make_leaves(args.workload, B=Bc, seed=1234)
    n, dt = cpu_steps(cfg, t, 1, budget_s=args.cpu_budget_s)
    return {"value": Bc * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d of %d samples per step of %s (same shapes, K=%d), %d steps, %.1f s" % (
                Bc, B, args.workload, cfg["K"], n, dt)}


class PluginStep:
    """The call a user of the reference makes: model.objective(batch) + backward through the drop-in plugin, batch from
    pinned host memory, replicated-parameter gradients all-reduced (SUM) at N>1, loss read back."""

    def __init__(self, cfg, B, dev, group, world, rank, fused_tail, graphed, fused_enc=False):
        import torch
        import mmvae_b200
        import mmvae_b200.parallel as par
        import mmvae_b200.synthetic as syn
        self.torch = torch
        torch.manual_seed(99)  # identical replicated parameters on every rank
        g = syn.gen(4321 + rank)
        vaes, self.host = {}, {}
        pv = cfg.get("private")
        for i, m in enumerate(cfg["mods"]):
            name = "mod_%d" % (i + 1)
            dz = cfg["D"] + (pv or 0)
            enc = syn.LinearEncoder(m["data_dim"], dz, returns_raw_logvar=fused_enc)
            dec = syn.LinearDecoder(dz, m["data_dim"], squash=(m["ltype"] == "bce"), returns_logits=fused_tail)
            vaes[name] = syn.StubVAE(enc, dec, cfg["D"], m["ltype"], private_latents=pv, llik_scaling=m["lam"],
                                     prior_dist=m["dist"], id_name=name)
            self.host[name] = syn.make_target(g, m["target"], B, m["data_dim"]).pin_memory()
        self.model = mmvae_b200.MODEL_REGISTRY[cfg["model"]](
            vaes, cfg["D"], {"obj": cfg["obj"], "beta": 1.0, "K": cfg["K"]}, None).to(dev)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.dev, self.B, self.world = dev, B, world
        self.sync = None
        if world > 1:
            par.attach(self.model, group, B * world)
            self.sync = par.GradSync(par.nccl_synced_params(self.model), group)
        self.gobj = None
        if graphed:  # the whole plugin step (encoders, kernels, decoders, backward) as ONE captured graph
            self.gobj = mmvae_b200.GraphedObjective(
                self.model, {k: {"data": v.to(dev), "masks": None, "categorical": False} for k, v in self.host.items()})
        self.h2d = sum(v.numel() * v.element_size() for v in self.host.values())
        # input pipeline: the batch of step i+1 is copied host -> device on a copy stream while step i computes (what a
        # DataLoader with pinned memory and non_blocking copies gives a training loop); two sets of device buffers.
        # Every step still copies its own inputs from pinned host memory inside the timed region.
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._bufs = [{k: torch.empty_like(v, device=dev) for k, v in self.host.items()} for _ in range(2)]
        self._ready = [None, None]
        self._slot = 0
        self._prefetch(0)
        self.api = "mmvae_b200.%s(vaes, ...).objective(batch) + backward, linear stand-in encoders/decoders (torch), %s%s%s" % (
            cfg["model"], "mmvae_b200.GraphedObjective (one CUDA-graph replay per step), double-buffered H2D of the next batch behind the "
            "current step" if graphed else "eager launches, double-buffered H2D",
            ("; decoder tail sigmoid+clamp fused into the likelihood kernel (bce_logits)" if fused_tail else "") +
            ("; encoder tail softmax+1e-6 fused into the latent kernels" if fused_enc else ""),
            "; flat-bucket NCCL all-reduce(SUM) of the parameter gradients" if world > 1 else "")

    def _prefetch(self, slot):
        torch = self.torch
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))  # the slot's previous consumer is done
        with torch.cuda.stream(self.copy_stream):
            for k, v in self.host.items():
                self._bufs[slot][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready[slot] = ev

    def step(self):
        torch = self.torch
        slot = self._slot
        torch.cuda.current_stream(self.dev).wait_event(self._ready[slot])  # this step's inputs have arrived
        batch = {k: {"data": v, "masks": None, "categorical": False} for k, v in self._bufs[slot].items()}
        self._slot = 1 - slot
        self._prefetch(self._slot)  # the next step's inputs travel while this step computes
        if self.gobj is not None:
            loss = self.gobj.step(batch)["loss"]
        else:
            for p in self.params:
                p.grad = None
            loss = self.model.objective(batch)["loss"]
            loss.backward()
        if self.sync is not None:
            self.sync.sync()
        return float(loss.detach())  # device -> host read of the step's result

    def close(self):
        if self.gobj is not None:
            self.gobj.close()


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries print there too (NCCL's version banner at NCCL_DEBUG=VERSION
    / WARN goes to fd 1 of every rank).  Keep a private duplicate of fd 1 for the result and point fd 1 at stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import mmvae_b200._lib as L
    import mmvae_b200.parallel as par
    import mmvae_b200.synthetic as syn
    import mmvae_b200.workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)  # capture warm-up runs on a side stream
    if world > 1:
        # NCCL prints its version banner on stdout at VERSION and WARN level: send its log to a file so that stdout
        # carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/mmvae_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    L.load()
    # collectives of the step: fused into our kernels over NVLink peer memory (parallel.PeerGroup, csrc/peer.cuh) unless
    # --nccl-only or the symmetric-memory allocation is unavailable; NCCL stays for the plugin's parameter buckets
    coll, coll_note = group, "none (1 GPU)"
    if world > 1:
        coll_note = "NCCL (in-graph all-reduce on a side stream)"
        if not args.nccl_only:
            try:
                coll = par.PeerGroup(group, dev)
                coll_note = ("fused into the kernels over NVLink peer memory (prior-logit gradient sync inside the prior-scale "
                             "backward, DReG (M,K) batch sums inside stage 2, optimal_sigma sum+count): no NCCL call in the step")
            except Exception as ex:
                coll_note = "NCCL (peer memory unavailable: %s)" % repr(ex)[:160]
    ctx = dict(rank=rank, local_rank=local_rank, world=world, dev=dev, group=group, coll=coll, coll_note=coll_note)
    if args.sweep:
        # several configurations inside ONE process group (a torchrun start-up per point costs more GPU time than the
        # points themselves): "workload:global_batch[:w|s],..." -- w = weak (the number is the batch per GPU)
        for item in args.sweep.split(","):
            f = item.split(":")
            a = argparse.Namespace(**vars(args))
            a.workload, a.batch, a.global_batch = f[0], None, None
            if len(f) > 2 and f[2] == "w":
                a.batch = int(f[1])
            else:
                a.global_batch = int(f[1])
            line = one_config(a, ctx)
            if rank == 0:
                emit(line)
    else:
        line = one_config(args, ctx)
        if rank == 0:
            emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.destroy_process_group()


def one_config(args, ctx):
    """Measure one configuration (the body of the bench); returns the JSON line (a dict)."""
    import torch
    import torch.distributed as dist
    import mmvae_b200._lib as L
    import mmvae_b200.parallel as par
    import mmvae_b200.synthetic as syn
    import mmvae_b200.workloads as W
    rank, local_rank, world, dev = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    group, coll, coll_note = ctx["group"], ctx["coll"], ctx["coll_note"]
    rdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B, scaling = local_batch(args, syn.WORKLOADS[args.workload]["B"], world)
    cfg, t = W.make_leaves(args.workload, B=B, seed=1234 + rank, recon_dtype=rdt)
    # the only replicated parameter on this leaf protocol is the prior logit vector: all-reduce(SUM) its gradient.
    # Default: inside the step (parallel.GradSync hook -> side stream, overlapped with the likelihood backward and
    # captured in the step graph); --eager-sync: one eager NCCL call after the step.
    in_step_sync = world > 1 and not args.eager_sync
    step = W.LeafStep(cfg, t, device=dev, group=coll, global_batch=B * world, sync_grads=in_step_sync,
                      fold=args.fold_terms)
    step.streams = args.streams
    W_, K_ = max(args.warmup, 3), args.steps

    def sync_grads():
        if world > 1 and step.sync is None and step.peer is None and step.pz_logits.grad is not None:
            dist.all_reduce(step.pz_logits.grad, group=group)

    # launches of our kernels per step, counted on one eager step
    step.run()
    torch.cuda.synchronize()
    c0 = L.launch_count
    step.run()
    launches_per_step = L.launch_count - c0
    runner = step
    # forward collectives (DReG (M,K) batch sums, optimal_sigma sum + count) are issued on the current stream and are
    # captured in the step graph like the gradient all-reduce
    if not args.no_graph:
        runner = W.GraphedStep(step)
    sync_mode = "none (1 GPU)"
    if world > 1:
        if step.peer is not None:
            sync_mode = "fused into the prior-scale backward kernel (peer memory)%s" % (
                ", captured in the step graph" if runner is not step else "")
        elif step.sync is not None:
            sync_mode = ("in-step NCCL all-reduce on a side stream behind the latent backward, overlapped with the likelihood "
                         "backward%s" % (", captured in the step graph" if runner is not step else ""))
        else:
            sync_mode = "eager NCCL all-reduce after the step"

    def one():
        runner.run()
        sync_grads()

    def allmax(x):
        v = torch.tensor([float(x)], device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)

    def timed(nsteps, fn, flush=None):
        """Device time of nsteps back-to-back steps (one event pair), or -- with an L2 flush between the steps -- the sum
        of per-step event pairs; barrier + synchronize on both sides, MAX over ranks."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if flush is None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(nsteps):
                fn()
            b.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
        else:
            evs = []
            for _ in range(nsteps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        return allmax(ms)

    step_bytes = W.algorithmic_bytes(cfg, rdt) * B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_needed(step_bytes) else None
    for _ in range(W_):
        one()
    est_ms = timed(3, one, flush) / 3.0
    reps = max(1, int(math.ceil(MIN_TIMED_S * 1e3 / max(est_ms, 1e-3) / K_)))
    if flush is not None:
        reps = min(reps, max(1, 400 // K_))  # every flushed step costs a 256 MB write as well
    K_eff = K_ * reps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(K_eff, one, flush)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * K_eff / (ms / 1e3)

    # roofline of the dominant kernel: CUDA events around every launch of it, eager steps, same inputs
    dom = "mmvae_loglik_rowreduce_bwd" if cfg["obj"] != "elbo" else "mmvae_loglik_rowreduce_fused"
    if all(m["ltype"] == "optimal_sigma" for m in cfg["mods"]):
        dom = "mmvae_osigma_bwd"
    if cfg.get("latent_only"):
        dom = "mmvae_moe_logdens_bwd_rk"
    names = [dom, "mmvae_loglik_rowreduce_fwd", "mmvae_moe_logdens_fwd", "mmvae_moe_logdens_bwd_rk", "mmvae_osigma_fwd"]
    kt = KernelTimer(names)
    L.timer = kt
    step.streams = 1  # per-kernel durations: one kernel at a time (the step itself runs two streaming kernels at once)
    for _ in range(0 if args.no_roofline_timer else min(K_, 10)):
        step.run()
    torch.cuda.synchronize()
    step.streams = args.streams
    L.timer = None
    tms, dom_bytes = kt.biggest(dom)  # launches of the largest term, bytes from the actual arguments
    peak, peak_src = measured_peak()
    ach = dom_bytes / (statistics.mean(tms) * 1e-3) / 1e9 if tms else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom if args.workload == DEFAULT_WORKLOAD and B == 256 else "")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": dom_bytes, "avg_ms": statistics.mean(tms) if tms else None,
                "launches_timed": len(tms)}
    for key, nm in (("fwd_kernel", "mmvae_loglik_rowreduce_fwd"), ("moe_fwd_kernel", "mmvae_moe_logdens_fwd"),
                    ("moe_bwd_kernel", "mmvae_moe_logdens_bwd_rk"), ("osigma_fwd_kernel", "mmvae_osigma_fwd")):
        kms, kb = kt.biggest(nm)
        if kms and nm != dom:
            roofline[key] = {"achieved": kb / (statistics.mean(kms) * 1e-3) / 1e9, "bytes_per_launch": kb,
                             "avg_ms": statistics.mean(kms)}
    roofline["step"] = {"algorithmic_bytes": step_bytes, "achieved": step_bytes * K_eff / (ms * 1e-3) / 1e9,
                        "frac": step_bytes * K_eff / (ms * 1e-3) / 1e9 / peak}

    # ---- end to end --------------------------------------------------------------------------------------------
    e2e = e2e_leaf = e2e_variants = None
    ke = max(3, min(K_, 10))
    if not args.no_e2e:
        # (a) leaf protocol: every input of the step lives in pinned host memory; H2D + loss read-back in the timed region
        pairs = [(step.mu, t["mu"]), (step.s, t["s"]), (step.pz_logits, t["pz_logits"])]
        host_recon = t["recon"]
        if step.fold:  # the folded leaves are the per-term tensors of a modality back to back
            host_recon = [torch.cat([t["recon"][i] for i in step.fold_index[tm]], 0) for tm in sorted(step.fold_index)]
        pairs += list(zip(step.targets, t["targets"])) + list(zip(step.recon, host_recon))
        if cfg["model"] == "moe":
            pairs.append((step.eps_stacked, torch.stack(t["noise"])))
            if step.dz is not None:
                pairs.append((step.dz, t["dz"]))
        else:  # the draws kernel reads one packed noise buffer
            pairs.append((step.eps_packed, torch.cat([n.reshape(-1) for n in t["noise"]])))
        pairs = [(d, h.contiguous().pin_memory()) for d, h in pairs]
        h2d = sum(h.numel() * h.element_size() for _, h in pairs)
        copy_stream = torch.cuda.Stream()

        def e2e_step():
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                for d, h in pairs:
                    d.data.copy_(h, non_blocking=True)
            torch.cuda.current_stream().wait_stream(copy_stream)
            loss = runner.run()
            sync_grads()
            return float(loss.detach())  # device -> host read of the step's result

        for _ in range(3):
            e2e_step()
        ems = timed(ke, e2e_step)
        e2e_leaf = {"value": B * world * ke / (ems / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ems / ke, "steps": ke,
                    "api": "mmvae_b200.workloads.LeafStep / GraphedStep (C-ABI kernels), every leaf from pinned host memory"}
        del pairs
        # (b) plugin level: the call a user of the reference makes, batch from pinned host memory
        if not cfg.get("latent_only"):
            e2e_variants = {}
            fused = any(m["ltype"] == "bce" for m in cfg["mods"])
            for key, kw in (("graphed", dict(fused_tail=fused, graphed=True, fused_enc=True)),
                            ("eager", dict(fused_tail=False, graphed=False))):
                if key == "eager" and world > 1:
                    continue
                try:
                    ps = PluginStep(cfg, B, dev, coll, world, rank, **kw)
                    for _ in range(3):
                        ps.step()
                    pms = timed(ke, ps.step)
                    e2e_variants[key] = {"value": B * world * ke / (pms / 1e3), "unit": "samples/s",
                                         "h2d_bytes_per_step": ps.h2d, "d2h_bytes_per_step": 4, "ms_per_step": pms / ke,
                                         "steps": ke, "api": ps.api}
                    ps.close()
                    del ps
                except Exception as ex:  # never let an extra leg hide the main numbers
                    e2e_variants[key] = {"error": repr(ex)[:300]}
            if "value" in e2e_variants.get("graphed", {}):
                e2e = e2e_variants.pop("graphed")
        if e2e is None:
            e2e = e2e_leaf

    # ---- multi-GPU numerical parity (outside the timed region) --------------------------------------------------
    parity_n = None
    if world > 1 and not args.no_parity:
        # the benchmarked step's own ingredients on a small global batch that does NOT divide evenly: sharded (graph
        # captured, in-step gradient all-reduce, forward collectives) vs the full batch on one GPU
        per = {}
        for name in dict.fromkeys([args.workload, "c2_moe_iwae_cdsprites_l5", "c4_moe_dreg_mnistsvhn",
                                   "c3_mopoe_elbo_vilanro"]):
            try:
                per[name] = allmax(par.sharded_parity(name, 4 * world + 1, coll, dev))
            except Exception as ex:
                per[name] = repr(ex)[:200]
        nums = [v for v in per.values() if isinstance(v, float)]
        parity_n = {"max_rel": max(nums) if nums else None, "per_workload": per, "global_batch": 4 * world + 1,
                    "peer_timeouts": bool(coll.error()) if isinstance(coll, par.PeerGroup) else None,
                    "what": "sharded (CUDA graph, collectives inside the step) vs full batch on one GPU: summed loss, "
                            "all-reduced prior gradient, shard rows of d/dmu, d/ds, d/drecon[0]; max relative deviation "
                            "over ranks"}

    conf = config_dict(args, cfg, B, world, scaling)
    if step.fold:
        conf["leaf_protocol"] = ("folded: one (terms*B, ...) reconstruction leaf and one likelihood launch per modality "
                                 "(decoders that fold K); algorithmic bytes unchanged (2R+T per term)")
    run = {"grad_sync": sync_mode, "collectives": coll_note, "launch_mode": "cuda-graph" if runner is not step else "eager",
           "streams": "3 (likelihood terms alternate between two streams, latent kernels on a third; "
                      "forks/joins captured in the graph)" if step.streams == 3 else "single stream"}
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K_, "timed_steps": K_eff,
            "warmup": W_, "ms_per_step": ms / K_eff, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if rdt == torch.bfloat16 else "f32", "data": "synthetic", "impl": "ours",
            "config": conf, "run": run, "roofline": roofline, "gpu_launches": launches_per_step * K_eff,
            "launches_per_step": launches_per_step, "clocks": clocks}
    if e2e is not None:
        line["e2e"] = e2e
        line["e2e_leaf"] = e2e_leaf
        if e2e_variants:
            line["e2e_plugin_variants"] = e2e_variants
    if parity_n is not None:
        line["parity_n"] = parity_n
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, syn.WORKLOADS[args.workload], B)
    if runner is not step:
        runner.close()  # the graph holds the captured collectives: it must go before the communicator
    if step.sync is not None:
        step.sync.disarm()
    del runner, step
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    main()
