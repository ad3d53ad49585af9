"""Batch-sharded data parallelism for the hot path: one process per GPU (torchrun), NCCL over NVLink/NVSwitch.

The path shards naturally over batch rows: fusion, sampling, log-densities, KL and the likelihood row
reductions are row independent and every ELBO / IWAE loss is a plain sum over b.  So there is NO data-path
collective; the only communication is

  * one all-reduce(SUM) per step of a single flat bucket holding the gradients of the replicated parameters
    (encoder / decoder weights and the prior logits ``_pz_params[1]``) -- SUM, not mean: the reference's losses are
    sums over the batch;
  * three tiny forward exchanges that batch-GLOBAL statistics need for exact parity:
      (1) DReG parity mode: the (M,K) batch-summed log-weights are all-reduced between the two combine stages
          (ops._Dreg, include/mmvae_b200.h mmvae_objective_dreg_stage{1,2});
      (2) MoPoE: batch means use the global batch size (``model.global_batch``), no communication;
      (3) optimal_sigma: the scalar sum of squares is all-reduced between its stages (ops._OsigmaRows).
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


class PeerGroup:
    """One NVLink / NVSwitch domain seen through peer-mapped symmetric buffers: the handle the fused
    compute + collective kernels (include/mmvae_b200.h mmvae_*_peer) exchange their small reductions through, instead
    of going kernel -> NCCL all-reduce -> kernel.  PyTorch only provides the memory: the buffer comes from
    torch.distributed._symmetric_memory (CUDA VMM allocation exported to the peers of `group`), the protocol and the
    kernels are ours (csrc/peer.cuh).  Wherever this package takes a process `group`, a PeerGroup may be passed
    instead; NCCL (``.pg``) stays in use for what does not fit a 4 KB slot (encoder / decoder gradient buckets)."""

    CH_PRIOR_GRAD, CH_DREG, CH_OSIGMA, CH_OSIGMA_BWD = 0, 1, 2, 4  # channels (osigma: + term parity)

    def __init__(self, group=None, device=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.pg = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        if self.world > _lib.PEER_MAX_WORLD:
            raise RuntimeError("mmvae_b200: PeerGroup supports up to %d ranks" % _lib.PEER_MAX_WORLD)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty(_lib.PEER_BUFFER_BYTES, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.pg.group_name)
        torch.cuda.synchronize(device)
        dist.barrier(self.pg)  # every rank's buffer is zeroed before anyone publishes into it
        self.bufs_dev = ctypes.c_void_p(int(self.hdl.buffer_ptrs_dev))
        self._err_off = int(_lib.load().mmvae_peer_error_offset())

    def error(self) -> bool:
        """True if a wait of any fused collective on this rank ever timed out (a peer did not show up within ~2 s)."""
        return bool(self.buf[self._err_off:self._err_off + 4].view(torch.int32).item())


def process_group(group):
    """The torch process group behind `group` (a PeerGroup or a process group / None)."""
    return group.pg if isinstance(group, PeerGroup) else group


def shard_range(n_global: int, rank: int, world: int):
    """Contiguous global-row range [lo, hi) of `rank`: rows [g*B/G, (g+1)*B/G), remainder to the first ranks."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Slice every per-sample tensor of a reference-format batch dict ({"mod_i": {"data","masks","categorical"}})."""
    out = {}
    for mod, entry in batch.items():
        n = None
        for v in entry.values():
            if torch.is_tensor(v):
                n = v.shape[0]
                break
        new = dict(entry)
        if n is not None:
            lo, hi = shard_range(n, rank, world)
            for k, v in entry.items():
                if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                    new[k] = v[lo:hi]
        out[mod] = new
    return out


def shard_leaves(cfg: dict, tensors: dict, rank: int, world: int):
    """Rows [lo, hi) of a leaf-tensor set (workloads.make_leaves): encoder outputs (M,B,D), targets (B,..), noise
    (K,B,D) and the k-major reconstructions (row = k*B + b) / synthetic likelihood rows; the prior logits are replicated.
    Returns (cfg with the local B, sharded tensors, (lo, hi))."""
    B = cfg["B"]
    lo, hi = shard_range(B, rank, world)
    out = {"mu": tensors["mu"][:, lo:hi].contiguous(), "s": tensors["s"][:, lo:hi].contiguous(),
           "pz_logits": tensors["pz_logits"], "targets": [x[lo:hi].contiguous() for x in tensors["targets"]]}
    out["recon"] = [r.view(r.shape[0] // B, B, *r.shape[1:])[:, lo:hi].reshape(-1, *r.shape[1:]).contiguous()
                    for r in tensors["recon"]]
    out["noise"] = [n[:, lo:hi].contiguous() for n in tensors["noise"]]
    if tensors.get("dz") is not None:
        out["dz"] = tensors["dz"][:, :, lo:hi].contiguous()
    c = dict(cfg)
    c["B"] = hi - lo
    return c, out, (lo, hi)


def sharded_parity(name: str, B: int, group, device, seed: int = 11):
    """Numerical parity of the batch-sharded CUDA path: every rank runs its shard of a small global batch (captured in a
    CUDA graph together with its collectives, like the benchmarked step) and compares with the same objective on the
    FULL batch run locally: summed loss, all-reduced prior-logit gradient, the shard's rows of d/dmu, d/ds and of the
    first reconstruction gradient.  Returns the largest relative deviation seen on this rank."""
    from . import workloads as W
    pg = process_group(group)
    rank, world = dist.get_rank(pg), dist.get_world_size(pg)
    cfg, t = W.make_leaves(name, B=B, seed=seed)
    t["pz_logits"] = torch.randn(1, cfg["D"], generator=torch.Generator().manual_seed(2)) * 0.3
    full = W.LeafStep(cfg, t, device=device)
    full_loss = full.run().detach().clone()
    c2, t2, (lo, hi) = shard_leaves(cfg, t, rank, world)
    step = W.LeafStep(c2, t2, device=device, group=group, global_batch=B, sync_grads=True)
    gs = W.GraphedStep(step)
    for _ in range(2):
        loss = gs.run()
    loss = loss.detach().clone()
    if cfg["obj"] != "dreg":  # the DReG loss is already a global quantity (computed from all-reduced batch sums)
        dist.all_reduce(loss, group=pg)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    K = cfg["K"] if cfg["model"] == "moe" else 1
    errs = [rel(loss, full_loss), rel(step.mu.grad, full.mu.grad[:, lo:hi]), rel(step.s.grad, full.s.grad[:, lo:hi])]
    if full.pz_logits.grad is not None and float(full.pz_logits.grad.abs().max()) > 0:
        errs.append(rel(step.pz_logits.grad, full.pz_logits.grad))
    g_full = full.recon[0].grad
    g_full = g_full.view(K, B, *g_full.shape[1:])[:, lo:hi].reshape(step.recon[0].grad.shape)
    errs.append(rel(step.recon[0].grad, g_full))
    gs.close()  # a graph that captured the communicator must be destroyed before the process group
    if step.sync is not None:
        step.sync.disarm()
    return max(errs)


def attach(model, group=None, global_batch: int = None):
    """Tell a drop-in model plugin that it sees one shard of a global batch.  With a PeerGroup the forward exchanges AND
    the gradient sync of the replicated prior logits run inside the fused peer-memory kernels."""
    model.group = group
    model.obj_fn.group = group
    model.global_batch = global_batch
    return model


def nccl_synced_params(model):
    """Parameters whose gradients still need the flat-bucket NCCL all-reduce (GradSync): everything trainable, except
    the prior logits when the model is attached to a PeerGroup (their gradient leaves prior_scale's backward kernel
    already summed over the ranks)."""
    skip = id(model._pz_params[1]) if isinstance(getattr(model, "group", None), PeerGroup) else None
    return [p for p in model.parameters() if p.requires_grad and id(p) != skip]


class GradSync:
    """All-reduce(SUM) of the gradients of replicated parameters through ONE flat bucket per step.

    NVSwitch gives every GPU full bandwidth to every peer, so the bucket is sized for launch latency: a single
    collective per step instead of one per parameter.

    Two ways to run it:
      * ``sync()`` (or calling the object) after ``backward()``: the collective sits at the end of the step;
      * ``arm()`` once, then ``wait()`` after every ``backward()``: post-accumulate-grad hooks launch the collective on
        a SIDE stream as soon as the last bucket gradient has been accumulated, so it overlaps whatever backward work
        is still queued (on this path: the bandwidth-bound likelihood backward kernels, 250 us at C2, hide the
        ~30 us latency-bound all-reduce and the skew between ranks).  Both the fork and the join are plain stream
        waits, so the whole thing is capturable in the step's CUDA graph (NCCL supports capture).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group(group)
        self._flat = None
        self._side = None
        self._handles = []
        self._seen = 0
        self._inflight = False

    def active(self) -> bool:
        return bool(self.params) and dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(self.group) > 1

    def _bucket(self):
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        if self._flat is None or self._flat.numel() != n or self._flat.device != ref.device:
            self._flat = torch.empty(n, dtype=torch.float32, device=ref.device)
        return self._flat

    def _reduce(self):
        if len(self.params) == 1 and self.params[0].grad is not None and self.params[0].grad.dtype == torch.float32 \
                and self.params[0].grad.is_contiguous():
            dist.all_reduce(self.params[0].grad, op=dist.ReduceOp.SUM, group=self.group)  # in place, no bucket copy
            return
        flat = self._bucket()
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p in self.params:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n

    def sync(self):
        if self.active():
            self._reduce()

    __call__ = sync

    # -- overlapped mode -------------------------------------------------------------------------------------------
    def arm(self):
        """Register the hooks (idempotent).  Every parameter of the bucket must receive a gradient in every backward."""
        if self._handles or not self.active():
            return self
        dev = self.params[0].device
        self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None  # CPU tensors: reduce in the hook
        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def disarm(self):
        for h in self._handles:
            h.remove()
        self._handles, self._seen, self._inflight = [], 0, False

    def _on_grad(self, _p):
        self._seen += 1
        if self._seen < len(self.params):
            return
        self._seen = 0
        if self._side is None:
            self._reduce()
        else:
            cur = torch.cuda.current_stream()  # the stream the gradient was accumulated on
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._reduce()
        self._inflight = True

    def wait(self):
        """Join the side stream: after this the current stream sees the reduced gradients.  If the hooks did not
        complete the bucket in this backward (a parameter received no gradient -- e.g. the prior logits under MoE-ELBO),
        the collective runs here instead, so every rank still issues exactly one all-reduce per step."""
        if not self._handles:
            return
        if self._inflight:
            if self._side is not None:
                cur = torch.cuda.current_stream()
                cur.wait_stream(self._side)
                # tensors created on the side stream (the flat bucket, gradients materialised for parameters that got
                # none) are consumed on this stream from here on: tell the caching allocator, or a later free could hand
                # their memory back to the side stream while work queued here still reads it
                if self._flat is not None:
                    self._flat.record_stream(cur)
                for p in self.params:
                    if p.grad is not None:
                        p.grad.record_stream(cur)
        else:
            self._reduce()
        self._inflight, self._seen = False, 0
