"""torch.autograd.Functions over the C ABI (include/mmvae_b200.h).

PyTorch is plumbing here: it owns device memory, the CUDA stream and autograd bookkeeping.  All arithmetic on the
hot path happens in the sm_100a kernels of libmmvae_b200.so.  Every op raises on CPU tensors -- there is no
fallback (north_star: "no CPU fallback").
"""
import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import DrawDesc, call

_P = ctypes.c_void_p


# Upstream gradients that are known to be exactly 1 WITHOUT looking at device memory: a caller that starts its backward
# from a static ones tensor (``loss.backward(one)``) registers it here; a backward node whose incoming gradient has the
# same address can then skip its "scale by grad_output" launch altogether.  The registry keeps the tensors alive, so a
# registered address can never be recycled for other data (the owner must not write to it).
_UNIT_GRADS = {}
_TICKETS = {}


def mark_unit_grad(t: torch.Tensor) -> torch.Tensor:
    _UNIT_GRADS[t.data_ptr()] = t
    return t


def unmark_unit_grad(t: torch.Tensor) -> None:
    """Drop a tensor registered with mark_unit_grad (its owner is going away)."""
    _UNIT_GRADS.pop(t.data_ptr(), None)


def _is_unit(g: Optional[torch.Tensor]) -> bool:
    return g is not None and g.numel() == 1 and g.data_ptr() in _UNIT_GRADS


def new_ticket(device) -> torch.Tensor:
    """A zero-initialised counter for kernels that elect their last CTA; the kernel leaves it at zero, so one ticket
    serves every launch of its owner as long as those launches are stream ordered."""
    return torch.zeros(1, dtype=torch.int32, device=device)


def _ticket(device) -> torch.Tensor:
    """Default ticket: cached per (device, stream) for eager launches; a fresh one inside a CUDA-graph capture (it then
    lives in that graph's memory pool -- a cached tensor must not outlive the graph it was captured in)."""
    if torch.cuda.is_current_stream_capturing():
        return new_ticket(device)
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    if key not in _TICKETS:
        _TICKETS[key] = new_ticket(device)
    return _TICKETS[key]


def _ptr(t: Optional[torch.Tensor]):
    return _P(0) if t is None else _P(t.data_ptr())


def _stream():
    return _P(torch.cuda.current_stream().cuda_stream)


def _peer(group):
    """(peer, process_group): `group` may be a parallel.PeerGroup (fused peer-memory kernels) or a torch process group."""
    if group is not None and hasattr(group, "bufs_dev"):
        return group, group.pg
    return None, group


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmvae_b200 ops need CUDA tensors (no CPU fallback); got a tensor on %s" % t.device)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise RuntimeError("mmvae_b200: unsupported dtype %s (float32 / bfloat16 only)" % t.dtype)


def _rows2d(t: torch.Tensor, rows: int):
    """View a (rows, ...) tensor as rows x P with a row stride; copies only if the trailing dims are not dense."""
    t2 = t.reshape(rows, -1) if t.is_contiguous() else t
    if t2.dim() != 2:
        # trailing dims must be dense so that a row is one contiguous run of P elements
        P = 1
        for sz in t.shape[1:]:
            P *= sz
        exp = 1
        dense = True
        for sz, stt in zip(reversed(t.shape[1:]), reversed(t.stride()[1:])):
            if sz != 1 and stt != exp:
                dense = False
            exp *= sz
        if dense and t.stride(0) >= P:
            t2 = t.as_strided((rows, P), (t.stride(0), 1))
        else:
            t2 = t.contiguous().reshape(rows, -1)
    if t2.stride(1) != 1:
        t2 = t2.contiguous()
    return t2, t2.shape[1], t2.stride(0)


LTYPES = {"bce": _lib.LT_BCE, "mse": _lib.LT_MSE, "l1": _lib.LT_L1, "bce_logits": _lib.LT_BCE_LOGITS}


def ltype_code(ltype: str, likelihood: str) -> int:
    if ltype == "lprob":
        return _lib.LT_LPROB_LAPLACE if likelihood == "laplace" else _lib.LT_LPROB_NORMAL
    if ltype == "lprob_selfscale":  # lprob with padding masks: scale := loc (reference objectives.py:43-45)
        return _lib.LT_LPROB_LAPLACE_SELF if likelihood == "laplace" else _lib.LT_LPROB_NORMAL_SELF
    return LTYPES[ltype]


# ----------------------------------------------------------------------------------------------------------
# likelihood rows (element-wise families)
# ----------------------------------------------------------------------------------------------------------
class _LoglikRows(torch.autograd.Function):
    """rows[r] = lam * sum_p logp(recon[r,p] | target[r % B, p]); separate fwd / bwd kernels (weights unknown a
    priori: IWAE / DReG).  include/mmvae_b200.h mmvae_loglik_rowreduce_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, recon, target, lt, scale, lam, out=None):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(recon, target)
        rows, B = recon.shape[0], target.shape[0]
        x, P, ldx = _rows2d(recon.detach(), rows)
        t, Pt, ldt = _rows2d(target.detach(), B)
        if Pt != P or rows % B != 0:
            raise RuntimeError("mmvae_b200: recon rows x P (%d x %d) incompatible with target (%d x %d)" % (rows, P, B, Pt))
        out = _row_out(out, rows, recon.device)
        nws = _lib.load().mmvae_loglik_workspace_bytes(rows, P, _dt(x))
        ws = torch.empty(max(nws // 4, 1), dtype=torch.float32, device=recon.device)
        call("mmvae_loglik_rowreduce_fwd", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, lt, scale, lam,
             _ptr(out), _ptr(ws), _stream())
        ctx.save_for_backward(x, t)
        ctx.meta = (rows, B, P, ldx, ldt, lt, scale, lam, recon.shape)
        return out

    @staticmethod
    def backward(ctx, g_rows):
        if g_rows is None:
            return None, None, None, None, None, None
        x, t = ctx.saved_tensors
        rows, B, P, ldx, ldt, lt, scale, lam, shape = ctx.meta
        w = g_rows.detach().to(torch.float32).contiguous()
        g = torch.empty((rows, P), dtype=x.dtype, device=x.device)
        call("mmvae_loglik_rowreduce_bwd", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, lt, scale, lam,
             _ptr(w), _ptr(g), P, _stream())
        return g.view(shape), None, None, None, None, None


class _LoglikWeightedSum(torch.autograd.Function):
    """S = sum_r w_r * rows[r] with the gradient buffer produced in the SAME pass (weights known a priori: every
    ELBO).  Returns (S, rows); rows is a non-differentiable by-product for logging.  backward rescales the stored
    gradient in place by grad_output through a kernel that exits immediately when grad_output == 1."""

    @staticmethod
    def forward(ctx, recon, target, w_rows, w_const, lt, scale, lam, defer=False):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(recon, target, w_rows)
        rows, B = recon.shape[0], target.shape[0]
        x, P, ldx = _rows2d(recon.detach(), rows)
        t, Pt, ldt = _rows2d(target.detach(), B)
        if Pt != P or rows % B != 0:
            raise RuntimeError("mmvae_b200: recon rows x P (%d x %d) incompatible with target (%d x %d)" % (rows, P, B, Pt))
        out = torch.empty(rows, dtype=torch.float32, device=recon.device)
        g = torch.empty((rows, P), dtype=x.dtype, device=x.device)
        nws = _lib.load().mmvae_loglik_workspace_bytes(rows, P, _dt(x))
        ws = torch.empty(max(nws // 4, 1), dtype=torch.float32, device=recon.device)
        w = None if w_rows is None else w_rows.detach().to(torch.float32).contiguous()
        call("mmvae_loglik_rowreduce_fused", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, lt, scale, lam,
             _ptr(w), float(w_const), _ptr(out), _ptr(g), P, _ptr(ws), _stream())
        S = torch.empty((), dtype=torch.float32, device=recon.device)
        if defer:  # the batch sum is left to ops.elbo_combine (one launch for all terms): S is a placeholder
            if w is not None:
                raise RuntimeError("mmvae_b200: deferred term sums need a-priori constant weights (w_const)")
        elif w is None:
            call("mmvae_reduce_sum", _ptr(out), rows, float(w_const), _ptr(S), _stream())
        else:
            S = torch.dot(out, w)
        ctx.g = g
        ctx.rows_out = out
        ctx.shape = recon.shape
        ctx.w_needs = w_rows is not None and w_rows.requires_grad
        ctx.mark_non_differentiable(out)
        return S, out

    @staticmethod
    def backward(ctx, gS, _g_rows):
        if gS is None:
            return None, None, None, None, None, None, None, None
        if ctx.g is None:
            raise RuntimeError("mmvae_b200: the fused ELBO gradient buffer is single-use (retain_graph unsupported)")
        g, ctx.g = ctx.g, None
        gs = gS.detach().to(torch.float32).contiguous()
        if not _is_unit(gS):  # (the kernel itself exits at once when *gs == 1; a registered unit gradient skips the launch)
            call("mmvae_scale_inplace", _ptr(g), _dt(g), g.numel(), _ptr(gs), _stream())
        gw = (gs * ctx.rows_out) if ctx.w_needs else None
        return g.view(ctx.shape), None, gw, None, None, None, None, None


def _row_out(out, rows, device):
    """Destination of a row kernel: a fresh (rows,) fp32 tensor, or a caller-provided contiguous slice of a stacked
    buffer (so that the list of row vectors IS the stacked tensor, no concatenation copy)."""
    if out is None:
        return torch.empty(rows, dtype=torch.float32, device=device)
    if out.dtype != torch.float32 or out.numel() != rows or not out.is_contiguous() or out.device != device:
        raise RuntimeError("mmvae_b200: `out` must be a contiguous fp32 (rows,) slice on the same device")
    return out.view(rows)


def loglik_rows(recon, target, ltype, likelihood="normal", lam=1.0, scale=0.75, out=None):
    return _LoglikRows.apply(recon, target, ltype_code(ltype, likelihood), float(scale), float(lam), out)


def _deferred(S, rows, w_const, defer):
    """Tag a placeholder S with what ops.elbo_combine needs to do the sum itself: (row vector, coefficient)."""
    if defer:
        S._mmvae_deferred = (rows, float(w_const))
    return S, rows


def loglik_weighted_sum(recon, target, ltype, likelihood="normal", lam=1.0, scale=0.75, w_rows=None, w_const=1.0,
                        defer=False):
    """defer=True (constant weights only): skip the batch-sum launch; the returned S is only valid as a term of
    ops.elbo_combine, which sums the row vectors of every term in its single launch."""
    S, rows = _LoglikWeightedSum.apply(recon, target, w_rows, float(w_const), ltype_code(ltype, likelihood),
                                       float(scale), float(lam), bool(defer))
    return _deferred(S, rows, w_const, defer)


# ----------------------------------------------------------------------------------------------------------
# category_ce rows
# ----------------------------------------------------------------------------------------------------------
def _catce_geom(recon, target):
    """recon (rows, C, *rest) -> C, d = prod(rest); rows must be dense runs of C*d elements."""
    rows, B = recon.shape[0], target.shape[0]
    if recon.dim() < 2:
        raise RuntimeError("category_ce needs (rows, C, ...) reconstructions")
    C = recon.shape[1]
    d = 1
    for sz in recon.shape[2:]:
        d *= sz
    x, P, ldx = _rows2d(recon, rows)
    t, Pt, ldt = _rows2d(target, B)
    if P != C * d or Pt != P or rows % B != 0:
        raise RuntimeError("mmvae_b200: category_ce shapes %s vs %s" % (tuple(recon.shape), tuple(target.shape)))
    return x, t, rows, B, C, d, ldx, ldt


def _mask_bytes(mask, B, C):
    """(B, C) padding mask -> contiguous uint8 (what mmvae_catce_rows_masked reads); None stays None."""
    if mask is None:
        return None
    _need_cuda(mask)
    if mask.dim() != 2 or mask.shape[0] != B or mask.shape[1] < C:
        raise RuntimeError("mmvae_b200: padding mask %s does not cover the (B = %d, C = %d) class rows" % (
            tuple(mask.shape), B, C))
    m = mask.detach()
    return (m if m.dtype == torch.uint8 else m.to(torch.uint8)).contiguous()


class _CatceRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, target, lam, out=None, mask=None):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(recon, target)
        x, t, rows, B, C, d, ldx, ldt = _catce_geom(recon.detach(), target.detach())
        out = _row_out(out, rows, recon.device)
        stats = torch.empty((rows, 2, d), dtype=torch.float32, device=recon.device)  # cached column statistics
        mk = _mask_bytes(mask, B, C)
        call("mmvae_catce_rows_masked", 0, _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, C, d, lam, _P(0), 0.0,
             _ptr(out), _P(0), 0, _ptr(stats), _ptr(mk), 0 if mk is None else mk.stride(0), _stream())
        ctx.save_for_backward(x, t, stats, mk)
        ctx.meta = (rows, B, C, d, ldx, ldt, lam, recon.shape)
        return out

    @staticmethod
    def backward(ctx, g_rows):
        if g_rows is None:
            return None, None, None, None, None
        x, t, stats, mk = ctx.saved_tensors
        rows, B, C, d, ldx, ldt, lam, shape = ctx.meta
        w = g_rows.detach().to(torch.float32).contiguous()
        g = torch.empty((rows, C * d), dtype=x.dtype, device=x.device)
        call("mmvae_catce_rows_masked", 1, _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, C, d, lam, _ptr(w), 0.0,
             _P(0), _ptr(g), C * d, _ptr(stats), _ptr(mk), 0 if mk is None else mk.stride(0), _stream())
        return g.view(shape), None, None, None, None


class _CatceWeightedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, target, w_rows, w_const, lam, defer=False, mask=None):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(recon, target, w_rows)
        x, t, rows, B, C, d, ldx, ldt = _catce_geom(recon.detach(), target.detach())
        out = torch.empty(rows, dtype=torch.float32, device=recon.device)
        g = torch.empty((rows, C * d), dtype=x.dtype, device=x.device)
        w = None if w_rows is None else w_rows.detach().to(torch.float32).contiguous()
        mk = _mask_bytes(mask, B, C)
        call("mmvae_catce_rows_masked", 2, _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, C, d, lam, _ptr(w),
             float(w_const), _ptr(out), _ptr(g), C * d, _P(0), _ptr(mk), 0 if mk is None else mk.stride(0), _stream())
        S = torch.empty((), dtype=torch.float32, device=recon.device)
        if defer:  # the batch sum is left to ops.elbo_combine (one launch for all terms): S is a placeholder
            if w is not None:
                raise RuntimeError("mmvae_b200: deferred term sums need a-priori constant weights (w_const)")
        elif w is None:
            call("mmvae_reduce_sum", _ptr(out), rows, float(w_const), _ptr(S), _stream())
        else:
            S = torch.dot(out, w)
        ctx.g = g
        ctx.rows_out = out
        ctx.shape = recon.shape
        ctx.w_needs = w_rows is not None and w_rows.requires_grad
        ctx.mark_non_differentiable(out)
        return S, out

    @staticmethod
    def backward(ctx, gS, _g_rows):
        if gS is None:
            return None, None, None, None, None, None, None
        if ctx.g is None:
            raise RuntimeError("mmvae_b200: the fused ELBO gradient buffer is single-use (retain_graph unsupported)")
        g, ctx.g = ctx.g, None
        gs = gS.detach().to(torch.float32).contiguous()
        if not _is_unit(gS):
            call("mmvae_scale_inplace", _ptr(g), _dt(g), g.numel(), _ptr(gs), _stream())
        gw = (gs * ctx.rows_out) if ctx.w_needs else None
        return g.view(ctx.shape), None, gw, None, None, None, None


def catce_rows(recon, target, lam=1.0, out=None, mask=None):
    """mask: optional (B, C) padding mask of a text decoder that returns its output UNMASKED -- the kernel applies
    recon * mask[:, :, None] (reference decoders.py:722) itself and returns the gradient of the unmasked tensor."""
    return _CatceRows.apply(recon, target, float(lam), out, mask)


def catce_weighted_sum(recon, target, lam=1.0, w_rows=None, w_const=1.0, defer=False, mask=None):
    S, rows = _CatceWeightedSum.apply(recon, target, w_rows, float(w_const), float(lam), bool(defer), mask)
    return _deferred(S, rows, w_const, defer)


# ----------------------------------------------------------------------------------------------------------
# optimal_sigma rows
# ----------------------------------------------------------------------------------------------------------
class _OsigmaRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, target, lam, group, channel):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(recon, target)
        rows, B = recon.shape[0], target.shape[0]
        x, P, ldx = _rows2d(recon.detach(), rows)
        t, Pt, ldt = _rows2d(target.detach(), B)
        if Pt != P or rows % B != 0:
            raise RuntimeError("mmvae_b200: optimal_sigma shapes %s vs %s" % (tuple(recon.shape), tuple(target.shape)))
        stat = torch.zeros(2, dtype=torch.float64, device=recon.device)  # [sumsq, element count], both device side
        row_ss = torch.empty(rows, dtype=torch.float32, device=recon.device)  # per-row sums of squares: stage 2 needs no
        call("mmvae_osigma_sumsq", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, _ptr(stat), _ptr(row_ss),  # 2nd pass
             _stream())
        n_total = 0.0  # <= 0: the kernels read the count from stat[1]
        peer, pgroup = _peer(group)
        if peer is not None:  # global RMS over every shard: sum AND count through peer memory
            call("mmvae_peer_allreduce_f64", _ptr(stat), 2, peer.bufs_dev, peer.rank, peer.world,
                 peer.CH_OSIGMA + (channel & 1), _stream())
        elif pgroup is not None:  # ... or ONE tiny NCCL all-reduce of both, so uneven shards get the global mean
            import torch.distributed as dist
            dist.all_reduce(stat, group=pgroup)
        out = torch.empty(rows, dtype=torch.float32, device=recon.device)
        stats2 = torch.empty(2, dtype=torch.float32, device=recon.device)
        call("mmvae_osigma_fwd", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, lam, _ptr(stat), n_total,
             _ptr(out), _ptr(stats2), _ptr(row_ss), _stream())
        ctx.save_for_backward(x, t, stat)
        ctx.meta = (rows, B, P, ldx, ldt, lam, n_total, recon.shape, group, channel)
        return out

    @staticmethod
    def backward(ctx, g_rows):
        if g_rows is None:
            return None, None, None, None, None
        x, t, stat = ctx.saved_tensors
        rows, B, P, ldx, ldt, lam, n_total, shape, group, channel = ctx.meta
        w = g_rows.detach().to(torch.float32).contiguous()
        wsum = torch.empty(1, dtype=torch.float32, device=x.device)
        g = torch.empty((rows, P), dtype=x.dtype, device=x.device)
        peer, pgroup = _peer(group)
        if peer is not None or pgroup is not None:
            # the shards' rows all depend on every shard's x through the global sigma: sum_r w_r must be global too
            wtot = w.sum(dtype=torch.float64).reshape(1)
            if peer is not None:
                call("mmvae_peer_allreduce_f64", _ptr(wtot), 1, peer.bufs_dev, peer.rank, peer.world,
                     peer.CH_OSIGMA_BWD + (channel & 1), _stream())
            else:
                import torch.distributed as dist
                dist.all_reduce(wtot, group=pgroup)
            w = (wtot / rows).float() * torch.ones_like(w)  # same row-weight sum, fed through the unchanged kernel
        call("mmvae_osigma_bwd", _ptr(x), ldx, _dt(x), _ptr(t), ldt, _dt(t), rows, B, P, lam, _ptr(stat), n_total,
             _ptr(w), _ptr(wsum), _ptr(g), P, _stream())
        return g.view(shape), None, None, None, None


def osigma_rows(recon, target, lam=1.0, group=None, channel=0):
    """group: torch process group or parallel.PeerGroup of a batch-sharded run; channel: distinguishes optimal_sigma terms
    that may be in flight at the same time on different streams (peer mode)."""
    return _OsigmaRows.apply(recon, target, float(lam), group, int(channel))


# ----------------------------------------------------------------------------------------------------------
# latent draws
# ----------------------------------------------------------------------------------------------------------
class Draw:
    """Host-side description of one latent draw (mirrors mmvae_draw_desc)."""

    def __init__(self, mods: Sequence[int] = (), prior=False, direct=False, laplace=False, rowmask=False, kl_mode=0,
                 col0=0, width=0, K=0, want_params=False):
        self.mods, self.prior, self.direct, self.laplace, self.rowmask = tuple(mods), prior, direct, laplace, rowmask
        self.kl_mode, self.col0, self.width, self.K, self.want_params = kl_mode, col0, width, K, want_params


def _pack_descs(draws: List[Draw], B: int):
    n = len(draws)
    if n > _lib.MAX_DRAWS:
        raise RuntimeError("mmvae_b200: %d latent draws exceed the limit of %d" % (n, _lib.MAX_DRAWS))
    arr = (DrawDesc * n)()
    eo = po = ko = 0
    for j, d in enumerate(draws):
        mask = 0
        for m in d.mods:
            mask |= 1 << m
        flags = (_lib.DRAW_PRIOR if d.prior else 0) | (_lib.DRAW_DIRECT if d.direct else 0) | \
                (_lib.DRAW_LAPLACE if d.laplace else 0) | (_lib.DRAW_ROWMASK if d.rowmask else 0)
        arr[j].mask, arr[j].flags, arr[j].kl_mode = mask, flags, d.kl_mode
        arr[j].col0, arr[j].width, arr[j].K = d.col0, d.width, d.K
        arr[j].eps_off = arr[j].z_off = eo
        eo += d.K * B * d.width
        arr[j].par_off = po if d.want_params else -1
        if d.want_params:
            po += B * d.width
        arr[j].kl_off = ko if d.kl_mode else -1
        if d.kl_mode:
            ko += B
    return arr, eo, po, ko


class _LatentDraws(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_latent_draws_{fwd,bwd}.  Inputs: mu, s (M,B,Dtot), prior (mu0, s0) (D), packed
    noise.  Outputs: packed z, packed loc, packed scale, packed kl rows."""

    @staticmethod
    def forward(ctx, mu, s, mu0, s0, eps, draws, row_masks, s_raw=False):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(mu, s, mu0, s0, eps, row_masks)
        M, B, Dtot = mu.shape
        arr, ne, npar, nkl = _pack_descs(draws, B)
        mu_c, s_c = mu.detach().float().contiguous(), s.detach().float().contiguous()
        mu0_c = None if mu0 is None else mu0.detach().float().contiguous().reshape(-1)
        s0_c = None if s0 is None else s0.detach().float().contiguous().reshape(-1)
        eps_c = None if eps is None else eps.detach().float().contiguous()
        if ne and (eps_c is None or eps_c.numel() != ne):
            raise RuntimeError("mmvae_b200: packed noise has %s elements, draws need %d" % (
                None if eps_c is None else eps_c.numel(), ne))
        dev = mu.device
        z = torch.empty(max(ne, 1), dtype=torch.float32, device=dev)
        ploc = torch.empty(max(npar, 1), dtype=torch.float32, device=dev)
        pscale = torch.empty(max(npar, 1), dtype=torch.float32, device=dev)
        kl = torch.empty(max(nkl, 1), dtype=torch.float32, device=dev)
        # s_raw: `s` is the raw output of the encoders' second head; the kernel applies softmax(.,-1) + 1e-6 itself
        scales = torch.empty_like(s_c) if s_raw else None
        call("mmvae_latent_draws_fwd_tail", _ptr(mu_c), _ptr(s_c), M, B, Dtot, arr, len(draws), _ptr(row_masks),
             _ptr(mu0_c), _ptr(s0_c), _ptr(eps_c), _ptr(z), _ptr(ploc), _ptr(pscale), _ptr(kl), int(bool(s_raw)),
             _ptr(scales), _stream())
        ctx.save_for_backward(mu_c, s_c, mu0_c, s0_c, eps_c, row_masks)
        ctx.arr, ctx.n, ctx.dims = arr, len(draws), (M, B, Dtot, ne, npar, nkl)
        ctx.prior_shape = None if mu0 is None else (mu0.shape, s0.shape)
        ctx.s_raw = bool(s_raw)
        if scales is not None:
            ctx.mark_non_differentiable(scales)
        return z, ploc, pscale, kl, scales

    @staticmethod
    def backward(ctx, dz, dploc, dpscale, dkl, _dscales=None):
        mu_c, s_c, mu0_c, s0_c, eps_c, row_masks = ctx.saved_tensors
        M, B, Dtot, ne, npar, nkl = ctx.dims
        dev = mu_c.device
        f = lambda t: None if t is None else t.detach().float().contiguous()
        dz, dkl, dploc, dpscale = f(dz), f(dkl), f(dploc), f(dpscale)
        if (dploc is None) != (dpscale is None):
            ref = dploc if dploc is not None else dpscale
            dploc = torch.zeros_like(ref) if dploc is None else dploc
            dpscale = torch.zeros_like(ref) if dpscale is None else dpscale
        if ne == 0:
            dz = None
        if npar == 0:
            dploc = dpscale = None
        if nkl == 0:
            dkl = None
        dmu = torch.empty_like(mu_c)
        ds = torch.empty_like(s_c)
        nws = _lib.load().mmvae_latent_draws_bwd_ws_floats(B, Dtot)
        ws = torch.empty(nws, dtype=torch.float32, device=dev)
        dprior = torch.empty(2, Dtot, dtype=torch.float32, device=dev)  # fully written by the finalisation kernel
        call("mmvae_latent_draws_bwd_tail", _ptr(mu_c), _ptr(s_c), M, B, Dtot, ctx.arr, ctx.n, _ptr(row_masks),
             _ptr(mu0_c), _ptr(s0_c), _ptr(eps_c), _ptr(dz), _ptr(dkl), _ptr(dploc), _ptr(dpscale), int(ctx.s_raw),
             _ptr(dmu), _ptr(ds), _ptr(ws), _ptr(dprior[0]), _ptr(dprior[1]), _stream())
        gm0 = gs0 = None
        if ctx.prior_shape is not None:
            D = mu0_c.numel()
            gm0 = dprior[0, :D].reshape(ctx.prior_shape[0])
            gs0 = dprior[1, :D].reshape(ctx.prior_shape[1])
        return dmu, ds, gm0, gs0, None, None, None, None


class DrawResults(list):
    """Per-draw result dicts of latent_draws; ``kl_packed`` is the kernel's packed KL output (what elbo_combine reads)."""
    kl_packed = None
    scales = None


def latent_draws(mu, s, mu0, s0, eps, draws: List[Draw], row_masks=None, s_raw=False):
    """Run a list of draws.  Returns per-draw dicts with views: z (K,B,w) | None, loc/scale (B,w) | None, kl (B) | None.
    s_raw=True: `s` holds the raw logits of the encoders' second head and the kernels apply the encoder tail
    softmax(., -1) + 1e-6 themselves (reference encoders.py:49-54); the result's ``scales`` is the (M,B,Dtot) tensor of
    the scales they computed (no gradient), and the gradient returned for `s` is the one of the raw logits."""
    M, B, Dtot = mu.shape
    z, ploc, pscale, kl, scales = _LatentDraws.apply(mu, s, mu0, s0, eps, draws, row_masks, s_raw)
    out = DrawResults()
    out.scales = scales
    out.kl_packed = kl if any(d.kl_mode for d in draws) else None  # (n_kl * B): rows of the draws with a KL, in order
    eo = po = ko = 0
    for d in draws:
        e = {"z": None, "loc": None, "scale": None, "kl": None}
        n = d.K * B * d.width
        if d.K:
            e["z"] = z[eo:eo + n].view(d.K, B, d.width)
        eo += n
        if d.want_params:
            e["loc"] = ploc[po:po + B * d.width].view(B, d.width)
            e["scale"] = pscale[po:po + B * d.width].view(B, d.width)
            po += B * d.width
        if d.kl_mode:
            e["kl"] = kl[ko:ko + B]
            ko += B
        out.append(e)
    return out


# ----------------------------------------------------------------------------------------------------------
# MoE sample + log-densities
# ----------------------------------------------------------------------------------------------------------
class _MoeLogdens(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_moe_logdens_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, mu, s, mu0, s0, eps, dists, through_z, s_raw=False):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(mu, s, mu0, s0, eps)
        M, B, D = mu.shape
        K = eps.shape[1]
        mu_c, s_c = mu.detach().float().contiguous(), s.detach().float().contiguous()
        mu0_c, s0_c = mu0.detach().float().contiguous().reshape(-1), s0.detach().float().contiguous().reshape(-1)
        eps_c = eps.detach().float().contiguous()
        dev = mu.device
        z = torch.empty((M, K, B, D), dtype=torch.float32, device=dev)
        lq = torch.empty((M, M, K, B), dtype=torch.float32, device=dev)
        lpz = torch.empty((M, K, B), dtype=torch.float32, device=dev)
        darr = (ctypes.c_int32 * M)(*[int(x) for x in dists])
        scales = None
        if s_raw:  # encoder tail softmax(., -1) + 1e-6 evaluated inside the kernel (flat kernels only)
            if not moe_rk_supported(M, D):
                raise RuntimeError("mmvae_b200: the fused encoder tail of moe_logdens needs D % 4 == 0, D <= 128, M <= 3")
            scales = torch.empty_like(s_c)
            call("mmvae_moe_logdens_fwd_tail", _ptr(mu_c), _ptr(s_c), M, B, D, K, darr, _ptr(mu0_c), _ptr(s0_c),
                 _ptr(eps_c), _ptr(z), _ptr(lq), _ptr(lpz), _ptr(scales), _stream())
            ctx.mark_non_differentiable(scales)
        else:
            call("mmvae_moe_logdens_fwd", _ptr(mu_c), _ptr(s_c), M, B, D, K, darr, _ptr(mu0_c), _ptr(s0_c), _ptr(eps_c),
                 _ptr(z), _ptr(lq), _ptr(lpz), _stream())
        ctx.save_for_backward(mu_c, s_c, mu0_c, s0_c, eps_c)
        ctx.meta = (M, B, D, K, darr, int(through_z), mu0.shape, s0.shape)
        ctx.s_raw = bool(s_raw)
        return z, lq, lpz, scales

    @staticmethod
    def backward(ctx, dz, dlq, dlpz, _dscales=None):
        mu_c, s_c, mu0_c, s0_c, eps_c = ctx.saved_tensors
        M, B, D, K, darr, through_z, sh0, sh1 = ctx.meta
        f = lambda t: None if t is None else t.detach().float().contiguous()
        dz, dlq, dlpz = f(dz), f(dlq), f(dlpz)
        dev = mu_c.device
        # "rk" hand-over from a DReG combine that consumed lq / lpz (ops._DregRows.backward): its gradients are
        # c[r,k] * softmax_j(lq) and -c[r,k] with c constant over b, so instead of materialising (M,M,K,B) / (M,K,B)
        # tensors it leaves (wt, g, 1/M, softmax_j) here and the kernel folds them into its coefficients
        rk, ctx.rk = getattr(ctx, "rk", None), None
        rk_w = rk_g = None
        rk_mul = 1.0
        packed = 0
        if rk is not None:
            wt, g, mul, soft, soft_packed = rk
            if dlq is None and dlpz is None and moe_rk_supported(M, D):
                rk_w, rk_g, rk_mul, dlq, packed = wt, g, mul, soft, int(soft_packed)
                if packed and dz is None:  # the packed layout exists in the training-step kernel variant only
                    dz = torch.zeros((M, K, B, D), dtype=torch.float32, device=dev)
            else:  # lq / lpz also feed something else (or an unsupported shape): materialise and add
                if soft_packed:  # (K, B, M*M) [r*M + j] -> (M, M, K, B)
                    soft = soft.view(K, B, M, M).permute(2, 3, 0, 1)
                c = wt * mul if g is None else wt * (g * mul)
                d1, d2 = c.view(M, 1, K, 1) * soft, (-c).view(M, K, 1).expand(M, K, B)
                dlq = d1.contiguous() if dlq is None else dlq + d1
                dlpz = d2.contiguous() if dlpz is None else dlpz + d2
        dmu, ds = torch.empty_like(mu_c), torch.empty_like(s_c)
        nws = _lib.load().mmvae_moe_logdens_bwd_ws_floats(B, D, K)
        ws = torch.empty(nws, dtype=torch.float32, device=dev)
        dprior = torch.empty(2, D, dtype=torch.float32, device=dev)
        if ctx.s_raw:
            call("mmvae_moe_logdens_bwd_tail", _ptr(mu_c), _ptr(s_c), M, B, D, K, darr, _ptr(mu0_c), _ptr(s0_c),
                 _ptr(eps_c), _ptr(dz), _ptr(dlq), _ptr(dlpz), through_z, _ptr(rk_w), _ptr(rk_g), float(rk_mul), packed, 1,
                 _ptr(dmu), _ptr(ds), _ptr(ws), _ptr(dprior[0]), _ptr(dprior[1]), _stream())
        else:
            call("mmvae_moe_logdens_bwd_rk", _ptr(mu_c), _ptr(s_c), M, B, D, K, darr, _ptr(mu0_c), _ptr(s0_c), _ptr(eps_c),
                 _ptr(dz), _ptr(dlq), _ptr(dlpz), through_z, _ptr(rk_w), _ptr(rk_g), float(rk_mul), packed, _ptr(dmu),
                 _ptr(ds), _ptr(ws), _ptr(dprior[0]), _ptr(dprior[1]), _stream())
        return dmu, ds, dprior[0].reshape(sh0), dprior[1].reshape(sh1), None, None, None, None


def moe_rk_supported(M, D):
    """Shapes the flat MoE kernels (and with them the folded DReG coefficients) cover: include/mmvae_b200.h
    mmvae_moe_logdens_bwd_rk."""
    return D % 4 == 0 and D <= 128 and M <= 3


def moe_logdens(mu, s, mu0, s0, eps, dists, through_z=True):
    return _MoeLogdens.apply(mu, s, mu0, s0, eps, tuple(dists), through_z, False)[:3]


def moe_logdens_tail(mu, s_raw, mu0, s0, eps, dists, through_z=True):
    """moe_logdens with the encoder tail fused in: `s_raw` holds the raw logits of the encoders' second head, the kernel
    applies softmax(., -1) + 1e-6 (reference encoders.py:49-54).  Returns (z, lq, lpz, scales); the gradient that comes
    back for `s_raw` is the one of the raw logits.  Shapes: moe_rk_supported(M, D)."""
    return _MoeLogdens.apply(mu, s_raw, mu0, s0, eps, tuple(dists), through_z, True)


# ----------------------------------------------------------------------------------------------------------
# objective combination
# ----------------------------------------------------------------------------------------------------------
class _Iwae(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_objective_iwae: loss + the softmax weights its backward needs in one launch."""

    @staticmethod
    def forward(ctx, lpz, lq, lpx, beta):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(lpz, lq, lpx)
        M, K, B = lpz.shape
        L = lpx.shape[1]
        f = lambda t: t.detach().float().contiguous()
        lpz_c, lq_c, lpx_c = f(lpz), f(lq), f(lpx)
        dev = lpz.device
        lw = torch.empty((M * K, B), dtype=torch.float32, device=dev)
        loss_b = torch.empty(B, dtype=torch.float32, device=dev)
        w = torch.empty((M, K, B), dtype=torch.float32, device=dev)
        dlq = torch.empty((M, M, K, B), dtype=torch.float32, device=dev)
        call("mmvae_objective_iwae", _ptr(lpz_c), _ptr(lq_c), _ptr(lpx_c), M, L, K, B, float(beta), _ptr(lw),
             _ptr(loss_b), _ptr(w), _ptr(dlq), _stream())
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("mmvae_reduce_sum", _ptr(loss_b), B, 1.0, _ptr(loss), _stream())
        ctx.save_for_backward(w, dlq)
        ctx.L = L
        ctx.mark_non_differentiable(lw)
        return loss, lw

    @staticmethod
    def backward(ctx, g, _glw):
        if g is None:
            return None, None, None, None
        w, dlq = ctx.saved_tensors
        gw = g * w  # (M,K,B) -- tiny
        return -gw, g * dlq, (-gw).unsqueeze(1).expand(-1, ctx.L, -1, -1), None


def iwae_combine(lpz, lq, lpx, beta):
    return _Iwae.apply(lpz, lq, lpx, float(beta))


class _IwaeRows(torch.autograd.Function):
    """IWAE combine over a LIST of likelihood row vectors (index r*L + l, (K*B,) each) read through a pointer table
    (no stack copy); the backward is one launch and hands every row vector of modality r the same gradient -g*w[r]."""

    @staticmethod
    def forward(ctx, lpz, lq, beta, L, ticket, *rows):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(lpz, lq, *rows)
        M, K, B = lpz.shape
        f = lambda t: t.detach().float().contiguous()
        lpz_c, lq_c = f(lpz), f(lq)
        rows_c = [f(r).reshape(-1) for r in rows]
        if len(rows_c) != M * L or any(r.numel() != K * B for r in rows_c):
            raise RuntimeError("mmvae_b200: iwae needs M*L row vectors of K*B elements")
        ptrs = (ctypes.c_void_p * (M * L))(*[r.data_ptr() for r in rows_c])
        dev = lpz.device
        lw = torch.empty((M * K, B), dtype=torch.float32, device=dev)
        loss_b = torch.empty(B, dtype=torch.float32, device=dev)
        w = torch.empty((M, K, B), dtype=torch.float32, device=dev)
        dlq = torch.empty((M, M, K, B), dtype=torch.float32, device=dev)
        nw = torch.empty((M, K, B), dtype=torch.float32, device=dev)  # -w: the gradients for a unit upstream gradient
        loss = torch.empty((), dtype=torch.float32, device=dev)
        # one launch: log-weights, per-sample loss, softmax weights, d/dlq, -w and the batch sum of the loss
        call("mmvae_objective_iwae_fused", _ptr(lpz_c), _ptr(lq_c), _P(0), ptrs, M, L, K, B, float(beta), _ptr(lw),
             _ptr(loss_b), _ptr(w), _ptr(dlq), _ptr(nw), _ptr(loss), _ptr(ticket if ticket is not None else _ticket(dev)),
             _stream())
        ctx.save_for_backward(w, dlq)
        ctx.nw = nw
        ctx.meta = (M, L, K, B, [r.shape for r in rows])
        ctx.mark_non_differentiable(lw)
        return loss, lw

    @staticmethod
    def backward(ctx, g, _glw):
        w, dlq = ctx.saved_tensors
        M, L, K, B, shapes = ctx.meta
        if g is None:
            return (None, None, None, None, None) + (None,) * (M * L)
        dlq_out = dlq  # scaled in place by g (single-use buffer, like the fused ELBO gradient)
        if _is_unit(g):  # registered unit gradient: the forward already produced every gradient, nothing to launch
            if ctx.nw is None:
                raise RuntimeError("mmvae_b200: the IWAE gradient buffers are single-use (retain_graph unsupported)")
            dlpz, ctx.nw = ctx.nw, None
        else:
            # the kernel scales the saved dlq IN PLACE through a raw pointer (autograd sees no version bump): a second
            # backward would silently return dlq * g^2 -- single-use like the unit path and the fused ELBO buffers
            if getattr(ctx, "dlq_consumed", False):
                raise RuntimeError("mmvae_b200: the IWAE gradient buffers are single-use (retain_graph unsupported)")
            ctx.dlq_consumed = True
            gs = g.detach().float().contiguous()
            dlpz = torch.empty_like(w)
            call("mmvae_objective_iwae_bwd", _ptr(gs), _ptr(w), _ptr(dlq_out), _ptr(dlpz), w.numel(), dlq_out.numel(),
                 _stream())
        row_grads = [dlpz[i // L].reshape(shapes[i]) for i in range(M * L)]
        return (dlpz, dlq_out, None, None, None) + tuple(row_grads)


def iwae_combine_rows(lpz, lq, rows, L, beta, ticket=None):
    """rows: list of M*L tensors (K*B,), index r*L + l.  ticket: optional counter from new_ticket() owned by the caller."""
    return _IwaeRows.apply(lpz, lq, float(beta), int(L), ticket, *rows)


class _PriorScale(torch.autograd.Function):
    """s0 = softmax(logits, 1) * D of the learnable prior (reference mmvae_models.py:28-30), one tiny kernel each way."""

    @staticmethod
    def forward(ctx, logits, peer):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(logits)
        lg = logits.detach().float().contiguous()
        D = lg.shape[-1]
        s0 = torch.empty_like(lg)
        call("mmvae_prior_scale_fwd", _ptr(lg), D, _ptr(s0), _stream())
        ctx.save_for_backward(s0)
        ctx.peer = peer
        return s0

    @staticmethod
    def backward(ctx, ds0):
        if ds0 is None:
            return None, None
        (s0,) = ctx.saved_tensors
        d = ds0.detach().float().contiguous()
        out = torch.empty_like(s0)
        if ctx.peer is not None:  # gradient sync of the replicated logits fused into this kernel (peer memory)
            pg = ctx.peer
            call("mmvae_prior_scale_bwd_peer", _ptr(s0), _ptr(d), s0.shape[-1], _ptr(out), pg.bufs_dev, pg.rank, pg.world,
                 pg.CH_PRIOR_GRAD, _stream())
        else:
            call("mmvae_prior_scale_bwd", _ptr(s0), _ptr(d), s0.shape[-1], _ptr(out), _stream())
        return out, None


def prior_scale(logits, peer=None):
    """s0 = softmax(logits, 1) * D.  peer: a parallel.PeerGroup -- the backward then returns the gradient already
    all-reduced (SUM) over its ranks (one fused kernel, no NCCL); every rank must run this backward once per step."""
    return _PriorScale.apply(logits, peer)


class _Dreg(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_objective_dreg_stage{1,2} (parity mode: batch-summed log-weights).  `group`: the
    process group of a batch-sharded run -- the (M,K) partial sums are all-reduced between the stages."""

    @staticmethod
    def forward(ctx, lpz, lq, lpx, group):
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not as zero tensors
        _need_cuda(lpz, lq, lpx)
        M, K, B = lpz.shape
        L = lpx.shape[1]
        f = lambda t: t.detach().float().contiguous()
        lpz_c, lq_c, lpx_c = f(lpz), f(lq), f(lpx)
        dev = lpz.device
        part = torch.empty(((_lib.DREG_MAX_SPLIT + 1), M * K), dtype=torch.float64, device=dev)
        lq_soft = torch.empty((M, M, K, B), dtype=torch.float32, device=dev)
        call("mmvae_objective_dreg_stage1", _ptr(lpz_c), _ptr(lq_c), _ptr(lpx_c), M, L, K, B, _ptr(part),
             _ptr(lq_soft), _stream())
        lw = part[0]
        _, pgroup = _peer(group)
        if pgroup is not None:
            import torch.distributed as dist
            dist.all_reduce(lw, group=pgroup)
        wt = torch.empty((M, K), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("mmvae_objective_dreg_stage2", _ptr(lw), M, K, _ptr(wt), _ptr(loss), _stream())
        ctx.save_for_backward(wt, lq_soft)
        ctx.meta = (M, L, K, B)
        lw_out = lw.clone().view(M, K)
        ctx.mark_non_differentiable(lw_out)
        return loss, lw_out

    @staticmethod
    def backward(ctx, g, _glw):
        if g is None:
            return None, None, None, None
        wt, lq_soft = ctx.saved_tensors
        M, L, K, B = ctx.meta
        c = (g / M) * wt  # (M,K): -dloss/dlw
        d_rows = (-c).unsqueeze(-1).expand(M, K, B)
        return d_rows, c.view(M, 1, K, 1) * lq_soft, d_rows.unsqueeze(1).expand(M, L, K, B), None


def dreg_combine(lpz, lq, lpx, group=None):
    return _Dreg.apply(lpz, lq, lpx, group)


class _DregRows(torch.autograd.Function):
    """DReG combine over a LIST of likelihood row vectors (index r*L + l, (K*B,) each) read through a pointer table (no
    stack copy).  Backward: one launch writes the (M,L,K,B) row gradients -(g/M) wt[r,k]; the gradients of lq / lpz
    are not materialised -- when `node` (the backward node of the ops.moe_logdens call that produced lq and lpz) is
    given, (wt, g, 1/M, softmax_j lq) is handed to it and folded into the coefficients of its kernel."""

    @staticmethod
    def forward(ctx, lpz, lq, L, group, node, *rows):
        ctx.set_materialize_grads(False)
        _need_cuda(lpz, lq, *rows)
        M, K, B = lpz.shape
        f = lambda t: t.detach().float().contiguous()
        lpz_c, lq_c = f(lpz), f(lq)
        rows_c = [f(r).reshape(-1) for r in rows]
        if len(rows_c) != M * L or any(r.numel() != K * B for r in rows_c):
            raise RuntimeError("mmvae_b200: dreg needs M*L row vectors of K*B elements")
        ptrs = (ctypes.c_void_p * (M * L))(*[r.data_ptr() for r in rows_c])
        dev = lpz.device
        part = torch.empty(((_lib.DREG_MAX_SPLIT + 1), M * K), dtype=torch.float64, device=dev)
        # softmax_j(lq) for the backward.  When the MoE backward kernel will fold the log q gradients in (node given,
        # M == 2, one posterior family, gradient through z), it is written as (K, B, 4) vectors: that kernel then needs
        # one 16-byte copy per (k, b) instead of four scattered 4-byte ones
        soft_packed = False
        if node is not None and M == 2:
            nM, nB, nD, nK, ndarr, nthru = node.meta[:6]
            soft_packed = bool(nthru) and moe_rk_supported(nM, nD) and ndarr[0] == ndarr[1]
        lq_soft = torch.empty((K, B, M * M) if soft_packed else (M, M, K, B), dtype=torch.float32, device=dev)
        call("mmvae_objective_dreg_stage1_ptrs", _ptr(lpz_c), _ptr(lq_c), _P(0), ptrs, M, L, K, B, _ptr(part),
             _ptr(lq_soft), int(soft_packed), _stream())
        lw = part[0]
        peer, pgroup = _peer(group)
        wt = torch.empty((M, K), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        if peer is not None and M * K * 8 <= 4096:
            # the (M,K) batch sums are global -- exchanged through peer memory INSIDE stage 2
            call("mmvae_objective_dreg_stage2_peer", _ptr(lw), M, K, _ptr(wt), _ptr(loss), peer.bufs_dev, peer.rank,
                 peer.world, peer.CH_DREG, _stream())
        else:
            if pgroup is not None:  # NCCL on the current stream (capturable)
                import torch.distributed as dist
                dist.all_reduce(lw, group=pgroup)
            call("mmvae_objective_dreg_stage2", _ptr(lw), M, K, _ptr(wt), _ptr(loss), _stream())
        ctx.save_for_backward(wt, lq_soft)
        ctx.meta = (M, L, K, B, [r.shape for r in rows])
        ctx.node = node
        ctx.soft_packed = soft_packed
        lw_out = lw.clone().view(M, K)
        ctx.mark_non_differentiable(lw_out)
        return loss, lw_out

    @staticmethod
    def backward(ctx, g, _glw):
        M, L, K, B, shapes = ctx.meta
        if g is None:
            return (None,) * (5 + M * L)
        wt, lq_soft = ctx.saved_tensors
        gs = None if _is_unit(g) else g.detach().float().contiguous()
        d_rows = torch.empty((M, L, K, B), dtype=torch.float32, device=wt.device)
        call("mmvae_objective_dreg_rowgrads", _ptr(gs), _ptr(wt), M, L, K, B, _ptr(d_rows), _stream())
        row_grads = tuple(d_rows[i // L, i % L].reshape(shapes[i]) for i in range(M * L))
        if ctx.node is not None:
            ctx.node.rk = (wt, gs, 1.0 / M, lq_soft, ctx.soft_packed)
            return (None, None, None, None, None) + row_grads
        c = wt / M if gs is None else wt * (gs / M)
        return (((-c).view(M, K, 1).expand(M, K, B)), c.view(M, 1, K, 1) * lq_soft, None, None, None) + row_grads


def dreg_combine_rows(lpz, lq, rows, L, group=None):
    """rows: list of M*L tensors (K*B,), index r*L + l.  If lq and lpz are the direct outputs of one ops.moe_logdens
    call, their gradients are folded into that call's backward kernel (no (M,M,K,B) temporaries)."""
    node = lq.grad_fn
    if node is None or node is not lpz.grad_fn or not isinstance(node, _MoeLogdens._backward_cls):
        node = None
    return _DregRows.apply(lpz, lq, int(L), group, node, *rows)


class _ElboCombine(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_objective_elbo: loss = sum_i coef_i * sum(term_i) + sum_j kl_coef_j * sum_b kl[j,b]
    in one launch.  Differentiable inputs: the packed KL rows and the term scalars S_i (placeholders when their sum was
    deferred to this launch); `vecs` are the vectors actually read (row vectors or the S_i themselves)."""

    @staticmethod
    def forward(ctx, kl, kl_coef, kl_log_coef, coefs, vecs, *S):
        ctx.set_materialize_grads(False)
        _need_cuda(kl, *vecs)
        n_t, n_kl = len(vecs), len(kl_coef)
        if n_t > _lib.ELBO_MAX_TERMS or n_kl > _lib.ELBO_MAX_TERMS:
            raise RuntimeError("mmvae_b200: elbo_combine takes at most %d terms / KL segments" % _lib.ELBO_MAX_TERMS)
        dev = (kl if kl is not None else vecs[0]).device
        B = 0
        klc = None
        if n_kl:
            klc = kl.detach().float().contiguous().reshape(-1)
            if klc.numel() % n_kl:
                raise RuntimeError("mmvae_b200: packed KL rows (%d) do not split into %d segments" % (klc.numel(), n_kl))
            B = klc.numel() // n_kl
        vs = [v.detach().float().contiguous().reshape(-1) for v in vecs]
        ptrs = (ctypes.c_void_p * max(n_t, 1))(*[v.data_ptr() for v in vs])
        ns = (ctypes.c_int64 * max(n_t, 1))(*[v.numel() for v in vs])
        cf = (ctypes.c_float * max(n_t, 1))(*[float(c) for c in coefs])
        kc = (ctypes.c_float * max(n_kl, 1))(*[float(c) for c in kl_coef])
        kg = (ctypes.c_float * max(n_kl, 1))(*[float(c) for c in (kl_log_coef or [0.0] * n_kl)])
        loss = torch.empty((), dtype=torch.float32, device=dev)
        kld = torch.empty((), dtype=torch.float32, device=dev)  # the logged "kld"
        dkl = torch.empty(max(n_kl * B, 1), dtype=torch.float32, device=dev) if n_kl else None
        ws = torch.empty(2 * _lib.ELBO_MAX_CTAS, dtype=torch.float32, device=dev)  # per-CTA partials
        call("mmvae_objective_elbo", ptrs, ns, cf, n_t, _ptr(klc), B, kc, kg, n_kl, _ptr(loss), _ptr(kld), _ptr(dkl),
             _ptr(ws), _ptr(_ticket(dev)), _stream())
        ctx.dkl = dkl
        ctx.n_S = len(S)
        ctx.kl_shape = None if kl is None else kl.shape
        ctx.mark_non_differentiable(kld)
        return loss, kld

    @staticmethod
    def backward(ctx, g, _gk):
        if g is None:
            return (None,) * (5 + ctx.n_S)
        dkl = None
        if ctx.dkl is not None:
            # unit upstream gradient (registered static ones tensor): the forward already wrote the KL-row gradients
            dkl = ctx.dkl if _is_unit(g) else ctx.dkl * g
            dkl = dkl.view(ctx.kl_shape)
        return (dkl, None, None, None, None) + (g,) * ctx.n_S  # every S_i enters with coefficient 1


def elbo_combine(terms, kl=None, kl_coef=(), kl_log_coef=None):
    """loss, kld = sum of the likelihood terms + sum_j kl_coef[j] * sum_b kl[j, b]  (kld: same with kl_log_coef).
    terms: 0-d tensors S_i from loglik_weighted_sum / catce_weighted_sum (defer=True: their batch sum happens here) or
    any other differentiable 0-d tensor (coefficient 1).  kl: packed (n_kl * B) KL rows (latent_draws(...).kl_packed)."""
    vecs, coefs = [], []
    for S in terms:
        d = getattr(S, "_mmvae_deferred", None)
        if d is not None:
            vecs.append(d[0])
            coefs.append(d[1])
        else:
            vecs.append(S)
            coefs.append(1.0)
    return _ElboCombine.apply(kl, tuple(kl_coef), None if kl_log_coef is None else tuple(kl_log_coef), tuple(coefs),
                              tuple(vecs), *terms)


class _KlElementwise(torch.autograd.Function):
    """include/mmvae_b200.h mmvae_kl_elementwise_{fwd,bwd}: (n, D) KL(q || N(loc0, scale0)), prior broadcast over n."""

    @staticmethod
    def forward(ctx, loc, scale, loc0, scale0, dist_code):
        ctx.set_materialize_grads(False)
        _need_cuda(loc, scale, loc0, scale0)
        D = loc.shape[-1]
        f = lambda t: t.detach().float().contiguous()
        l, sg = f(loc).reshape(-1, D), f(scale).reshape(-1, D)
        l0 = f(loc0).reshape(-1).expand(D).contiguous() if loc0.numel() == 1 else f(loc0).reshape(-1)
        s0 = f(scale0).reshape(-1).expand(D).contiguous() if scale0.numel() == 1 else f(scale0).reshape(-1)
        if l0.numel() != D or s0.numel() != D:
            raise RuntimeError("mmvae_b200: the prior must broadcast as a single (D) row")
        out = torch.empty_like(l)
        call("mmvae_kl_elementwise_fwd", _ptr(l), _ptr(sg), _ptr(l0), _ptr(s0), int(dist_code), l.shape[0], D, _ptr(out),
             _stream())
        ctx.save_for_backward(l, sg, l0, s0)
        ctx.meta = (int(dist_code), loc.shape, scale.shape, loc0.shape, scale0.shape)
        return out.view(loc.shape)

    @staticmethod
    def backward(ctx, up):
        if up is None:
            return None, None, None, None, None
        l, sg, l0, s0 = ctx.saved_tensors
        code, sh_l, sh_s, sh_l0, sh_s0 = ctx.meta
        n, D = l.shape
        u = up.detach().float().contiguous().reshape(n, D)
        dl, ds = torch.empty_like(l), torch.empty_like(sg)
        ws = torch.empty(_lib.load().mmvae_kl_elementwise_ws_floats(n, D), dtype=torch.float32, device=l.device)
        dp = torch.empty(2, D, dtype=torch.float32, device=l.device)
        call("mmvae_kl_elementwise_bwd", _ptr(l), _ptr(sg), _ptr(l0), _ptr(s0), code, n, D, _ptr(u), _ptr(dl), _ptr(ds),
             _ptr(ws), _ptr(dp[0]), _ptr(dp[1]), _stream())
        red = lambda g, shp: g.sum().reshape(shp) if len(shp) == 0 or int(torch.tensor(shp).prod()) == 1 else g.reshape(shp)
        return dl.view(sh_l), ds.view(sh_s), red(dp[0], sh_l0), red(dp[1], sh_s0), None


def kl_elementwise(loc, scale, loc0, scale0, laplace=False):
    return _KlElementwise.apply(loc, scale, loc0, scale0, 1 if laplace else 0)


def kl_table(loc, scale, loc0, scale0, laplace=False):
    """include/mmvae_b200.h mmvae_kl_table (forward only: analysis).  loc, scale (M, n, D); prior (D) ->
    (M + M(M-1)/2, n, D): KL(q_i || prior) rows, then the symmetric J divergences of the pairs i < j."""
    _need_cuda(loc, scale, loc0, scale0)
    M, n, D = loc.shape
    f = lambda t: t.detach().float().contiguous()
    l, sg = f(loc), f(scale)
    l0 = f(loc0).reshape(-1).expand(D).contiguous() if loc0.numel() == 1 else f(loc0).reshape(-1)
    s0 = f(scale0).reshape(-1).expand(D).contiguous() if scale0.numel() == 1 else f(scale0).reshape(-1)
    if l0.numel() != D or s0.numel() != D:
        raise RuntimeError("mmvae_b200: the prior must broadcast as a single (D) row")
    out = torch.empty((M + M * (M - 1) // 2, n, D), dtype=torch.float32, device=loc.device)
    darr = (ctypes.c_int32 * M)(*([1 if laplace else 0] * M))
    call("mmvae_kl_table", _ptr(l), _ptr(sg), M, darr, _ptr(l0), _ptr(s0), n, D, _ptr(out), _stream())
    return out


def reduce_sum(x, scale=1.0):
    """Deterministic single-CTA sum (forward only helper for logging values)."""
    _need_cuda(x)
    xc = x.detach().float().contiguous().reshape(-1)
    out = torch.empty((), dtype=torch.float32, device=x.device)
    call("mmvae_reduce_sum", _ptr(xc), xc.numel(), float(scale), _ptr(out), _stream())
    return out
