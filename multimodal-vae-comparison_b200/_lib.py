"""ctypes binding of libmmvae_b200.so (the C ABI declared in include/mmvae_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MMVAE_B200_LIB: alternative build of the same library (kernel tuning variants, tools/tune_*.sh); never a fallback
LIB_PATH = os.environ.get("MMVAE_B200_LIB") or os.path.join(HERE, "libmmvae_b200.so")

c_i, c_i64, c_f, c_d, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p

F32, BF16 = 0, 1
NORMAL, LAPLACE = 0, 1
LT_BCE, LT_LPROB_NORMAL, LT_LPROB_LAPLACE, LT_MSE, LT_L1, LT_BCE_LOGITS = 0, 1, 2, 3, 4, 5
LT_LPROB_NORMAL_SELF, LT_LPROB_LAPLACE_SELF = 6, 7
DRAW_PRIOR, DRAW_DIRECT, DRAW_LAPLACE, DRAW_ROWMASK = 1, 2, 4, 8
MAX_MODS, MAX_COLS, MAX_DRAWS, DREG_MAX_SPLIT = 8, 256, 64, 64
PEER_CHANNELS, PEER_MAX_WORLD, PEER_BUFFER_BYTES = 8, 32, 72 * 1024
ELBO_MAX_TERMS, ELBO_MAX_CTAS = 48, 64


class DrawDesc(ctypes.Structure):
    """mmvae_draw_desc (include/mmvae_b200.h)."""
    _fields_ = [("mask", ctypes.c_uint32), ("flags", ctypes.c_int32), ("kl_mode", ctypes.c_int32),
                ("col0", ctypes.c_int32), ("width", ctypes.c_int32), ("K", ctypes.c_int32),
                ("eps_off", c_i64), ("z_off", c_i64), ("par_off", c_i64), ("kl_off", c_i64)]


_RECON = [c_p, c_i64, c_i, c_p, c_i64, c_i]  # recon, ld, dtype, target, ld, dtype
SIGNATURES = {
    "mmvae_version": (c_i, []),
    "mmvae_loglik_workspace_bytes": (c_i64, [c_i64, c_i64, c_i]),
    "mmvae_loglik_rowreduce_fwd": (c_i, _RECON + [c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_p]),
    "mmvae_loglik_rowreduce_bwd": (c_i, _RECON + [c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_i64, c_p]),
    "mmvae_loglik_rowreduce_fused": (c_i, _RECON + [c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_f, c_p, c_p, c_i64,
                                                    c_p, c_p]),
    "mmvae_catce_rows": (c_i, [c_i] + _RECON + [c_i64, c_i64, c_i64, c_i64, c_f, c_p, c_f, c_p, c_p, c_i64, c_p, c_p]),
    "mmvae_catce_rows_masked": (c_i, [c_i] + _RECON + [c_i64, c_i64, c_i64, c_i64, c_f, c_p, c_f, c_p, c_p, c_i64, c_p, c_p,
                                             c_i64, c_p]),
    "mmvae_osigma_sumsq": (c_i, _RECON + [c_i64, c_i64, c_i64, c_p, c_p, c_p]),
    "mmvae_osigma_fwd": (c_i, _RECON + [c_i64, c_i64, c_i64, c_f, c_p, c_d, c_p, c_p, c_p, c_p]),
    "mmvae_osigma_bwd": (c_i, _RECON + [c_i64, c_i64, c_i64, c_f, c_p, c_d, c_p, c_p, c_p, c_i64, c_p]),
    "mmvae_latent_draws_fwd": (c_i, [c_p, c_p, c_i, c_i64, c_i, ctypes.POINTER(DrawDesc), c_i, c_p, c_p, c_p, c_p,
                                     c_p, c_p, c_p, c_p, c_p]),
    "mmvae_latent_draws_bwd_ws_floats": (c_i64, [c_i64, c_i]),
    "mmvae_latent_draws_bwd": (c_i, [c_p, c_p, c_i, c_i64, c_i, ctypes.POINTER(DrawDesc), c_i, c_p, c_p, c_p, c_p,
                                     c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_latent_draws_fwd_tail": (c_i, [c_p, c_p, c_i, c_i64, c_i, ctypes.POINTER(DrawDesc), c_i, c_p, c_p, c_p, c_p,
                                          c_p, c_p, c_p, c_p, c_i, c_p, c_p]),
    "mmvae_latent_draws_bwd_tail": (c_i, [c_p, c_p, c_i, c_i64, c_i, ctypes.POINTER(DrawDesc), c_i, c_p, c_p, c_p, c_p,
                                          c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_moe_logdens_fwd_tail": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p,
                                         c_p, c_p, c_p, c_p, c_p]),
    "mmvae_moe_logdens_bwd_tail": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p,
                                         c_p, c_p, c_p, c_i, c_p, c_p, c_f, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_kl_elementwise_fwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i64, c_i, c_p, c_p]),
    "mmvae_kl_elementwise_ws_floats": (c_i64, [c_i64, c_i]),
    "mmvae_kl_elementwise_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i64, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_kl_table": (c_i, [c_p, c_p, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_i64, c_i, c_p, c_p]),
    "mmvae_moe_logdens_fwd": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p,
                                    c_p, c_p, c_p, c_p]),
    "mmvae_moe_logdens_bwd_ws_floats": (c_i64, [c_i64, c_i, c_i]),
    "mmvae_moe_logdens_bwd": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p,
                                    c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_moe_logdens_bwd_rk": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p,
                                       c_p, c_p, c_p, c_i, c_p, c_p, c_f, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_objective_iwae": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i64, c_f, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_objective_iwae_ptrs": (c_i, [c_p, c_p, c_p, ctypes.POINTER(c_p), c_i, c_i, c_i, c_i64, c_f, c_p, c_p, c_p,
                                        c_p, c_p]),
    "mmvae_objective_iwae_fused": (c_i, [c_p, c_p, c_p, ctypes.POINTER(c_p), c_i, c_i, c_i, c_i64, c_f, c_p, c_p, c_p,
                                         c_p, c_p, c_p, c_p, c_p]),
    "mmvae_objective_iwae_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p]),
    "mmvae_prior_scale_fwd": (c_i, [c_p, c_i, c_p, c_p]),
    "mmvae_prior_scale_bwd": (c_i, [c_p, c_p, c_i, c_p, c_p]),
    "mmvae_objective_dreg_stage1": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_p]),
    "mmvae_objective_dreg_stage1_ptrs": (c_i, [c_p, c_p, c_p, ctypes.POINTER(c_p), c_i, c_i, c_i, c_i64, c_p, c_p, c_i,
                                               c_p]),
    "mmvae_objective_dreg_stage2": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p]),
    "mmvae_objective_dreg_rowgrads": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i64, c_p, c_p]),
    "mmvae_reduce_sum": (c_i, [c_p, c_i64, c_f, c_p, c_p]),
    "mmvae_objective_elbo": (c_i, [ctypes.POINTER(c_p), ctypes.POINTER(c_i64), ctypes.POINTER(c_f), c_i, c_p, c_i64,
                                   ctypes.POINTER(c_f), ctypes.POINTER(c_f), c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmvae_peer_error_offset": (c_i64, []),
    "mmvae_prior_scale_bwd_peer": (c_i, [c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_p]),
    "mmvae_objective_dreg_stage2_peer": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_p]),
    "mmvae_peer_allreduce_f64": (c_i, [c_p, c_i, c_p, c_i, c_i, c_i, c_p]),
    "mmvae_scale_inplace": (c_i, [c_p, c_i, c_i64, c_p, c_p]),
}

_lib = None
# launches of OUR kernels issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0
_LAUNCHES = {"mmvae_loglik_rowreduce_fwd": 1, "mmvae_loglik_rowreduce_bwd": 1, "mmvae_loglik_rowreduce_fused": 1,
             "mmvae_catce_rows": 1, "mmvae_osigma_sumsq": 1, "mmvae_osigma_fwd": 1, "mmvae_osigma_bwd": 2,
             "mmvae_latent_draws_fwd": 1, "mmvae_latent_draws_bwd": 2, "mmvae_moe_logdens_fwd": 1,
             "mmvae_latent_draws_fwd_tail": 1, "mmvae_latent_draws_bwd_tail": 2, "mmvae_moe_logdens_fwd_tail": 1,
             "mmvae_moe_logdens_bwd_tail": 2,
             "mmvae_moe_logdens_bwd": 2, "mmvae_moe_logdens_bwd_rk": 2, "mmvae_objective_iwae": 1, "mmvae_objective_iwae_fused": 1, "mmvae_objective_dreg_stage1": 2, "mmvae_objective_dreg_stage1_ptrs": 2,
             "mmvae_objective_dreg_stage2": 1, "mmvae_reduce_sum": 1, "mmvae_scale_inplace": 1}


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU / eager fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("mmvae_b200: %s not found -- build it with `python multimodal-vae-comparison_b200/build.py` "
                           "(or __graft_entry__.build()); there is no fallback path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mmvae_version() != 1:
        raise RuntimeError("mmvae_b200: ABI version mismatch")
    _lib = lib
    return lib


# optional per-kernel CUDA-event timer (bench.py roofline leg): object with .names and .wrap(name, fn)
timer = None


def call(name, *args):
    """Invoke an int-returning entry point and turn a non-zero status into a RuntimeError."""
    global launch_count
    fn = getattr(load(), name)
    if timer is not None and name in timer.names:
        rc = timer.wrap(name, lambda: fn(*args), args)
    else:
        rc = fn(*args)
    if rc != 0:
        kind = {-1: "bad argument", -2: "unknown enum", -3: "size limit"}.get(rc, "cudaError_t %d" % rc)
        raise RuntimeError("mmvae_b200: %s failed: %s" % (name, kind))
    launch_count += _LAUNCHES.get(name, 1)
    return rc
