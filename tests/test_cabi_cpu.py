"""CPU: the C-ABI library loads and exports exactly what include/mmvae_b200.h declares; the product package is
isolated from the oracle; ops refuse CPU tensors (no fallback); host-side integer logic."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-vae-comparison_b200")
HEADER = os.path.join(ROOT, "include", "mmvae_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import mmvae_b200._lib as L
    return L


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MMVAE_API\s+(?:int64_t|int)\s+(mmvae_\w+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), "symbol %s declared in the header but not exported" % n
    assert sorted(lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_only_cabi_symbols_are_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(exported) == declared_symbols()


def test_version_and_argument_errors_without_gpu(lib):
    L = lib.load()
    assert L.mmvae_version() == 1
    # argument validation happens before any launch: callable without a GPU
    assert L.mmvae_loglik_rowreduce_fwd(None, 0, 0, None, 0, 0, 0, 0, 0, 0, 0.75, 1.0, None, None, None) == -1
    assert L.mmvae_loglik_workspace_bytes(32, 12288, 0) > 0      # few long rows are split over CTAs
    assert L.mmvae_loglik_workspace_bytes(7680, 12288, 0) == 0   # one CTA per row: no workspace
    assert L.mmvae_reduce_sum(None, 0, 1.0, None, None) == -1


def test_fused_iwae_argument_validation_without_gpu(lib):
    """mmvae_objective_iwae_fused rejects missing outputs and a loss_sum without its ticket before any launch."""
    L = lib.load()
    one = ctypes.c_void_p(16)  # never dereferenced on the host: validation only looks at NULL-ness
    ptrs = (ctypes.c_void_p * 1)(16)
    args = dict(lpz=one, lq=one, lpx=None, ptrs=ptrs, M=1, L=1, K=2, B=4, beta=1.0)

    def run(lw, loss_b, w, loss_sum, ticket):
        return L.mmvae_objective_iwae_fused(args["lpz"], args["lq"], args["lpx"], args["ptrs"], args["M"], args["L"],
                                            args["K"], args["B"], args["beta"], lw, loss_b, w, None, None, loss_sum,
                                            ticket, None)
    assert run(None, one, one, None, None) == -1          # lw missing
    assert run(one, one, one, one, None) == -1            # loss_sum without ticket
    assert run(one, one, one, None, one) == -1            # ticket without loss_sum


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's stdout contract: ONE JSON line, whatever libraries print (fd 1 is pointed at stderr for the run)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "1", "--cpu-batch", "2"], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:400]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "objective fwd+bwd samples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_draw_desc_layout_matches_header(lib):
    assert ctypes.sizeof(lib.DrawDesc) == 56


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_ops_refuse_cpu_tensors():
    import mmvae_b200.ops as ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.loglik_rows(torch.rand(4, 8), torch.rand(4, 8), "bce")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.moe_logdens(torch.rand(2, 3, 4), torch.rand(2, 3, 4), torch.zeros(1, 4), torch.ones(1, 4),
                        torch.rand(2, 1, 3, 4), [0, 0])


def test_graphed_objective_refuses_cpu_models_and_tracks_batch_signature():
    import mmvae_b200
    from mmvae_b200.graphed import _signature
    lin = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        mmvae_b200.GraphedObjective(lin, {"mod_1": {"data": torch.zeros(2, 2), "masks": None, "categorical": False}})
    a = {"mod_1": {"data": torch.zeros(4, 3), "masks": None, "categorical": False}}
    b = {"mod_1": {"data": torch.ones(4, 3), "masks": None, "categorical": False}}
    c = {"mod_1": {"data": torch.zeros(5, 3), "masks": None, "categorical": False}}
    d = {"mod_1": {"data": torch.zeros(4, 3), "masks": torch.ones(4, 3, dtype=torch.bool), "categorical": False}}
    assert _signature(a) == _signature(b) and _signature(a) != _signature(c) and _signature(a) != _signature(d)


def test_missing_library_fails_loudly(monkeypatch):
    import mmvae_b200._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no fallback"):
        L.load()


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "refmath" in src or "/root/reference" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_subset_orders_and_chunk_maps(golden):
    import mmvae_b200.mmvae_models as mm
    from oracle import refmath
    assert mm.poe_subsets(range(3)) == refmath.poe_subsets(range(3))
    assert mm.mopoe_subsets(range(4)) == refmath.mopoe_subsets(range(4))
    assert mm.mopoe_subsets(["mod_1", "mod_2", "mod_3"])[3] == ("mod_1", "mod_2")
    for (S, B), ends in golden["chunk_ends"].items():
        assert list(mm.mopoe_chunk_bounds(S, B)[1]) == ends, (S, B)  # bit exact vs the reference run
    for S in (1, 3, 7, 15, 31, 63):
        assert mm.mopoe_inmodel_component(S) == S - 1
    masks = mm.subset_bitmasks(mm.mopoe_subsets(range(3)), 3)
    assert masks.tolist()[:6] == [1, 2, 4, 3, 5, 6] and (masks[6].item() & 0xffffffff) == (7 | 1 << 31)


def test_registry_and_plugin_surface():
    import mmvae_b200
    from oracle import cases
    assert set(mmvae_b200.MODEL_REGISTRY) == {"moe", "poe", "mopoe", "dmvae"}
    case = cases.case_list()[0]
    m = mmvae_b200.poe(cases.build_vaes(case), case["D"], {"obj": "elbo", "beta": 1.0, "K": 1}, None)
    assert isinstance(m, mmvae_b200.TorchMMVAE) and m.modelName == "poe" and m.K == 1
    mu0, s0 = m.pz_params
    assert mu0.shape == (1, case["D"]) and torch.allclose(s0, torch.ones(1, case["D"]))
    assert {"_pz_params.0", "_pz_params.1"} <= set(m.state_dict())
    assert m.obj_fn.obj_name == "elbo"
    with pytest.raises(AssertionError):
        mmvae_b200.MultimodalObjective("nope")
    m.obj_fn.set_ltype("bce")
    with pytest.raises(AssertionError):
        m.obj_fn.set_ltype("does_not_exist")
    with pytest.raises(ValueError):  # the reference also needs every modality in objective()
        m.objective({"mod_1": {"data": None, "masks": None}, "mod_2": {"data": torch.zeros(2, 5, 27), "masks": None}})


def test_algorithmic_bytes_match_survey():
    import mmvae_b200.synthetic as syn
    import mmvae_b200.workloads as W
    b = {k: W.algorithmic_bytes(dict(v)) for k, v in syn.WORKLOADS.items()}
    assert abs(b["c1_poe_elbo_cdsprites_l1"] - 449e3) / 449e3 < 0.01
    assert abs(b["c2_moe_iwae_cdsprites_l5"] - 9.95e6) / 9.95e6 < 0.01
    assert abs(b["c3_mopoe_elbo_sprites"] - 1.18e6) / 1.18e6 < 0.01
    # C5: bf16 reconstructions, gradients AND targets (3 x (3*12288)*2 + 3 x (3*6642)*2 = 341 KB): 3 terms
    # per modality, 6 bytes per feature
    c5 = W.algorithmic_bytes(dict(syn.WORKLOADS["c5_dmvae_elbo_cub"]), torch.bfloat16)
    assert abs(c5 - 341e3) / 341e3 < 0.01
    assert abs(c5 - 3 * 6 * (12288 + 6642)) / c5 < 0.01


def test_helper_module_matches_reference_semantics(golden):
    import mmvae_b200.utils as U
    from oracle import refmath
    v = torch.randn(6, 5)
    assert torch.equal(U.log_mean_exp(v), refmath.log_mean_exp(v))
    assert U.combinatorial(["a", "b", "c"]) == [("a", "b"), ("a", "c"), ("b", "c")]
    batch = {"mod_1": {"data": torch.zeros(3, 2), "masks": None, "categorical": False},
             "mod_2": {"data": torch.ones(3, 4), "masks": torch.ones(3, 4, dtype=torch.bool), "categorical": True}}
    subs = U.subsample_input_modalities(batch)
    assert len(subs) == 3 and U.find_out_batch_size(subs[1]) == 3
    assert subs[0]["mod_2"]["data"] is None and subs[0]["mod_2"]["masks"] is None and subs[0]["mod_2"]["categorical"] is True
    assert subs[2]["mod_1"]["data"] is batch["mod_1"]["data"]  # shared, not deep-copied
    import torch.distributions as dist
    p, q = dist.Normal(torch.zeros(4, 3), torch.ones(4, 3) * 0.5), dist.Normal(torch.zeros(1, 3), torch.ones(1, 3))
    assert torch.allclose(U.kl_divergence(p, q), refmath.kl_normal_normal(p.loc, p.scale, q.loc, q.scale))  # CPU: torch registry


def test_encoder_tail_plumbing_on_cpu():
    """Host logic of the fused encoder tail (no kernels involved): encode() keeps reference semantics for every caller
    except the objectives; _stack_raw hands raw logits to the kernels only when every encoder is raw and of one width."""
    import torch
    import mmvae_b200
    import mmvae_b200.synthetic as syn
    torch.manual_seed(0)
    D = 8

    def vaes(raw1, raw2):
        return {"mod_1": syn.StubVAE(syn.LinearEncoder((3, 4, 4), D, returns_raw_logvar=raw1),
                                     syn.LinearDecoder(D, (3, 4, 4), squash=True), D, "bce", id_name="mod_1"),
                "mod_2": syn.StubVAE(syn.LinearEncoder((5, 27), D, returns_raw_logvar=raw2),
                                     syn.LinearDecoder(D, (5, 27), squash=False), D, "category_ce", id_name="mod_2")}

    batch = {"mod_1": {"data": torch.rand(6, 3, 4, 4), "masks": None, "categorical": False},
             "mod_2": {"data": torch.rand(6, 5, 27), "masks": None, "categorical": False}}
    m = mmvae_b200.MODEL_REGISTRY["poe"](vaes(True, True), D, {"obj": "elbo", "beta": 1.0, "K": 1}, None)
    enc = m.encode(batch)  # reference semantics: (mu, softmax(raw) + 1e-6)
    s = enc["mod_1"]["shared"][1]
    assert not enc["mod_1"]["raw"] and torch.allclose(s.sum(-1), torch.full((6,), 1 + D * 1e-6), atol=1e-5) and float(s.min()) > 0
    enc_raw = m.encode(batch, raw_ok=True)
    assert enc_raw["mod_1"]["raw"] and enc_raw["mod_1"]["shared"] is None
    mu, sraw, flag = m._stack_raw(enc_raw, ["mod_1", "mod_2"], "shared")
    assert flag and sraw.shape == (2, 6, D) and not torch.allclose(sraw.sum(-1), torch.ones(2, 6))
    assert torch.allclose(torch.softmax(sraw[0], -1) + 1e-6, s, atol=1e-6)
    # mixed encoders: the tail is applied on the host, the kernels get scales
    m2 = mmvae_b200.MODEL_REGISTRY["poe"](vaes(True, False), D, {"obj": "elbo", "beta": 1.0, "K": 1}, None)
    mu2, s2, flag2 = m2._stack_raw(m2.encode(batch, raw_ok=True), ["mod_1", "mod_2"], "shared")
    assert not flag2 and torch.allclose(s2.sum(-1), torch.full((2, 6), 1 + D * 1e-6), atol=1e-5)


def test_bench_sweep_and_fold_flags_parse():
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True).stdout
    for flag in ("--sweep", "--fold-terms", "--nccl-only", "--cpu-budget-s", "--impl"):
        assert flag in out


def test_round2_entry_points_validate_arguments_without_gpu(lib):
    """Every check below returns before the first CUDA call: bad arguments are MMVAE_E_ARG (-1) / MMVAE_E_LIMIT (-3)."""
    import ctypes
    c_p, c_i, c_i64, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    P = lambda v=0x1000: c_p(v)  # never dereferenced on these paths
    lib = ctypes.CDLL(lib.LIB_PATH)  # a private handle: its argtypes do not touch the package's binding
    # ELBO combine: no terms and no KL segments; too many terms
    f = lib.mmvae_objective_elbo
    f.restype = c_i
    f.argtypes = [c_p, c_p, c_p, c_i, c_p, c_i64, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
    assert f(None, None, None, 0, None, 0, None, None, 0, P(), None, None, None, None, None) == -1
    assert f(P(), P(), P(), 49, None, 0, None, None, 0, P(), None, None, None, None, None) == -3
    assert f(None, None, None, 0, None, 8, P(), None, 2, P(), None, None, None, None, None) == -1  # KL segments without rows
    # masked category_ce: mask rows shorter than the class axis
    g = lib.mmvae_catce_rows_masked
    g.restype = c_i
    g.argtypes = [c_i, c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i64, c_f, c_p, c_f, c_p, c_p, c_i64, c_p,
                  c_p, c_i64, c_p]
    assert g(0, P(), 45 * 27, 0, P(), 45 * 27, 0, 8, 8, 45, 27, 1.0, None, 0.0, P(), None, 0, None, P(), 44, None) == -1
    # fused encoder tail: only the flat MoE kernels have it (D % 4 == 0, D <= 128, M <= 3)
    h = lib.mmvae_moe_logdens_fwd_tail
    h.restype = c_i
    h.argtypes = [c_p, c_p, c_i, c_i64, c_i, c_i, ctypes.POINTER(ctypes.c_int32), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
    dists = (ctypes.c_int32 * 2)(0, 0)
    assert h(P(), P(), 2, 8, 10, 3, dists, P(), P(), P(), P(), P(), P(), None, None) == -3
    # peer-memory collectives: more values than a 4 KB slot, bad rank
    k = lib.mmvae_peer_allreduce_f64
    k.restype = c_i
    k.argtypes = [c_p, c_i, c_p, c_i, c_i, c_i, c_p]
    assert k(P(), 513, P(), 0, 2, 0, None) == -3
    assert k(P(), 4, P(), 2, 2, 0, None) == -1
    assert k(P(), 4, P(), 0, 2, 8, None) == -3  # channel >= MMVAE_PEER_CHANNELS
    # DReG packed coefficients need M == 2
    s1 = lib.mmvae_objective_dreg_stage1_ptrs
    s1.restype = c_i
    s1.argtypes = [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_i, c_p]
    assert s1(P(), P(), P(), None, 3, 1, 4, 8, P(), P(), 1, None) == -1
