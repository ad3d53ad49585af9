#!/bin/bash
# Compile variants of csrc/catce.cu with tuning macros (in parallel) and time the category_ce kernels alone at the
# shapes of the benchmark configurations; every variant is checked against a torch fp64 evaluation of
# nn.CrossEntropyLoss semantics (class axis = dim 1).  Run on the GPU box: `gpurun -- tools/tune_catce.sh`.
set -u
cd "$(dirname "$0")/.."
SRC=multimodal-vae-comparison_b200/csrc
OUT=gpurun_out/tune_catce; mkdir -p $OUT; rm -f $OUT/lib_*.so
declare -A V
V[r1_staged]="-DMMVAE_CATCE_PAIRS=0 -DMMVAE_CATCE_RESIDENT=0"
V[r2_tma_w4]="-DMMVAE_CATCE_RESIDENT=0"
V[r2_flat_r4w4]="-DMMVAE_CATCE_RESIDENT=0 -DMMVAE_CATCE_PAIRS_STAGE_X=0"
V[resident_occ3]=""
V[resident_occ2]="-DMMVAE_CATCE_RESIDENT_OCC=2"
V[resident_occ4]="-DMMVAE_CATCE_RESIDENT_OCC=4"
for k in "${!V[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude ${V[$k]} -shared $SRC/catce.cu -o $OUT/lib_$k.so -lcudart > $OUT/build_$k.log 2>&1 &
done
wait
python - <<'PY'
import ctypes, glob, os, torch
c_p, c_i, c_i64, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
flush = torch.empty(64 << 20, device="cuda")
# (name, rows, B, C, d, recon dtype)
SHAPES = [("c5_text_bf16", 4096, 4096, 246, 27, torch.bfloat16), ("bf16_c128_d27", 4096, 4096, 128, 27, torch.bfloat16),
          ("bf16_c100_d40", 4096, 1024, 100, 40, torch.bfloat16), ("bf16_c64_d6", 8192, 4096, 64, 6, torch.bfloat16),
          ("c2_text", 7680, 256, 45, 27, torch.float32), ("c5_text_f32", 4096, 4096, 246, 27, torch.float32)]
def ref(x, t, rows, B, C, d, w):
    xd = x.double().view(rows, C, d).requires_grad_(True)
    td = t.double().view(B, C, d).repeat(rows // B, 1, 1)
    val = (td * torch.log_softmax(xd, 1)).sum((1, 2))
    (val * w.double()).sum().backward()
    return val, xd.grad.view(rows, C * d)
lines = []
for name, rows, B, C, d, dt in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    n = C * d
    x = torch.randn(rows, n, device="cuda", generator=g).to(dt)
    t = torch.nn.functional.one_hot(torch.randint(d, (B, C), device="cuda", generator=g), d).float().view(B, n)
    t = t * (torch.rand(B, C, 1, device="cuda", generator=g) > 0.3).float().expand(B, C, d).reshape(B, n)  # padded positions
    t = t.to(dt).contiguous()  # bf16 configurations carry bf16 targets
    w = torch.randn(rows, device="cuda", generator=g)
    rv, rg = ref(x, t, rows, B, C, d, w)
    out = torch.empty(rows, device="cuda"); grad = torch.empty_like(x); stats = torch.empty(rows, 2, d, device="cuda")
    dtc = 0 if dt == torch.float32 else 1
    nb_f, nb_b = x.numel() * x.element_size() + t.numel() * t.element_size(), 2 * x.numel() * x.element_size() + t.numel() * t.element_size()
    for lib in sorted(glob.glob("gpurun_out/tune_catce/lib_*.so")):
        L = ctypes.CDLL(lib)
        f = L.mmvae_catce_rows; f.restype = c_i
        f.argtypes = [c_i, c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i64, c_f, c_p, c_f, c_p, c_p, c_i64, c_p, c_p]
        st = torch.cuda.current_stream().cuda_stream
        def run(mode, with_stats=True):
            rc = f(mode, x.data_ptr(), n, dtc, t.data_ptr(), n, dtc, rows, B, C, d, 1.0, w.data_ptr(), 0.0, out.data_ptr(),
                   grad.data_ptr(), n, stats.data_ptr() if with_stats else None, st)
            assert rc == 0, (lib, mode, rc)
        tm = {}
        for key, fn in (("fwd", lambda: run(0)), ("bwd", lambda: run(1)), ("fused", lambda: run(2, False))):
            for _ in range(3): fn()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
            ts.sort(); tm[key] = sum(ts[2:-2]) / len(ts[2:-2])
        tol = 2e-2 if dt == torch.bfloat16 else 1e-5
        errs = []
        out.zero_(); grad.zero_(); run(0); run(1); torch.cuda.synchronize()
        errs.append(float((out.double() - rv).abs().max() / rv.abs().max()))
        errs.append(float((grad.double() - rg).abs().max() / rg.abs().max()))
        out.zero_(); grad.zero_(); run(2, False); torch.cuda.synchronize()
        errs.append(float((out.double() - rv).abs().max() / rv.abs().max()))
        errs.append(float((grad.double() - rg).abs().max() / rg.abs().max()))
        grad.zero_(); run(1, False); torch.cuda.synchronize()
        errs.append(float((grad.double() - rg).abs().max() / rg.abs().max()))
        line = "%-13s %-16s fwd %6.1f us (%4.0f GB/s)  bwd %6.1f us (%4.0f GB/s)  fused %6.1f us (%4.0f GB/s)  max rel err %.1e %s" % (
            name, os.path.basename(lib)[4:-3], tm["fwd"] * 1e3, nb_f / tm["fwd"] / 1e6, tm["bwd"] * 1e3, nb_b / tm["bwd"] / 1e6,
            tm["fused"] * 1e3, nb_b / tm["fused"] / 1e6, max(errs), "OK" if max(errs) < tol else "MISMATCH")
        print(line); lines.append(line)
open("gpurun_out/tune_catce/results.txt", "w").write("\n".join(lines) + "\n")
PY
grep -l error $OUT/build_*.log | head
rm -f $OUT/lib_*.so
