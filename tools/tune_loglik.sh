#!/bin/bash
# Compile variants of csrc/loglik.cu with tuning macros (in parallel) and time the BCE forward / backward kernels
# alone at the C2 shape (7680 x 12288 fp32, target 256 x 12288).  Run on the GPU box: `gpurun -- tools/tune_loglik.sh`.
set -u
cd "$(dirname "$0")/.."
SRC=multimodal-vae-comparison_b200/csrc
OUT=gpurun_out/tune; mkdir -p $OUT; rm -f $OUT/lib_*.so
declare -A V
V[cur]=""
V[nofast]="-DMMVAE_BCE_FAST=0"
for k in "${!V[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude ${V[$k]} -shared $SRC/loglik.cu -o $OUT/lib_$k.so -lcudart > $OUT/build_$k.log 2>&1 &
done
# optional A/B against an older source dropped next to this script (not committed)
[ -f tools/_loglik_prev.cu ] && cp tools/_loglik_prev.cu $SRC/_loglik_prev.cu && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -shared $SRC/_loglik_prev.cu -o $OUT/lib_a_prev.so -lcudart > $OUT/build_a_prev.log 2>&1
rm -f $SRC/_loglik_prev.cu
wait
python - <<'PY'
import ctypes, glob, os, torch
c_p, c_i, c_i64, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
flush = torch.empty(64 << 20, device="cuda")
lines = []
for tag, rows, B, P, dt in (("c2_image_f32", 7680, 256, 12288, torch.float32), ("c5_image_bf16", 4096, 4096, 12288, torch.bfloat16),
                            ("c1_image_f32", 4096, 4096, 12288, torch.float32)):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.sigmoid(torch.randn(rows, P, device="cuda", generator=g)).clamp(1e-6, 1 - 1e-6).to(dt)
    t = torch.rand(B, P, device="cuda", generator=g).to(dt)
    w = torch.randn(rows, device="cuda", generator=g)
    out = torch.empty(rows, device="cuda"); grad = torch.empty_like(x)
    dc = 0 if dt == torch.float32 else 1
    es = x.element_size()
    ref = None
    for lib in sorted(glob.glob("gpurun_out/tune/lib_*.so")):
        L = ctypes.CDLL(lib)
        f = L.mmvae_loglik_rowreduce_fwd; f.restype = c_i
        f.argtypes = [c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_p]
        b = L.mmvae_loglik_rowreduce_bwd; b.restype = c_i
        b.argtypes = [c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_i64, c_p]
        u = L.mmvae_loglik_rowreduce_fused; u.restype = c_i
        u.argtypes = [c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_f, c_p, c_p, c_i64, c_p, c_p]
        st = torch.cuda.current_stream().cuda_stream
        def run_f(): assert f(x.data_ptr(), P, dc, t.data_ptr(), P, dc, rows, B, P, 0, 0.75, 1.0, out.data_ptr(), None, st) == 0
        def run_b(): assert b(x.data_ptr(), P, dc, t.data_ptr(), P, dc, rows, B, P, 0, 0.75, 1.0, w.data_ptr(), grad.data_ptr(), P, st) == 0
        def run_u(): assert u(x.data_ptr(), P, dc, t.data_ptr(), P, dc, rows, B, P, 0, 0.75, 1.0, None, -1.0, out.data_ptr(), grad.data_ptr(), P, None, st) == 0
        tm = {}
        for name, fn in (("fwd", run_f), ("bwd", run_b), ("fused", run_u)):
            for _ in range(5): fn()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
            ts.sort(); tm[name] = sum(ts[2:-2]) / len(ts[2:-2])
        run_u(); torch.cuda.synchronize()
        cur = (out.clone(), grad.float().clone())
        if ref is None: ref = cur
        err = max(float((cur[0] - ref[0]).abs().max() / ref[0].abs().max()), float((cur[1] - ref[1]).abs().max() / ref[1].abs().max()))
        R, T = rows * P * es, B * P * es
        line = "%-14s %-12s fwd %6.1f us (%4.0f GB/s)  bwd %6.1f us (%4.0f GB/s)  fused %6.1f us (%4.0f GB/s)  max rel diff vs first %.1e" % (
            tag, os.path.basename(lib)[4:-3], tm["fwd"] * 1e3, (R + T) / tm["fwd"] / 1e6, tm["bwd"] * 1e3, (2 * R + T) / tm["bwd"] / 1e6,
            tm["fused"] * 1e3, (2 * R + T) / tm["fused"] / 1e6, err)
        print(line); lines.append(line)
open("gpurun_out/tune/results.txt", "w").write("\n".join(lines) + "\n")
PY
grep -l error $OUT/build_*.log | head
