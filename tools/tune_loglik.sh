#!/bin/bash
# Compile variants of csrc/loglik.cu with tuning macros (in parallel) and time the BCE forward / backward kernels
# alone at the C2 shape (7680 x 12288 fp32, target 256 x 12288).  Run on the GPU box: `gpurun -- tools/tune_loglik.sh`.
set -u
cd "$(dirname "$0")/.."
SRC=multimodal-vae-comparison_b200/csrc
OUT=gpurun_out/tune; mkdir -p $OUT
declare -A V
V[base]="-DMMVAE_BCE_FAST=0"
V[fast]="-DMMVAE_BCE_FAST=1"
V[fast_mb6]="-DMMVAE_BCE_FAST=1 -DMMVAE_FWD_MINBLOCKS=6"
V[fast_mb4]="-DMMVAE_BCE_FAST=1 -DMMVAE_FWD_MINBLOCKS=4"
V[fast_mb6_u3]="-DMMVAE_BCE_FAST=1 -DMMVAE_FWD_MINBLOCKS=6 -DMMVAE_FWD_UNROLL=3"
for k in "${!V[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude ${V[$k]} -shared $SRC/loglik.cu -o $OUT/lib_$k.so -lcudart > $OUT/build_$k.log 2>&1 &
done
wait
python - <<'PY'
import ctypes, glob, os, torch
c_p, c_i, c_i64, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
rows, B, P = 7680, 256, 12288
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.sigmoid(torch.randn(rows, P, device="cuda", generator=g)).clamp(1e-6, 1 - 1e-6)
t = torch.rand(B, P, device="cuda", generator=g)
w = torch.randn(rows, device="cuda", generator=g)
out = torch.empty(rows, device="cuda"); grad = torch.empty_like(x)
flush = torch.empty(64 << 20, device="cuda")
ref = None
res = []
for lib in sorted(glob.glob("gpurun_out/tune/lib_*.so")):
    L = ctypes.CDLL(lib)
    f = L.mmvae_loglik_rowreduce_fwd; f.restype = c_i
    f.argtypes = [c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_p]
    b = L.mmvae_loglik_rowreduce_bwd; b.restype = c_i
    b.argtypes = [c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_p, c_p, c_i64, c_p]
    st = torch.cuda.current_stream().cuda_stream
    def run_f(): assert f(x.data_ptr(), P, 0, t.data_ptr(), P, 0, rows, B, P, 0, 0.75, 1.0, out.data_ptr(), None, st) == 0
    def run_b(): assert b(x.data_ptr(), P, 0, t.data_ptr(), P, 0, rows, B, P, 0, 0.75, 1.0, w.data_ptr(), grad.data_ptr(), P, st) == 0
    tm = {}
    for name, fn in (("fwd", run_f), ("bwd", run_b)):
        for _ in range(5): fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
        ts.sort(); tm[name] = sum(ts[2:-2]) / len(ts[2:-2])
    run_f(); torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    res.append((os.path.basename(lib), tm["fwd"], tm["bwd"], err))
with open("gpurun_out/tune/results.txt", "w") as fh:
    for n, a, b_, e in res:
        line = "%-22s fwd %.1f us (%.0f GB/s)   bwd %.1f us (%.0f GB/s)   max rel diff vs base %.1e" % (
            n, a * 1e3, (rows * P * 4 + B * P * 4) / a / 1e6, b_ * 1e3, (2 * rows * P * 4 + B * P * 4) / b_ / 1e6, e)
        print(line); fh.write(line + "\n")
PY
grep -l error $OUT/build_*.log | head
