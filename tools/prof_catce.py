"""ncu target: a few category_ce launches at the C2 text shape from one of the tune_catce variant libraries.
    ncu --set full ... python tools/prof_catce.py gpurun_out/tune_catce/lib_ring_xb24.so"""
import ctypes, sys, torch
c_p, c_i, c_i64, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
rows, B, C, d = 7680, 256, 45, 27
n = C * d
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, n, device="cuda", generator=g)
t = torch.nn.functional.one_hot(torch.randint(d, (B, C), device="cuda", generator=g), d).float().view(B, n)
w = torch.randn(rows, device="cuda", generator=g)
out = torch.empty(rows, device="cuda"); grad = torch.empty_like(x); stats = torch.empty(rows, 2, d, device="cuda")
L = ctypes.CDLL(sys.argv[1])
f = L.mmvae_catce_rows; f.restype = c_i
f.argtypes = [c_i, c_p, c_i64, c_i, c_p, c_i64, c_i, c_i64, c_i64, c_i64, c_i64, c_f, c_p, c_f, c_p, c_p, c_i64, c_p, c_p]
st = torch.cuda.current_stream().cuda_stream
for mode in (0, 1, 2, 0, 1, 2):
    rc = f(mode, x.data_ptr(), n, 0, t.data_ptr(), n, 0, rows, B, C, d, 1.0, w.data_ptr(), 0.0, out.data_ptr(),
           grad.data_ptr(), n, stats.data_ptr() if mode != 2 else None, st)
    assert rc == 0
torch.cuda.synchronize()
