"""Reference rates on this box: read-only reduction vs copy (context for the roofline of read-only kernels)."""
import torch
x = torch.randn(7680 * 12288, device="cuda")
y = torch.empty_like(x)
flush = torch.empty(64 << 20, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return sum(ts[2:-2]) / len(ts[2:-2])
nb = x.numel() * 4
r = t(lambda: x.sum())
c = t(lambda: y.copy_(x))
m = t(lambda: torch.max(x))
big = torch.randn(1 << 30, device="cuda")
rb = t(lambda: big.sum(), 10)
print("read-only sum   377 MB: %.1f us  %.0f GB/s" % (r * 1e3, nb / r / 1e6))
print("read-only max   377 MB: %.1f us  %.0f GB/s" % (m * 1e3, nb / m / 1e6))
print("copy (rd+wr)    755 MB: %.1f us  %.0f GB/s" % (c * 1e3, 2 * nb / c / 1e6))
print("read-only sum   4.3 GB: %.1f us  %.0f GB/s" % (rb * 1e3, big.numel() * 4 / rb / 1e6))
