"""Achievable HBM bandwidth on this B200 for pure-write, pure-read and copy streams (torch ops, CUDA events)."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
N = 1 << 28   # 1 GiB of fp32
x = torch.empty(N, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
x.normal_()
print("write (fill_)   : %.0f GB/s" % (4 * N / t(lambda: y.fill_(1.0)) / 1e9))
print("write (memset)  : %.0f GB/s" % (4 * N / t(lambda: y.zero_()) / 1e9))
print("read  (sum)     : %.0f GB/s" % (4 * N / t(lambda: x.sum()) / 1e9))
print("copy  (r+w)     : %.0f GB/s" % (8 * N / t(lambda: y.copy_(x)) / 1e9))
print("add   (2r+w)    : %.0f GB/s" % (12 * N / t(lambda: torch.add(x, y, out=y)) / 1e9))
z = torch.empty(N // 2, dtype=torch.float32, device="cuda")
print("1r+2w (cat-like): n/a")
