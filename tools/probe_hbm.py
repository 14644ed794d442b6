"""What HBM delivers for simple streaming patterns (torch elementwise kernels as a yardstick, not product code)."""
import torch
n = 256 * 64 * 64 * 256  # elements of a (256,64,64,256) bf16 tensor = 537 MB
a = torch.randn(n, device="cuda").to(torch.bfloat16)
b = torch.randn(n, device="cuda").to(torch.bfloat16)
c = torch.empty_like(a)
small = a[: n // 4]


def t(fn, bytes_, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:40s} {ms * 1e3:8.1f} us  {bytes_ / ms / 1e6:8.1f} GB/s")


t(lambda: c.copy_(a), 2 * n * 2, "copy (1R + 1W)")
t(lambda: torch.add(a, b, out=c), 3 * n * 2, "add (2R + 1W)")
t(lambda: a.sum(), n * 2, "sum (1R)")
t(lambda: c.zero_(), n * 2, "zero (1W)")
t(lambda: torch.relu(a, out=c) if False else torch.clamp_min(a, 0, out=c), 2 * n * 2, "relu (1R + 1W)")
