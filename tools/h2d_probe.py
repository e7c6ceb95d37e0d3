"""Raw pinned H2D / D2H bandwidth of the box (context for the e2e number)."""
import torch, time
n = 338 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        a.copy_(b, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s %d MB: %.2f ms  %.1f GB/s" % (name, n // 1000000, ms, n / ms / 1e6))
