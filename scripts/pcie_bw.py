"""PCIe copy bandwidth of the box (pinned host memory, 128 MB, CUDA events): the floor under the
host-buffer (e2e) legs of bench.py."""
import torch
n = 128 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
h2d = t(lambda: d.copy_(h, non_blocking=True)); d2h = t(lambda: h.copy_(d, non_blocking=True))
print("H2D %.1f GB/s  D2H %.1f GB/s" % (n / h2d / 1e6, n / d2h / 1e6))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
both(); torch.cuda.synchronize(); a.record()
for _ in range(10): both()
torch.cuda.synchronize(); b.record(); torch.cuda.synchronize()
print("both directions at once: %.1f GB/s each" % (n / (a.elapsed_time(b) / 10) / 1e6))
