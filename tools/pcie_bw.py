"""Diagnostic: pinned host<->device copy bandwidth on this box (bounds the e2e numbers of bench.py)."""
import torch, time
n = 788_070_400
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, "%.2f ms for %d MB -> %.1f GB/s" % (ms, n >> 20, n / ms / 1e6))
