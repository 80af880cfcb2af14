"""Pinned host -> device copy rate of the box (context for bench.py's e2e leg)."""
import torch
for mb in (8, 32, 128):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D %4d MiB: %.1f GB/s" % (mb, 20 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9))
