#!/usr/bin/env python
"""Raw pinned host -> device copy rate of this box (105 MB, the bench's per-step input), CUDA events."""
import torch
n = 104857600
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"pinned H2D {n / 1e6:.0f} MB: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
