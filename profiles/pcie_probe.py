#!/usr/bin/env python
"""Host<->device copy bandwidth of the box: each direction alone, then both at once (the ceiling of the pipelined C-ABI calls)."""
import time
import torch

n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


for name, a, b in (("H2D", 1, 0), ("D2H", 0, 1), ("both", 1, 1)):
    t = run(a, b)
    print(f"{name}: {1e3 * t:.1f} ms for {a + b} GiB -> {(a + b) * n / t / 1e9:.1f} GB/s")
