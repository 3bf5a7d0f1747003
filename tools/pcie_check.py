"""Host<->device copy rates of the box (pinned memory): H2D alone, D2H alone, both at once; and the e2e solve."""
import time
import torch
n = 1342210048 // 8
h1 = torch.empty(n, dtype=torch.float64).pin_memory()
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda")
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
gb = n * 8 / 1e9
print(f"H2D alone: {gb / t(lambda: d1.copy_(h1, non_blocking=True)):.1f} GB/s")
print(f"D2H alone: {gb / t(lambda: h2.copy_(d2, non_blocking=True)):.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both)
print(f"both at once: {gb / dt:.1f} GB/s each direction ({dt * 1e3:.1f} ms for 1.34 GB each way)")
