"""Pinned host->device copy bandwidth of this box (the ceiling of bench.py's e2e)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print("pinned H2D 1 GiB: %.2f ms  %.1f GB/s  (%.0f Gbit/s)" % (dt * 1e3, n / dt / 1e9, n * 8 / dt / 1e9))
