"""Page-locked host <-> device copy bandwidth of the box (the ceiling of the CSV-returning end-to-end figure)."""
import time
import torch
n = 1 << 28                                   # 2 GiB of doubles
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print("%s pinned: %.1f GB/s" % (name, 3 * n * 8 / (time.perf_counter() - t0) / 1e9))
