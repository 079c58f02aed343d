"""Pinned host<->device copy bandwidth of the box (context for the e2e number; not a product path)."""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (3 * n / (time.perf_counter() - t) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("both directions at once: %.1f GB/s each" % (3 * n / (time.perf_counter() - t) / 1e9))
