"""PCIe ceiling of the box the bench runs on: pinned H2D / D2H copies of the e2e leg's sizes (not part of the product)."""
import time
import torch

def bw(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

n = 1024 * 196608
h = torch.empty((1024, 196608), dtype=torch.complex64).pin_memory()
d = torch.empty_like(h, device="cuda")
hb = torch.empty((1024, 230400), dtype=torch.int8).pin_memory()
db = torch.empty_like(hb, device="cuda")
print("H2D one 1.6 GB copy      GB/s", round(bw(lambda: d.copy_(h, non_blocking=True), h.numel() * 8), 1))
print("D2H one 236 MB copy      GB/s", round(bw(lambda: hb.copy_(db, non_blocking=True), hb.numel()), 1))
def rows():
    for s in range(1024):
        d[s].copy_(h[s], non_blocking=True)
print("H2D 1024 x 1.5 MB copies GB/s", round(bw(rows, h.numel() * 8), 1))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        hb.copy_(db, non_blocking=True)
print("H2D + D2H concurrent     GB/s (H2D bytes only)", round(bw(both, h.numel() * 8), 1))
