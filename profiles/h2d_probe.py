"""Pinned host<->device copy rates of the box (the ceiling of the e2e leg): 185 MB H2D, 36 MB D2H, alone and together."""
import time
import torch

h = torch.empty(185 * 2 ** 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
h2 = torch.empty(36 * 2 ** 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty_like(h2, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(both, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


run(False, 3)
t = run(False); print("H2D alone: %.2f ms per 185 MB = %.1f GB/s" % (t * 1e3, h.numel() / t / 1e9))
t = run(True); print("H2D with concurrent 36 MB D2H: %.2f ms = %.1f GB/s H2D" % (t * 1e3, h.numel() / t / 1e9))
