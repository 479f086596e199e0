"""PCIe ceiling of this box: pinned D2H / H2D copies alone and both directions at once."""
import torch, time
dev = torch.device("cuda", 0)
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device=dev); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 8, dtype=torch.uint8, device=dev); h2 = torch.empty(n // 8, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 1 GiB pinned: {n / a / 1e9:.1f} GB/s")
b = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB pinned: {n / b / 1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
c = t(both); print(f"D2H 1 GiB + H2D 128 MiB concurrently: {n / c / 1e9:.1f} GB/s D2H")
