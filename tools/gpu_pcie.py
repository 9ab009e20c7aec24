"""PCIe ceiling probe: pinned H2D / D2H alone and simultaneously (the e2e host-slice path is bound by this)."""
import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
a = t(lambda: d1.copy_(h1, non_blocking=True)); print(f"H2D alone {n/a/1e9:.1f} GB/s")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H alone {n/b/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"H2D+D2H concurrent: {n/c/1e9:.1f} GB/s each direction")
