"""N = 8192 (u64, q50) forward / inverse / fused product rates, batch 16384 (used for the cluster A/B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
q = 1125899906826241
t = P.U64NttTable(13, q)
x = torch.randint(0, q, (16384, 8192), dtype=torch.int64, device="cuda"); y = x.flip(0).contiguous(); z = torch.empty_like(x)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = timeit(lambda: t.polymul_batch(x, y, z)); print(f"polymul N=8192: {16384/ms*1e3:.4e} /s")
ms = timeit(lambda: t.forward_batch(x)); print(f"fwd N=8192: {16384/ms*1e3:.4e} /s")
ms = timeit(lambda: t.inverse_batch(x)); print(f"inv N=8192: {16384/ms*1e3:.4e} /s")
