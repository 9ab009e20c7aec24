import sys, os
sys.path.insert(0, "/root/repo")
import torch, primus_fhe_b200 as P
sys.path.insert(0, "/root/repo")
from bench import _c3_primes
c3 = _c3_primes()
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
dc = P.U64DcrtTable(14, c3)
t1 = P.U64NttTable(14, c3[0])
for nrns in (256, 1024):
    ra = torch.stack([torch.randint(0, m, (nrns, 16384), dtype=torch.int64, device="cuda") for m in c3], dim=1).contiguous()
    rb, rc = ra.flip(0).contiguous(), torch.empty_like(ra)
    ms = timeit(lambda: dc.polymul_batch(ra, rb, rc))
    print(f"DCRT polymul N=16384 L=8 batch {nrns}: {nrns/ms*1e3:.3e} RNS/s = {8*nrns/ms*1e3:.3e} limb-products/s")
    a1 = ra.view(-1, 16384); b1 = rb.view(-1, 16384); c1 = rc.view(-1, 16384)
    ms = timeit(lambda: t1.polymul_batch(a1, b1, c1))
    print(f"single-modulus polymul same shape ({8*nrns} polys): {8*nrns/ms*1e3:.3e} /s")
    ms = timeit(lambda: dc.forward_batch(ra)); print(f"DCRT fwd: {8*nrns/ms*1e3:.3e} limb NTT/s")
    ms = timeit(lambda: t1.forward_batch(a1)); print(f"single fwd: {8*nrns/ms*1e3:.3e} NTT/s")
    del ra, rb, rc

t13 = P.U64NttTable(13, 1125899906826241)
x = torch.randint(0, 1125899906826241, (16384, 8192), dtype=torch.int64, device="cuda"); y = x.flip(0).contiguous(); z = torch.empty_like(x)
ms = timeit(lambda: t13.polymul_batch(x, y, z)); print(f"polymul N=8192 u64 q50 batch 16384: {16384/ms*1e3:.3e} /s")
ms = timeit(lambda: t13.forward_batch(x)); print(f"fwd N=8192: {16384/ms*1e3:.3e} /s")
