"""Tuning experiment for the N=4096 u64 forward NTT: LOGE / PPB variants (env hooks in ntt.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
q, log_n, n, batch = 1125899906826241, 12, 4096, 65536
x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
def timeit(fn, reps=6):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for loge, ppb in [("4", "1"), ("3", "1"), ("5", "1"), ("5", "2")]:
    os.environ["PFHE_LOGE12"] = loge; os.environ["PFHE_PPB12"] = ppb
    t = P.U64NttTable(log_n, q)
    y = x.clone(); t.forward_batch(y); t.inverse_batch(y)
    ok = bool((y == x).all())
    msf = timeit(lambda: t.forward_batch(x)); msi = timeit(lambda: t.inverse_batch(x))
    print(f"LOGE={loge} PPB={ppb}: fwd {batch/msf*1e3:.3e}  inv {batch/msi*1e3:.3e} NTT/s roundtrip_ok={ok}")
