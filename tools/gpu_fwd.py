"""Forward-NTT timing helper for one size (used under ncu): python tools/gpu_fwd.py LOG_N BATCH [q]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
log_n, batch = int(sys.argv[1]), int(sys.argv[2])
q = int(sys.argv[3]) if len(sys.argv) > 3 else (1125899904679937 if log_n == 14 else 1125899906826241)
t = P.U64NttTable(log_n, q)
x = torch.randint(0, q, (batch, 1 << log_n), dtype=torch.int64, device="cuda")
for _ in range(3):
    t.forward_batch(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); t.forward_batch(x); e1.record(); torch.cuda.synchronize()
print(f"fwd N=2^{log_n} batch {batch}: {batch / e0.elapsed_time(e1) * 1e3:.3e} NTT/s")
