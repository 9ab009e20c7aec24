import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
q, n, batch = 1125899906826241, 4096, 65536
t = P.U64NttTable(12, q)
host = torch.randint(0, q, (batch, n), dtype=torch.int64).pin_memory()
t.transform_slices(host)
best = 1e9
for _ in range(4):
    t0 = time.perf_counter(); t.transform_slices(host); best = min(best, time.perf_counter() - t0)
dt = best
print(f"chunk={os.environ.get('PFHE_PIPE_CHUNK_MB','64')}MB: e2e {batch/dt:.4e} NTT/s ({2*batch*n*8/dt/1e9:.1f} GB/s both ways)")
