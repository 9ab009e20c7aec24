"""Host-slice shim on PAGEABLE memory (numpy buffer = what a Rust Vec<u64> is): staged path vs the driver's own staging."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import primus_fhe_b200 as P
q, n, batch = 1125899906826241, 4096, 32768
t = P.U64NttTable(12, q)
rng = np.random.default_rng(1)
host = rng.integers(0, q, (batch, n), dtype=np.uint64)
ref = host[:4].copy()
t.transform_slices(host)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); t.transform_slices(host); best = min(best, time.perf_counter() - t0)
print(f"PFHE_STAGE={os.environ.get('PFHE_STAGE','1')} threads={os.environ.get('PFHE_STAGE_THREADS','4')} chunk={os.environ.get('PFHE_PIPE_CHUNK_MB','32')}MB: "
      f"pageable e2e {batch/best:.4e} NTT/s ({batch*n*8/best/1e9:.1f} GB/s per direction)")
