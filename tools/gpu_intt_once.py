"""Three inverse NTT launches of the headline shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, primus_fhe_b200 as P
t = P.U64NttTable(12, 1125899906826241)
x = torch.randint(0, 1125899906826241, (65536, 4096), dtype=torch.int64, device="cuda")
for _ in range(3):
    t.inverse_batch(x)
torch.cuda.synchronize()
