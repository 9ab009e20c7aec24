"""Is the 8-limb N = 16384 product sensitive to what ran before it (board power / clocks)?  Measure it cold, after 10 s of FP64-heavy
transforms, and after a 5 s pause, sampling the SM clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import primus_fhe_b200 as P
from bench import _c3_primes
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def clk(): return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
c3 = _c3_primes()
dc = P.U64DcrtTable(14, c3)
nrns = 1024
ra = torch.stack([torch.randint(0, m, (nrns, 16384), dtype=torch.int64, device="cuda") for m in c3], dim=1).contiguous()
rb, rc = ra.flip(0).contiguous(), torch.empty_like(ra)
def measure(tag):
    best = 1e9; seen = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dc.polymul_batch(ra, rb, rc); e1.record(); seen.append(clk()); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{tag}: {nrns / best * 1e3:.4e} RNS products/s; (SM MHz, W) while running: {seen[2:]}", flush=True)
measure("cold")
t = P.U64NttTable(12, 1125899906826241)
x = torch.randint(0, 1125899906826241, (65536, 4096), dtype=torch.int64, device="cuda")
t0 = time.time()
while time.time() - t0 < 10:
    for _ in range(50): t.forward_batch(x)
    torch.cuda.synchronize()
print("after 10 s of N = 4096 transforms:", clk())
measure("hot")
time.sleep(5)
measure("after 5 s pause")
