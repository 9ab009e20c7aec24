"""A/B of environment tuning hooks on the headline kernel (fresh process per setting): python tools/gpu_ab_env.py "" A=1 A=1,B=2"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys; sys.path.insert(0, %r)
import torch, numpy as np, primus_fhe_b200 as P
from oracle import oracle as O
q, batch = 1125899906826241, 65536
t = P.U64NttTable(12, q)
x = torch.randint(0, q, (batch, 4096), dtype=torch.int64, device="cuda")
x0 = x.clone()
def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ot = O.U64NttTable(12, q)
want = x0[:4].cpu().numpy().view(np.uint64).copy(); ot.forward_batch(want)
t.forward_batch(x); exact = np.array_equal(x[:4].cpu().numpy().view(np.uint64), want)
t.inverse_batch(x); ok = torch.equal(x, x0)
y = torch.empty_like(x); t.forward_batch_to(x, y); exact2 = np.array_equal(y[:4].cpu().numpy().view(np.uint64), want)
f = timeit(lambda: t.forward_batch(x)); i = timeit(lambda: t.inverse_batch(x))
print("RES fwd %%.3e inv %%.3e fwd_exact=%%s/%%s roundtrip=%%s" %% (batch/f*1e3, batch/i*1e3, exact, exact2, ok))
''' % ROOT
for st in [dict(a.split("=") for a in s.split(",")) if s else {} for s in (sys.argv[1:] or [""])]:
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **st), capture_output=True, text=True)
    out = [l for l in p.stdout.splitlines() if l.startswith("RES")]
    print(st, out[0] if out else "FAILED " + p.stderr[-1500:])
