"""Throughput table of the NTT family (fresh process per PFHE_F64_LAZY setting).  Not the bench: quick numbers to steer kernel work."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
import torch
import primus_fhe_b200 as P
def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for bits, log_n, q, batch in json.loads(os.environ["AB_CASES"]):
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    n = 1 << log_n
    x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
    if bits == 32: x = x.to(torch.int32)
    x0 = x.clone()
    t.forward_batch(x); t.inverse_batch(x); torch.cuda.synchronize()
    ok = bool(torch.equal(x, x0))
    msf = timeit(lambda: t.forward_batch(x)); msi = timeit(lambda: t.inverse_batch(x))
    y = x.clone(); z = torch.empty_like(x)
    msp = timeit(lambda: t.polymul_batch(x, y, z))
    out[f"u{bits}_q{int(q).bit_length()}_n{n}"] = dict(fwd=batch / msf * 1e3, inv=batch / msi * 1e3, polymul=batch / msp * 1e3, roundtrip_ok=ok)
print("AB_RESULT " + json.dumps(out))
''' % ROOT
CASES = [(64, 12, 1125899906826241, 65536), (64, 12, 1152921504606830593, 65536), (64, 13, 1125899906826241, 32768),
         (64, 14, 1125899904679937, 16384), (64, 11, 1125899906826241, 131072), (64, 10, 1125899906826241, 262144),
         (32, 10, 132120577, 262144), (32, 11, 132120577, 131072), (32, 12, 268369921, 65536)]
res = {}
for lazy in ((1,) if "quick" in sys.argv else (0, 1)):
    env = dict(os.environ, PFHE_F64_LAZY=str(lazy), AB_CASES=json.dumps(CASES))
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("AB_RESULT ")]
    if not line:
        print(f"lazy={lazy} FAILED\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}")
        continue
    r = json.loads(line[0][10:])
    res[f"lazy{lazy}"] = r
    for k, v in r.items():
        w = 8 if k.startswith("u64") else 4
        n = int(k.split("_n")[1])
        print(f"lazy={lazy} {k}: fwd {v['fwd']:.3e} ({v['fwd']*2*n*w/1e9:.0f} GB/s) inv {v['inv']:.3e} polymul {v['polymul']:.3e} ({v['polymul']*3*n*w/1e9:.0f} GB/s) rt_ok={v['roundtrip_ok']}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab.json"), "w"), indent=1)
