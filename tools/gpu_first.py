"""First-contact GPU script: integer-pipe microbenchmark + quick throughput numbers (not the bench)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import primus_fhe_b200 as P

res = {}
for kind, name in ((0, "u32"), (1, "u64")):
    blocks, iters = 148 * 8, 4096
    ms = P.modmul_microbench(kind, blocks, iters)
    bfly = blocks * 256 * 8 * iters
    res[f"bfly_{name}_per_s"] = bfly / (ms * 1e-3)
    print(name, "butterflies/s", f"{bfly/(ms*1e-3):.3e}", "ms", ms)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for bits, log_n, q, batch in [(64, 12, 1125899906826241, 65536), (64, 13, 1125899906826241, 32768), (64, 14, 1125899904679937 if False else 1125899906826241, 0),
                              (64, 11, 1125899906826241, 131072), (64, 10, 1125899906826241, 262144), (32, 10, 132120577, 262144), (32, 11, 132120577, 131072), (32, 12, 268369921, 65536)]:
    if batch == 0: continue
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    n = 1 << log_n
    x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
    if bits == 32: x = x.to(torch.int32)
    msf = timeit(lambda: t.forward_batch(x))
    msi = timeit(lambda: t.inverse_batch(x))
    y = x.clone(); z = torch.empty_like(x)
    msp = timeit(lambda: t.polymul_batch(x, y, z))
    by = 2 * batch * n * (bits // 8)
    print(f"u{bits} N={n} batch={batch}: fwd {batch/msf*1e3:.3e} NTT/s ({by/msf/1e6:.0f} GB/s)  inv {batch/msi*1e3:.3e}  polymul {batch/msp*1e3:.3e}/s ({1.5*by/msp/1e6:.0f} GB/s)")
    res[f"fwd_u{bits}_n{n}"] = batch / msf * 1e3
    res[f"inv_u{bits}_n{n}"] = batch / msi * 1e3
    res[f"polymul_u{bits}_n{n}"] = batch / msp * 1e3

# external product C4-A / C4-B, blind rotation C5 (small n_lwe sample)
for bits, q, log_n, batch in [(32, 132120577, 11, 4096), (64, 1125899906826241, 11, 4096)]:
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    n = 1 << log_n
    b = P.ApproxSignedBasis(q, 7, None, bits)
    lv = b.decompose_length()
    dt = torch.int64 if bits == 64 else torch.int32
    key = torch.randint(0, q, (2 * lv * 2 * n,), dtype=torch.int64, device="cuda").to(dt)
    cin = torch.randint(0, q, (batch, 2 * n), dtype=torch.int64, device="cuda").to(dt)
    out = torch.empty_like(cin)
    ms = timeit(lambda: t.external_product_batch(1, 7, None, key, cin, out, True))
    print(f"extprod u{bits} N={n} l={lv} batch={batch}: {batch/ms*1e3:.3e} /s")
    res[f"extprod_u{bits}"] = batch / ms * 1e3
q, log_n, n_lwe, batch = 132120577, 10, 512, 2048
t = P.U32NttTable(log_n, q); n = 1 << log_n
lv = P.ApproxSignedBasis(q, 7, None, 32).decompose_length()
bsk = torch.randint(0, q, (n_lwe * 2 * lv * 2 * n,), dtype=torch.int64, device="cuda").to(torch.int32)
lwe = torch.randint(0, 2 * n, (batch, n_lwe + 1), dtype=torch.int64, device="cuda").to(torch.int32)
tv = torch.randint(0, q, (n,), dtype=torch.int64, device="cuda").to(torch.int32)
acc = torch.empty((batch, 2 * n), dtype=torch.int32, device="cuda")
ms = timeit(lambda: t.blind_rotate_batch(7, None, bsk, n_lwe, lwe, tv, acc), reps=2)
print(f"blind rotate u32 N=1024 n=512 batch={batch}: {batch/ms*1e3:.3e} /s ({ms:.1f} ms)")
res["blind_rotate_u32"] = batch / ms * 1e3
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/first.json", "w"), indent=1)
