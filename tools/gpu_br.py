"""Blind-rotation / external-product timing helper (used under ncu and for tuning)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
which = sys.argv[1] if len(sys.argv) > 1 else "br"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1250
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
if which == "br":
    q, log_n, n_lwe = 132120577, 10, int(os.environ.get("NLWE", "512")); n = 1 << log_n
    t = P.U32NttTable(log_n, q)
    lv = P.ApproxSignedBasis(q, 7, None, 32).decompose_length()
    bsk = torch.randint(0, q, (n_lwe * 2 * lv * 2 * n,), dtype=torch.int64, device="cuda").to(torch.int32)
    lwe = torch.randint(0, 2 * n, (nb, n_lwe + 1), dtype=torch.int64, device="cuda").to(torch.int32)
    tv = torch.randint(0, q, (n,), dtype=torch.int64, device="cuda").to(torch.int32)
    acc = torch.empty((nb, 2 * n), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: t.blind_rotate_batch(7, None, bsk, n_lwe, lwe, tv, acc), reps=int(os.environ.get("REPS", "3")))
    print(f"blind rotate u32 N=1024 n={n_lwe} batch={nb}: {nb/ms*1e3:.4e} bootstraps/s ({ms:.2f} ms)")
    if os.environ.get("TERNARY"):
        bsk2 = torch.randint(0, q, (n_lwe * 2 * lv * 2 * n,), dtype=torch.int64, device="cuda").to(torch.int32)
        ms = timeit(lambda: t.blind_rotate_ternary_batch(7, None, bsk, bsk2, n_lwe, lwe, tv, acc), reps=3)
        print(f"ternary blind rotate u32 N=1024 n={n_lwe} batch={nb}: {nb/ms*1e3:.4e} bootstraps/s ({ms:.2f} ms)")
else:
    for bits, q in ((32, 132120577), (64, 1125899906826241)):
        t = (P.U64NttTable if bits == 64 else P.U32NttTable)(11, q); n = 2048
        lv = P.ApproxSignedBasis(q, 7, None, bits).decompose_length()
        dt = torch.int64 if bits == 64 else torch.int32
        key = torch.randint(0, q, (2 * lv * 2 * n,), dtype=torch.int64, device="cuda").to(dt)
        cin = torch.randint(0, q, (4096, 2 * n), dtype=torch.int64, device="cuda").to(dt)
        out = torch.empty_like(cin)
        ms = timeit(lambda: t.external_product_batch(1, 7, None, key, cin, out, True))
        print(f"extprod u{bits} N=2048 l={lv} batch=4096: {4096/ms*1e3:.4e} /s ({ms:.3f} ms)")
