"""GB/s of the RNS streaming kernels (compose, decompose, multi-word gadget, base conversion, scaled lift) at operands larger than L2.
Every timed output is compared with the oracle on a leading slice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import primus_fhe_b200 as P
from oracle import oracle as O
PEAK = 6436.4
Q, QB, QC = 1125899906826241, 1125899906629633, 1125899905744897
g = torch.Generator(device="cuda"); g.manual_seed(3)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


def u64(t): return t.cpu().numpy().view(np.uint64)
def rnd(m, shape): return torch.randint(0, m, shape, dtype=torch.int64, device="cuda", generator=g)
def show(name, nbytes, dt, ok): print(f"{name:44s} {nbytes / dt / 1e9:8.1f} GB/s  {nbytes / dt / 1e9 / PEAK:5.2f} of HBM copy peak  parity={ok}", flush=True)


polys, n = 16384, 2048                      # 32 Mi coefficients: 256 MiB per u64 limb
count = polys * n
for mods in ([Q, QB], [Q, QB, QC]):
    L = len(mods)
    rns = P.RNSBase(mods, 64); orns = O.RNSBase(mods, 64)
    vl = rns.big_uint_value_len()
    flat = torch.stack([rnd(m, (count,)) for m in mods], dim=0).contiguous()
    big = torch.empty((count, vl), dtype=torch.int64, device="cuda")
    dt = timeit(lambda: rns.compose_multiple_values_to(flat, big))
    want = orns.compose_multiple_values_to(u64(flat[:, :4096]).copy().reshape(-1), 4096)
    show(f"rns_compose L={L} (value_len {vl})", count * 8 * (L + vl), dt, bool(np.array_equal(u64(big[:4096]).reshape(-1), np.asarray(want).reshape(-1))))
    back = torch.empty_like(flat)
    dt = timeit(lambda: rns.decompose_big_uint_values_to(big, back))
    show(f"rns_decompose L={L}", count * 8 * (L + vl), dt, bool(torch.equal(back, flat)))
    bb = P.BigUintApproxSignedBasis(rns, 7, None); lv = bb.decompose_length()
    gp = 2048 if L == 2 else 1024               # digits are l*L words per coefficient
    res = torch.stack([rnd(m, (gp, n)) for m in mods], dim=1).contiguous()
    digs = torch.empty((gp, lv, L, n), dtype=torch.int64, device="cuda")
    dt = timeit(lambda: bb.gadget_decompose_batch(res, digs, n))
    obb = O.BigUintApproxSignedBasis(orns, 7, None)
    ob = orns.compose_multiple_values_to(u64(res[0]).copy().reshape(-1), n)
    car = obb.init_value_carry_slice_inplace(ob)
    want = np.stack([orns.wrapping_decompose_small_values_to(obb.unsigned_decompose_slice_to(l, ob, car), 1 << 7).reshape(L, n) for l in range(lv)])
    ok = bool(np.array_equal(u64(digs[0]).reshape(-1), want.reshape(-1)))
    show(f"rns_gadget L={L} l={lv}", gp * n * 8 * (L + L * lv), dt, ok)
    del flat, big, back, res, digs
    small = rnd(1 << 20, (count,))
    acc = torch.stack([rnd(m, (count,)) for m in mods], dim=0).contiguous()
    scal = [12345 % m for m in mods]
    dt = timeit(lambda: rns.wrapping_decompose_small_values_scaled_add_to(small, acc, 1 << 20, scal))
    show(f"rns_lift_scaled_acc L={L}", count * 8 * (1 + 2 * L), dt, None)
    del small, acc

for ins, outs in (([137438822401, 137438814209, 137438773249], [Q, QB]), ([Q, QB, QC], [1125899905351681, 562949953392641]),
                  ([Q, 1152921504606830593], [QB, QC])):
    bits = max(int(m).bit_length() for m in ins + outs)
    cin = torch.stack([rnd(m, (polys, n)) for m in ins], dim=1).contiguous()
    for exact in (False, True):
        om = outs[:1] if exact else outs
        bc = P.BaseConverter(ins, om, 64); obc = O.BaseConverter(ins, om, 64)
        cout = torch.empty((polys, len(om), n), dtype=torch.int64, device="cuda")
        run = (lambda: bc.exact_convert_array(cin, cout, n)) if exact else (lambda: bc.fast_convert_array(cin, cout, n))
        dt = timeit(run)
        ofn = obc.exact_convert_array if exact else obc.fast_convert_array
        want = np.stack([np.asarray(ofn(u64(cin[p]).copy().reshape(-1), n)).reshape(len(om), n) for p in (0, polys - 1)])
        got = np.stack([u64(cout[p]) for p in (0, polys - 1)])
        show(f"baseconv {'exact' if exact else 'fast'} {len(ins)}->{len(om)} ({bits}-bit moduli)", count * 8 * (len(ins) + len(om)), dt,
             bool(np.array_equal(got, want)))
        del cout
    del cin
