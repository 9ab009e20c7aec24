"""CPU emulation of the register-pass NTT index math (Plan / fill_pass_tables / fwd_pass_regs /
inv_pass_regs in primus_fhe_b200/csrc) against the oracle -- catches layout bugs without a GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O


class Plan:
    def __init__(s, logn, loge):
        s.logn, s.loge = logn, loge
        s.npass = (logn + loge - 1) // loge
        s.first = logn - (s.npass - 1) * loge
    def nstages(s, p): return s.first if p == 0 else s.loge
    def s0(s, p): return 0 if p == 0 else s.first + (p - 1) * s.loge
    def fb(s, p): return s.logn - s.loge if p == 0 else s.logn - (s.s0(p) + s.loge)
    def nh(s, p):
        tpp = 1 << (s.logn - s.loge)
        return max(tpp >> s.fb(p), 1)
    def entries(s, p): return ((1 << s.nstages(p)) - 1) * s.nh(p)
    def off(s, p): return sum(s.entries(i) for i in range(p))


def run(bits, logn, loge, q):
    t = (O.U64NttTable if bits == 64 else O.U32NttTable)(logn, q)
    n = 1 << logn; e = 1 << loge; tpp = n // e
    roots, inv_roots = [int(v) for v in t.roots()], [int(v) for v in t.inv_roots()]
    inv_n = t.inv_n(); inv_n_w = inv_n * inv_roots[n - 1] % q
    pl = Plan(logn, loge)
    total = pl.off(pl.npass)
    fwd, inv = [None] * total, [None] * total
    for p in range(pl.npass):
        ns, fb, nh, off, s0 = pl.nstages(p), pl.fb(p), pl.nh(p), pl.off(p), pl.s0(p)
        for ls in range(ns):
            jb = loge - 1 - ls; b = fb + jb
            inv_base = 1 + n - (n >> b)
            for jp in range(1 << ls):
                for high in range(nh):
                    blk = (high << ls) | jp
                    slot = off + ((1 << ls) - 1 + jp) * nh + high
                    fwd[slot] = roots[(1 << (s0 + ls)) + blk]
                    inv[slot] = inv_n_w if b == logn - 1 else inv_roots[inv_base + blk]
    assert all(v is not None for v in fwd)
    rng = np.random.default_rng(0)
    dt = np.uint64 if bits == 64 else np.uint32
    x = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    want = x.copy(); t.transform_slice(want)
    mem = [int(v) for v in x]

    def elem_index(fb, th, j):
        low, high = th & ((1 << fb) - 1), th >> fb
        return (high << (fb + loge)) | (j << fb) | low
    for p in range(pl.npass):
        ns, fb, nh, off = pl.nstages(p), pl.fb(p), pl.nh(p), pl.off(p)
        for th in range(tpp):
            xs = [mem[elem_index(fb, th, j)] for j in range(e)]
            high = th >> fb
            for ls in range(ns):
                jb = loge - 1 - ls
                for jp in range(1 << ls):
                    w = fwd[off + ((1 << ls) - 1 + jp) * nh + high]
                    for jl in range(1 << jb):
                        j0 = (jp << (jb + 1)) | jl; j1 = j0 | (1 << jb)
                        u, v = xs[j0], xs[j1] * w % q
                        xs[j0], xs[j1] = (u + v) % q, (u - v) % q
            for j in range(e): mem[elem_index(fb, th, j)] = xs[j]
    assert mem == [int(v) for v in want], ("fwd mismatch", bits, logn, loge)
    for p in range(pl.npass - 1, -1, -1):
        ns, fb, nh, off = pl.nstages(p), pl.fb(p), pl.nh(p), pl.off(p)
        for th in range(tpp):
            xs = [mem[elem_index(fb, th, j)] for j in range(e)]
            high = th >> fb
            for ls in range(ns - 1, -1, -1):
                jb = loge - 1 - ls
                for jp in range(1 << ls):
                    w = inv[off + ((1 << ls) - 1 + jp) * nh + high]
                    for jl in range(1 << jb):
                        j0 = (jp << (jb + 1)) | jl; j1 = j0 | (1 << jb)
                        u, v = xs[j0], xs[j1]
                        if p == 0 and ls == 0:
                            xs[j0], xs[j1] = (u + v) * inv_n % q, (u - v) * w % q
                        else:
                            xs[j0], xs[j1] = (u + v) % q, (u - v) * w % q
            for j in range(e): mem[elem_index(fb, th, j)] = xs[j]
    assert mem == [int(v) for v in x], ("inv mismatch", bits, logn, loge)


if __name__ == "__main__":
    for bits, logn, loge, q in [(64, 6, 2, 1125899906826241), (64, 7, 3, 1125899906826241), (64, 10, 5, 1125899906826241), (64, 11, 4, 1125899906826241),
                                (64, 12, 4, 1125899906826241), (64, 13, 5, 1125899906826241), (64, 11, 3, 1125899906826241), (32, 10, 5, 132120577),
                                (32, 11, 6, 132120577), (32, 12, 6, 132120577), (32, 10, 4, 132120577), (32, 13, 5, 132120577), (64, 10, 3, 1125899906826241)]:
        run(bits, logn, loge, q); print("ok", bits, logn, loge)
