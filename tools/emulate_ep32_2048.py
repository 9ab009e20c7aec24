"""CPU emulation (numpy, wrapping uint32) of lattice32_ep.cu's N = 2048 schedule: 3 + 3 + 3 + 2 register passes, exchange-buffer
addresses (plain / idx + 4*(idx>>5)), twiddle indices, carry-free digits, Montgomery reduction, NTT-domain and coefficient
outputs -- checked against the oracle's external product.  Design aid that runs without a GPU (tests/test_oracle.py runs it).

    python tools/emulate_ep32_2048.py [seed]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O
from emulate_br32 import u32, mulhi, shoup_lazy, umin, bf_fwd, bf_inv, fwd_pass8, inv_pass8

LOGN, N, TPP, F1W, F2W = 11, 2048, 256, 2048, 36 * 64
t = np.arange(TPP); h = t >> 5; l = t & 31; g = t >> 2; lam = t & 3
zero = np.zeros(TPP, dtype=np.int64)


class Tab:
    def __init__(self, q):
        self.q = q
        tb = O.U32NttTable(LOGN, q)
        roots = np.array(tb.roots(), dtype=np.uint64); inv_roots = np.array(tb.inv_roots(), dtype=np.uint64)
        self.fwd = roots; self.fwd_q = np.array([(int(w) << 32) // q for w in roots], dtype=np.uint64)
        inv = inv_roots.copy(); inv_n = pow(N, -1, q)
        inv[N - 1] = inv_n * int(inv_roots[N - 1]) % q
        self.inv = inv; self.inv_q = np.array([(int(w) << 32) // q for w in inv], dtype=np.uint64)
        r32 = (1 << 32) % q
        self.r32, self.r32_q = r32, (r32 << 32) // q
        self.invn_r = inv_n * r32 % q; self.invn_r_q = (self.invn_r << 32) // q
        self.invnw_r = int(inv[N - 1]) * r32 % q; self.invnw_r_q = (self.invnw_r << 32) // q
        self.qinv = pow(q, -1, 1 << 32); self.one_q = (1 << 32) // q
        self.oracle = tb


def forward(x, T):
    q = T.q
    f1 = np.zeros(F1W, dtype=np.uint64); f2 = np.zeros(F2W, dtype=np.uint64)
    fwd_pass8(x, None, T, lambda a, k: zero + [1, 2 + k, 4 + k][a])
    for j in range(8): f1[j * 256 + t] = x[:, j]
    for j in range(8): x[:, j] = f1[h * 256 + l + j * 32]
    fwd_pass8(x, None, T, lambda a, k: [8 + h, 16 + 2 * h + k, 32 + 4 * h + k][a])
    for j in range(8): f2[h * 288 + l + j * 36] = x[:, j]
    for j in range(8): x[:, j] = f2[36 * g + lam + 4 * j]
    fwd_pass8(x, None, T, lambda a, k: [64 + g, 128 + 2 * g + k, 256 + 4 * g + k][a])
    for j in range(8): f2[36 * g + lam + 4 * j] = x[:, j]
    for m in range(8): x[:, m] = f2[36 * g + 8 * lam + m]
    W = lambda i: (T.fwd[i], T.fwd_q[i])
    for base in (0, 4):
        w = W(512 + 2 * t + base // 4)
        for m in (0, 1): x[:, base + m], x[:, base + m + 2] = bf_fwd(x[:, base + m], x[:, base + m + 2], *w, q)
    for k in range(4):
        x[:, 2 * k], x[:, 2 * k + 1] = bf_fwd(x[:, 2 * k], x[:, 2 * k + 1], *W(1024 + 4 * t + k), q)
    return x


def inverse(x, T, bias):
    q = T.q
    f1 = np.zeros(F1W, dtype=np.uint64); f2 = np.zeros(F2W, dtype=np.uint64)
    base = lambda lg: 1 + N - (N >> lg)
    W = lambda i: (T.inv[i], T.inv_q[i])
    for k in range(4):
        a, b = x[:, 2 * k].copy(), x[:, 2 * k + 1].copy()
        x[:, 2 * k] = shoup_lazy(u32(a + b), 1, T.one_q, q)
        x[:, 2 * k + 1] = shoup_lazy(u32(a + bias - b), *W(1 + 4 * t + k), q)
    for half in (0, 1):
        w = W(base(1) + 2 * t + half)
        for m in (0, 1): x[:, 4 * half + m], x[:, 4 * half + m + 2] = bf_inv(x[:, 4 * half + m], x[:, 4 * half + m + 2], *w, q)
    for m in range(8): f2[36 * g + 8 * lam + m] = x[:, m]
    for j in range(8): x[:, j] = f2[36 * g + lam + 4 * j]
    inv_pass8(x, T, lambda a, k: [base(2) + 4 * g + k, base(3) + 2 * g + k, base(4) + g][a])
    for j in range(8): f2[36 * g + lam + 4 * j] = x[:, j]
    for j in range(8): x[:, j] = f2[h * 288 + l + j * 36]
    inv_pass8(x, T, lambda a, k: [base(5) + 4 * h + k, base(6) + 2 * h + k, base(7) + h][a])
    for j in range(8): f1[h * 256 + l + j * 32] = x[:, j]
    for j in range(8): x[:, j] = f1[j * 256 + t]
    tail = lambda k: (T.inv[N - 8 + k], T.inv_q[N - 8 + k])
    for k in range(4): x[:, 2 * k], x[:, 2 * k + 1] = bf_inv(x[:, 2 * k], x[:, 2 * k + 1], *tail(1 + k), q)
    for hi in range(2):
        for j in range(2): x[:, 4 * hi + j], x[:, 4 * hi + j + 2] = bf_inv(x[:, 4 * hi + j], x[:, 4 * hi + j + 2], *tail(5 + hi), q)
    for j in range(4):
        tx, ty = u32(x[:, j] + x[:, j + 4]), u32(x[:, j] + 2 * q - x[:, j + 4])
        a = shoup_lazy(tx, T.invn_r, T.invn_r_q, q); b = shoup_lazy(ty, T.invnw_r, T.invnw_r_q, q)
        x[:, j], x[:, j + 4] = umin(a, a - q), umin(b, b - q)
    return x


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    for q, beta_log, lv_in in ((132120577, 7, None), (132120577, 4, 5), (132120577, 9, None)):
        T = Tab(q)
        ob = O.ApproxSignedBasis(q, beta_log, lv_in, 32); levels, drop = ob.decompose_length(), ob.drop_bits()
        terms = 2 * levels
        assert (2 * LOGN + 2) * q < 1 << 32 and 2 * (terms + 1) * q < 1 << 32
        rng = np.random.default_rng(seed)
        xin = rng.integers(0, q, N, dtype=np.uint64)
        x = np.zeros((TPP, 8), dtype=np.uint64)
        for j in range(8): x[:, j] = xin[j * 256 + t]
        out = forward(x, T)
        want = xin.astype(np.uint32).copy(); T.oracle.transform_slice(want)
        assert np.array_equal((out.reshape(-1) % q).astype(np.uint32), want), "forward transform mismatch"
        half = 0 if beta_log == 1 else 1 << (beta_log - 1)
        R = (1 << (drop - 1)) if drop else 0
        for lvl in range(levels): R += half << (drop + lvl * beta_log)
        thr = ob.threshold(); thr = 0xffffffff if thr is None else thr
        add = (1 << q.bit_length()) - q; mask = (1 << beta_log) - 1; doff = q - half if half else 0
        bias = (terms + 1) * q
        key = rng.integers(0, q, 2 * levels * 2 * N, dtype=np.uint64).astype(np.uint32)
        cin = rng.integers(0, q, (1, 2 * N), dtype=np.uint64).astype(np.uint32)
        cin[0, :5] = (0, q - 1, 1, q // 2, q // 2 + 1)
        for to_coeff in (True, False):
            want = O.external_product_single(T.oracle, ob, 1, key, cin, to_coeff=to_coeff, batch=1)[0]
            got = np.zeros(2 * N, dtype=np.uint64)
            acc = np.zeros((2, TPP, 8), dtype=object)
            Wd = np.zeros((2, TPP, 8), dtype=np.uint64)
            for r in range(2):
                for j in range(8):
                    d = cin[0, r * N + j * 256 + t].astype(np.uint64)
                    Wd[r, :, j] = u32(d + np.where(d >= thr, add + R, R))
            for lvl in range(levels):
                shift = drop + lvl * beta_log
                for r in range(2):
                    x = ((Wd[r] >> np.uint64(shift)) & np.uint64(mask)) + np.uint64(doff)
                    o = forward(x.copy(), T)
                    assert int(o.max()) < 1 << 32
                    for c in range(2):
                        kp = key[((r * levels + lvl) * 2 + c) * N:][:N].astype(np.uint64).reshape(TPP, 8)
                        acc[c] = acc[c] + o.astype(object) * kp.astype(object)
            for c in range(2):
                lo = np.array([int(v) & 0xffffffff for v in acc[c].reshape(-1)], dtype=np.uint64).reshape(TPP, 8)
                hi = np.array([int(v) >> 32 for v in acc[c].reshape(-1)], dtype=np.uint64).reshape(TPP, 8)
                y = u32(hi + q - mulhi(u32(lo * T.qinv), q))
                assert int(y.max()) <= bias
                if to_coeff:
                    y = inverse(y, T, bias)
                    for j in range(8): got[c * N + j * 256 + t] = y[:, j]
                else:
                    v = shoup_lazy(y, T.r32, T.r32_q, q)
                    got[c * N:(c + 1) * N] = umin(v, v - q).reshape(-1)
            assert np.array_equal(got.astype(np.uint32), want), (q, beta_log, to_coeff)
        print(f"N=2048 q={q} log_basis={beta_log} levels={levels}: forward + external product (coeff and NTT output) == oracle")


if __name__ == "__main__":
    main()
