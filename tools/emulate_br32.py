"""CPU emulation of lattice32.cu's blind-rotation schedule (thread/register/shared-memory level, numpy uint32 with
wrapping arithmetic) checked against the oracle -- a design aid that validates the index maps, twiddle indices,
padded exchange addresses, the carry-free digits, the Montgomery reduction and the 2^32 compensation without a GPU.

    python tools/emulate_br32.py [n_lwe] [seed]
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O

N, TPP, E1W, E2W = 1024, 128, 1152, 1152
U = np.uint32
M32 = (1 << 32) - 1

def u32(x): return np.asarray(x, dtype=np.uint64) & M32
def mulhi(a, b): return (np.asarray(a, np.uint64) * np.asarray(b, np.uint64)) >> np.uint64(32)
def shoup_lazy(y, w, wq, q): return u32(u32(np.asarray(y, np.uint64) * w) - u32(mulhi(y, wq) * q))
def umin(a, b): return np.minimum(u32(a), u32(b))

class Tab:
    def __init__(self, q):
        self.q = q
        t = O.U32NttTable(10, q)
        roots = np.array(t.roots(), dtype=np.uint64); inv_roots = np.array(t.inv_roots(), dtype=np.uint64)
        self.fwd = roots; self.fwd_q = np.array([(int(w) << 32) // q for w in roots], dtype=np.uint64)
        inv = inv_roots.copy()
        inv_n = pow(N, -1, q)
        inv[N - 1] = inv_n * int(inv_roots[N - 1]) % q
        self.inv = inv; self.inv_q = np.array([(int(w) << 32) // q for w in inv], dtype=np.uint64)
        r32 = (1 << 32) % q
        self.invn_r = inv_n * r32 % q; self.invn_r_q = (self.invn_r << 32) // q
        self.invnw_r = int(inv[N - 1]) * r32 % q; self.invnw_r_q = (self.invnw_r << 32) // q
        self.qinv = pow(q, -1, 1 << 32); self.one_q = (1 << 32) // q
        self.oracle = t

def bf_fwd(x, y, w, wq, q):
    t = shoup_lazy(y, w, wq, q)
    return u32(x + t), u32(x + 2 * q - t)
def bf_inv(x, y, w, wq, q):
    tx, ty = u32(x + y), u32(x + 2 * q - y)
    return umin(tx, tx - 2 * q), shoup_lazy(ty, w, wq, q)

def fwd_pass8(x, tw, T, idx):  # x: [128, 8]; tw(a, k) -> per-thread table index arrays
    q = T.q
    def W(i): return T.fwd[i], T.fwd_q[i]
    w, wq = W(idx(0, 0))
    for j in range(4): x[:, j], x[:, j + 4] = bf_fwd(x[:, j], x[:, j + 4], w, wq, q)
    for hi in range(2):
        w, wq = W(idx(1, hi))
        for j in range(2): x[:, 4 * hi + j], x[:, 4 * hi + j + 2] = bf_fwd(x[:, 4 * hi + j], x[:, 4 * hi + j + 2], w, wq, q)
    for k in range(4):
        w, wq = W(idx(2, k))
        x[:, 2 * k], x[:, 2 * k + 1] = bf_fwd(x[:, 2 * k], x[:, 2 * k + 1], w, wq, q)

def inv_pass8(x, T, idx):
    q = T.q
    def W(i): return T.inv[i], T.inv_q[i]
    for k in range(4):
        w, wq = W(idx(0, k))
        x[:, 2 * k], x[:, 2 * k + 1] = bf_inv(x[:, 2 * k], x[:, 2 * k + 1], w, wq, q)
    for hi in range(2):
        w, wq = W(idx(1, hi))
        for j in range(2): x[:, 4 * hi + j], x[:, 4 * hi + j + 2] = bf_inv(x[:, 4 * hi + j], x[:, 4 * hi + j + 2], w, wq, q)
    w, wq = W(idx(2, 0))
    for j in range(4): x[:, j], x[:, j + 4] = bf_inv(x[:, j], x[:, j + 4], w, wq, q)

t = np.arange(TPP); h = t >> 4; l = t & 15; g = t >> 1; beta = t & 1
zero = np.zeros(TPP, dtype=np.int64)

def forward(x, T):
    """x[t, j] = coefficient j*128 + t  ->  x[t, m] = output word 8t + m"""
    q = T.q
    e1 = np.zeros(E1W, dtype=np.uint64); e2 = np.zeros(E2W, dtype=np.uint64)
    fwd_pass8(x, None, T, lambda a, k: zero + [1, 2 + k, 4 + k][a])
    for j in range(8): e1[j * 144 + t] = x[:, j]
    for j in range(8): x[:, j] = e1[h * 144 + l + j * 16]
    fwd_pass8(x, None, T, lambda a, k: [8 + h, 16 + 2 * h + k, 32 + 4 * h + k][a])
    for j in range(8): e2[h * 144 + l + j * 18] = x[:, j]
    for j in range(8): x[:, j] = e2[18 * g + beta + 2 * j]
    fwd_pass8(x, None, T, lambda a, k: [64 + g, 128 + 2 * g + k, 256 + 4 * g + k][a])
    o = np.zeros_like(x)
    for k in range(4):
        send = np.where(beta == 1, x[:, k], x[:, k + 4])
        recv = send[t ^ 1]
        X = np.where(beta == 1, recv, x[:, k]); Y = np.where(beta == 1, x[:, k + 4], recv)
        i = 512 + 4 * t + k
        X, Y = bf_fwd(X, Y, T.fwd[i], T.fwd_q[i], q)
        o[:, 2 * k], o[:, 2 * k + 1] = X, Y
    return o

def inverse(x, T, bias):
    q = T.q
    e1 = np.zeros(E1W, dtype=np.uint64); e2 = np.zeros(E2W, dtype=np.uint64)
    o = np.zeros_like(x)
    for k in range(4):
        a, b = x[:, 2 * k], x[:, 2 * k + 1]
        i = 1 + 4 * t + k
        X = shoup_lazy(u32(a + b), 1, T.one_q, q)
        Y = shoup_lazy(u32(a + bias - b), T.inv[i], T.inv_q[i], q)
        send = np.where(beta == 1, X, Y); recv = send[t ^ 1]
        o[:, k] = np.where(beta == 1, recv, X); o[:, k + 4] = np.where(beta == 1, Y, recv)
    x = o
    base = lambda lg: 1 + N - (N >> lg)
    inv_pass8(x, T, lambda a, k: [base(1) + 4 * g + k, base(2) + 2 * g + k, base(3) + g][a])
    for j in range(8): e2[18 * g + beta + 2 * j] = x[:, j]
    for j in range(8): x[:, j] = e2[h * 144 + l + j * 18]
    inv_pass8(x, T, lambda a, k: [base(4) + 4 * h + k, base(5) + 2 * h + k, base(6) + h][a])
    for j in range(8): e1[h * 144 + l + j * 16] = x[:, j]
    for j in range(8): x[:, j] = e1[j * 144 + t]
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
    n_lwe = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for q, beta_log, lv_in in ((132120577, 7, None), (132120577, 4, 5), (134215681, 7, None), (132120577, 1, 6), (132120577, 9, None), (132120577, 2, 7)):
        T = Tab(q)
        ob = O.ApproxSignedBasis(q, beta_log, lv_in, 32); levels, drop = ob.decompose_length(), ob.drop_bits()
        terms = 2 * levels
        assert (2 * 10 + 2) * q < 1 << 32 and 2 * (terms + 1) * q < 1 << 32, "preconditions"
        rng = np.random.default_rng(seed)
        # transforms alone
        xin = rng.integers(0, q, N, dtype=np.uint64)
        x = np.zeros((TPP, 8), dtype=np.uint64)
        for j in range(8): x[:, j] = xin[j * 128 + t]
        out = forward(x, T)
        want = xin.astype(np.uint32).copy(); T.oracle.transform_slice(want)
        got = (out.reshape(-1) % q).astype(np.uint32)
        assert np.array_equal(got, want), "forward transform mismatch"
        # gadget constants as in launch_blind_rotate_fast32
        half = 0 if beta_log == 1 else 1 << (beta_log - 1)
        R = (1 << (drop - 1)) if drop else 0
        for lvl in range(levels): R += half << (drop + lvl * beta_log)
        thr = ob.threshold(); thr = 0xffffffff if thr is None else thr
        bits = q.bit_length(); add = (1 << bits) - q
        mask = (1 << beta_log) - 1; doff = q - half if half else 0
        bias = (terms + 1) * q
        bsk = rng.integers(0, q, n_lwe * 2 * levels * 2 * N, dtype=np.uint64).astype(np.uint32)
        lwe = rng.integers(0, 2 * N, (1, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
        lwe[0, 0] = 0; 
        if n_lwe > 1: lwe[0, 1] = N
        tv = rng.integers(0, q, N, dtype=np.uint64).astype(np.uint32)
        tv[:4] = 0
        want = O.blind_rotate(T.oracle, ob, bsk, n_lwe, lwe, tv, batch=1)[0]
        accs = np.zeros(2 * N, dtype=np.uint64)
        b = int(lwe[0, n_lwe]) & 2047; rot = (2 * N - b) & 2047
        for i in range(N):
            srcw = (i - rot) & 2047; v = int(tv[srcw & 1023])
            accs[N + i] = ((q - v) % q) if srcw >= N else v
        for i in range(n_lwe):
            a = int(lwe[0, i]) & 2047
            key = bsk[i * 2 * levels * 2 * N:].astype(np.uint64)
            W = np.zeros((2, TPP, 8), dtype=np.uint64)
            basei = (t - a) & 2047
            for r in range(2):
                for j in range(8):
                    src = (basei + 128 * j) & 2047
                    v = accs[r * N + (src & 1023)]; p = accs[r * N + j * 128 + t]
                    s = np.where(src & N, q - v, v)
                    d = u32(s + q - p); d = umin(d, d - q); d = umin(d, d - q)
                    W[r, :, j] = u32(d + np.where(d >= thr, add + R, R))
            acc = np.zeros((2, TPP, 8), dtype=object)
            for lvl in range(levels):
                shift = drop + lvl * beta_log
                for r in range(2):
                    x = ((W[r] >> np.uint64(shift)) & np.uint64(mask)) + np.uint64(doff)
                    out = forward(x.copy(), T)
                    assert int(out.max()) < 1 << 32
                    for c in range(2):
                        kp = key[((r * levels + lvl) * 2 + c) * N:][:N].reshape(TPP, 8)
                        acc[c] = acc[c] + out.astype(object) * kp.astype(object)
            for c in range(2):
                assert max(int(v) for v in acc[c].reshape(-1)) < 1 << 64
                lo = np.array([int(v) & M32 for v in acc[c].reshape(-1)], dtype=np.uint64).reshape(TPP, 8)
                hi = np.array([int(v) >> 32 for v in acc[c].reshape(-1)], dtype=np.uint64).reshape(TPP, 8)
                y = u32(hi + q - mulhi(u32(lo * T.qinv), q))
                assert int(y.max()) <= bias
                y = inverse(y, T, bias)
                for j in range(8):
                    s = u32(accs[c * N + j * 128 + t] + y[:, j])
                    accs[c * N + j * 128 + t] = umin(s, s - q)
        assert np.array_equal(accs.astype(np.uint32), want), f"blind rotation mismatch q={q} beta={beta_log}"
        print(f"q={q} log_basis={beta_log} levels={levels} drop={drop}: forward + {n_lwe}-step blind rotation == oracle")

if __name__ == "__main__":
    main()
