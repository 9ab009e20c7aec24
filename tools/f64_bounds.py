"""Exactness budget of the lazy FP64-pipe butterflies (primus_fhe_b200/csrc/ntt_core.cuh, struct F64LazyField).

Every value is an integer held in a double.  One modular product mulmod_c(y, w), 0 <= w < q:
    h = RN(y*w); l = fma(y, w, -h)            (exact split P = h + l)
    t = fma(h, qinv, M_c), M_c = 1.5*2^(52+c) (one rounding to a multiple of 2^c, valid while |h*qinv| < 2^(51+c))
    k = t - M_c;  r = fma(-k, q, h) + l       (exact while the intermediate integers stay below 2^53)
    |k - P/q| <= 2^(c-1) + |P/q| * (2^-53 + 2^-53 + 2^-106)   =>   |r| <= q * (2^(c-1) + |y| * 2^-52 * (1 + 2^-53))
fold(x):  k = rint(x*qinv) by the same trick (c = 0), x - k q:  |result| <= q/2 + |x| * 2^-52 * ... < q/2 + 1.
This script walks the stage schedules used by the kernels with exact rationals for the worst admissible modulus
and asserts every precondition (k range, integer exactness of sums).  Run: python tools/f64_bounds.py
"""
from fractions import Fraction as Fr

QMAX = (1 << 50) - 1024          # use_f64 requires q <= 2^50 - 2^10 (capi.cu)
EPS = Fr(1, 1 << 52) * (1 + Fr(1, 1 << 53))
LIMIT = Fr(1 << 53)              # integers up to 2^53 are exact in a double

FWD_LEVEL = [0, 0, 0, 1, 2]      # magic level by stage-since-fold (forward, CT)
INV_LEVEL = [0, 1, 2, 3]         # inverse (GS): the sum path doubles every stage, so the level follows it


def mul_bound(b, c, q):
    # precondition of the magic-constant rounding
    assert b * (1 + Fr(1, 1 << 52)) < Fr(1 << (51 + c)), ("k range", float(b / q), c)
    t = q * (Fr(1 << c, 2) + b * EPS)
    # h - k q = r - l must be an exact integer double: |r| + |l| < 2^53, |l| <= ulp(P)/2 <= |P| 2^-53
    assert t + b * q * Fr(1, 1 << 53) < LIMIT
    return t


def fold_bound(b, q):
    assert b <= LIMIT
    return q / 2 + b * EPS + 1


def forward(stages_per_pass, q):
    b = (q + 1) / 2              # centred load (exact 64-bit compare): [0,q) -> [-(q-1)/2, (q+1)/2)
    worst = b
    for p, ns in enumerate(stages_per_pass):
        if p:
            b = fold_bound(b, q)
        for s in range(ns):
            t = mul_bound(b, FWD_LEVEL[s], q)
            b = b + t
            assert b <= LIMIT, ("fwd sum", p, s, float(b / q))
            worst = max(worst, b)
    out = fold_bound(b, q)       # canonical output: fold, + q bias, one conditional subtract
    assert out < q
    return float(worst / q)


def inverse(stages_per_pass, q):
    # stages_per_pass in execution order (last pass first); the final stage multiplies both outputs (n^-1 fused)
    b = (q + 1) / 2
    worst = b
    total = sum(stages_per_pass)
    done = 0
    for p, ns in enumerate(stages_per_pass):
        if p:
            b = fold_bound(b, q)
        since = 0
        for s in range(ns):
            if since == len(INV_LEVEL):
                b = fold_bound(b, q)
                since = 0
            sd = 2 * b           # s = x + y, d = x - y
            assert sd <= LIMIT, ("inv sum", p, s, float(sd / q))
            t = mul_bound(sd, INV_LEVEL[since], q)
            done += 1
            b = t if done == total else max(sd, t)
            worst = max(worst, sd)
            since += 1
    out = fold_bound(b, q)
    assert out < q
    return float(worst / q)


def mac_budget(q, every=8):
    """Lattice key multiply-accumulate (LatAcc<F64LazyField>): digit folded to q/2+1, products at level 0, the accumulator
    folded after `every` terms (the first batch starts from 0, later ones from a folded accumulator counted as one term)."""
    x = fold_bound(Fr(8) * q, q)                 # |transformed digit| after the fold
    t = mul_bound(x, 0, q)                       # |x * key mod q| with key < q
    first = every * t
    later = fold_bound(first, q) + (every - 1) * t
    assert first <= LIMIT and later <= LIMIT
    return float(max(first, later) / q)


def canonical_product_budget(qs_in, ps_out):
    """Round-2 uses of the same product outside the transforms (rns.cu baseconv_kernel<F64>, pointwise.cu f64_mul_canonical):
    operands are CANONICAL residues (0 <= y < q_i <= QMAX, multiplier below the modulus of the product), level 0.
      base conversion:  y_i = x_i * inv_i mod q_i   (|result| < q_i, made canonical by one conditional + q_i)
                        out_k = sum_i y_i * M[k][i] mod p_k,  y_i < q_i (NOT reduced mod p_k), M < p_k: at most 8 terms, then one fold
                        exact variant: minus v * (Q mod p_k), v <= 8
      reduce_mul:       a * b mod q with a, b < q, result + q in (0, 2q), one conditional subtraction."""
    worst = Fr(0)
    for q in qs_in:
        t = mul_bound(Fr(q), 0, Fr(q))           # x < q (non-canonical words are reduced first), inv < q
        assert t < q                              # so  v < 0 ? v + q : v  is the canonical residue
        worst = max(worst, t / q)
    for p in ps_out:
        acc = Fr(0)
        for q in qs_in[:8]:
            acc += mul_bound(Fr(q), 0, Fr(p))     # |y_i * M mod p| with y_i < q_i <= QMAX
        acc += mul_bound(Fr(8), 0, Fr(p))         # exact variant: v * (Q mod p), v <= number of input limbs
        assert acc <= LIMIT                        # the sum is exact
        assert fold_bound(acc, Fr(p)) < p          # canon(): fold, + p, one conditional subtraction
        worst = max(worst, acc / p)
    for q in set(qs_in) | set(ps_out):
        t = mul_bound(Fr(q), 0, Fr(q))
        assert t < q and t + q < 2 * q             # f64_mul_canonical: r + q in (0, 2q)
    return float(worst)


def plan(logn, loge):
    npass = (logn + loge - 1) // loge
    first = logn - (npass - 1) * loge
    return [first] + [loge] * (npass - 1)


def check_all():
    res = {}
    for q in (QMAX, (1 << 49) + 1, 1125899906826241):
        for logn in range(10, 15):
            for loge in (3, 4, 5):
                pl = plan(logn, loge)
                res[(q, logn, loge)] = (forward(pl, Fr(q)), inverse(list(reversed(pl)), Fr(q)))
        assert mac_budget(Fr(q)) < 8.0
    # base conversion / pointwise product: eight worst-case input limbs against worst-case and small output moduli
    res[(0, 0, 0)] = (canonical_product_budget([QMAX] * 8, [QMAX, (1 << 49) + 1, 3, 65537]), 0.0)
    return res


if __name__ == "__main__":
    for k, v in sorted(check_all().items()):
        if k == (0, 0, 0):
            print("canonical products (base conversion, reduce_mul): max |sum| / modulus %.3f" % v[0])
        else:
            print(k, "max |value|/q  fwd %.3f  inv %.3f" % v)
    print("all exactness preconditions hold for q <=", QMAX)
