"""CPU tests: the C oracle against (a) the reference's own deterministic test inputs, (b) the property and
cross-implementation checks the reference's tests make, (c) an independent big-int model (oracle/pymodel.py),
(d) the committed golden fixtures.  No GPU needed."""
import json
import os

import numpy as np
import pytest

from conftest import Q27, Q28, Q29, Q30, Q49, Q50, Q50B, Q60
from oracle import oracle as O
from oracle import pymodel as M

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- SURVEY Appendix B constants (computed there by a throw-away probe; independent check) ----------
@pytest.mark.parametrize("q,log_n,root", [(Q27, 10, 73993), (Q27, 11, 160172), (Q29, 11, 145054), (Q49, 11, 63091942852),
                                          (Q60, 11, 459811883340678), (Q28, 12, 62736), (Q30, 12, 236231),
                                          (Q50, 12, 46909545429), (Q50B, 12, 12064401162), (137438822401, 12, 8625844),
                                          (137438814209, 12, 52201411), (137438773249, 12, 22196635),
                                          (134215681, 10, 282116), (134176769, 10, 130311)])
def test_minimal_roots(q, log_n, root):
    assert O.min_primitive_root(log_n + 1, q) == root
    assert M.min_primitive_root(log_n + 1, q) == root
    if q < (1 << 30):
        assert O.min_primitive_root(log_n + 1, q, 32) == root


def test_constructor_errors_mirror_ntt_error():
    with pytest.raises(O.OracleError) as e:
        O.U64NttTable(12, 1125899906842597)       # root.rs:72-81
    assert e.value.code == 1
    with pytest.raises(O.OracleError) as e:
        O.U32NttTable(10, 2013265921)             # prime32/table.rs:195
    assert e.value.code == 5
    with pytest.raises(O.OracleError) as e:
        O.U64NttTable(10, 4611686018427394049)    # q >= 2^62, prime64/table.rs:318 (q = 2^62 + 6145 is 1 mod 2048)
    assert e.value.code in (1, 5)


# ---- NTT: the reference's own checks (prime64/tests.rs:14-272, prime32/tests.rs:14-236, tests/ntt.rs) --
@pytest.mark.parametrize("cls,q,log_n", [(O.U64NttTable, Q27, ln) for ln in range(3, 11)] +
                         [(O.U32NttTable, Q27, ln) for ln in range(3, 11)] +
                         [(O.U64NttTable, Q29, 11), (O.U64NttTable, Q49, 11), (O.U64NttTable, Q60, 11), (O.U64NttTable, Q50, 12)])
def test_ntt_cross_impl_roundtrip_ranges(cls, q, log_n):
    t = cls(log_n, q); n = 1 << log_n
    dt = np.uint64 if t.bits == 64 else np.uint32
    rng = np.random.default_rng(log_n)
    x = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    y = x.copy(); t.transform_slice(y)
    g = x.copy(); t.generic_transform_slice(g)
    assert np.array_equal(y, g)                                   # U64NttTable == UintNttTable (tests.rs:190-237)
    z = y.copy(); t.inverse_transform_slice(z); assert np.array_equal(z, x)   # round trip
    z = y.copy(); t.generic_inverse_transform_slice(z); assert np.array_equal(z, x)
    lz = x.copy(); t.lazy_transform_slice(lz)
    assert (lz < 4 * q).all() and np.array_equal(lz.astype(object) % q, y.astype(object))   # tests.rs:78-106
    li = y.copy(); t.lazy_inverse_transform_slice(li)
    assert (li < 2 * q).all() and np.array_equal(li.astype(object) % q, x.astype(object))   # tests.rs:113-142


@pytest.mark.parametrize("cls,q,log_n", [(O.U64NttTable, Q50, 5), (O.U32NttTable, Q27, 6), (O.U64NttTable, Q60, 4)])
def test_ntt_is_direct_evaluation_and_negacyclic(cls, q, log_n):
    t = cls(log_n, q); n = 1 << log_n
    dt = np.uint64 if t.bits == 64 else np.uint32
    rng = np.random.default_rng(2)
    x = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    y = x.copy(); t.transform_slice(y)
    assert [int(v) for v in y] == M.ntt_forward(x, q, t.root())
    assert np.array_equal(y, t.direct_transform(x))
    a = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    c = t.polymul_batch(a, x)
    assert [int(v) for v in c] == M.negacyclic_mul(a, x, q)
    assert np.array_equal(c, O.naive_mul(a, x, q, t.bits))
    for deg in (0, 1, n - 1):
        for coeff in (1, q - 1, 5):
            m = np.zeros(n, dtype=dt); m[deg] = coeff
            mm = m.copy(); t.transform_slice(mm)
            assert np.array_equal(mm, t.transform_monomial(coeff, deg))       # tests.rs monomial checks
    for r in (0, 3, n, n + 2, 2 * n - 1):
        assert [int(v) for v in O.mul_monomial(x, r, q, t.bits)] == M.mul_monomial(x, r, q)


# ---- Barrett / Shoup (barrett_modulus.rs:25-303, shoup_factor.rs:19-165) ---------------------------------
@pytest.mark.parametrize("bits,q", [(64, Q50), (64, Q60), (64, (1 << 62) - 57), (32, Q27), (32, (1 << 30) - 35), (64, 3), (32, 3)])
def test_barrett_and_shoup_exact(bits, q):
    m = O.BarrettModulus(q, bits)
    B = 1 << bits
    assert m.ratio[0] + (m.ratio[1] << bits) == (B * B) // q
    rng = np.random.default_rng(q % 1000)
    for _ in range(300):
        a, b, c = (int(rng.integers(0, q)) for _ in range(3))
        assert m.reduce_mul(a, b) == a * b % q
        assert m.reduce_mul_add(a, b, c) == (a * b + c) % q
        sf = O.ShoupFactor(a, q, bits)
        assert sf.quotient == (a << bits) // q
        y = int(rng.integers(0, B - 1, dtype=np.uint64, endpoint=True))
        lz = sf.lazy_factor_mul_modulo(y)
        assert lz < 2 * q and lz % q == a * y % q                      # shoup_factor.rs lazy < 2q
        assert sf.factor_mul_modulo(b) == m.reduce_mul(a, b)           # Shoup == Barrett
    for n in list(range(0, 66)) + [1000]:                              # odd lengths, SIMD-tail style (barrett_modulus.rs)
        dt = np.uint64 if bits == 64 else np.uint32
        a = rng.integers(0, q, n, dtype=np.uint64).astype(dt); b = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
        want = [(int(x) * int(y)) % q for x, y in zip(a, b)]
        assert [int(v) for v in m.reduce_mul_slice_to(a, b)] == want
        assert m.reduce_dot_product(a, b) == sum(want) % q
    with pytest.raises(ValueError):
        O.BarrettModulus(1 << (bits - 2), bits)                        # "modulus is too large" (barrett/mod.rs:40-42)


# ---- gadget decomposition (non_pow_of_2.rs:13-232, big_uint.rs:17-440) -----------------------------------
@pytest.mark.parametrize("bits,q,beta,rev", [(32, Q27, 7, None), (32, Q27, 7, 2), (64, Q50, 7, None), (32, 0b111000110, 3, 2),
                                             (64, Q29, 4, 7), (32, 1000003, 4, None), (64, Q60, 1, None), (32, 513, 4, None),
                                             (64, Q50, 13, 2), (32, Q27, 1, None)])
def test_decompose_properties_and_model(bits, q, beta, rev):
    b = O.ApproxSignedBasis(q, beta, rev, bits)
    g = M.Gadget(q, beta, rev)
    assert b.decompose_length() == g.levels and b.drop_bits() == g.drop and b.threshold() == g.threshold
    rng = np.random.default_rng(beta)
    dt = np.uint64 if bits == 64 else np.uint32
    v = rng.integers(0, q, 5000, dtype=np.uint64).astype(dt)
    v[:3] = [0, q - 1, q // 2]
    dig = b.decompose_slice(v)
    adj, car0 = b.init_value_carry_slice_to(v)
    bound = 0 if b.drop_bits() == 0 else 1 << (b.drop_bits() - 1)
    half = (1 << beta) // 2
    for i in range(len(v)):
        ds = [int(dig[l, i]) for l in range(g.levels)]
        signed = [d if d <= half else d - q for d in ds]
        assert all(-half <= s <= half for s in signed)                         # centred digits
        assert signed == [s if True else 0 for s in g.signed_digits(int(v[i]))] or beta == 1
        result = sum(s * d for s, d in zip(b.scalars(), ds)) % q
        diff = min((result - int(v[i])) % q, (int(v[i]) - result) % q)
        assert diff <= bound                                                     # non_pow_of_2.rs:140-141
        drop_m = 1 << b.drop_bits(); low = int(adj[i]) & (drop_m - 1)
        if car0[i]:
            assert int(v[i]) == (result - (drop_m - low)) % q                    # non_pow_of_2.rs:179-186
        else:
            assert int(v[i]) == (result + low) % q
    # slice == per level scalar path
    car = car0.copy()
    for l in range(g.levels):
        assert np.array_equal(b.decompose_level_slice_to(l, adj, car), dig[l])


# ---- RNS: the reference's deterministic inputs (primus_rns/tests/rns.rs:65-343) ------------------------------
def test_rns_reference_deterministic_cases():
    with pytest.raises(O.OracleError) as e:
        O.RNSBase([])
    assert e.value.code == 6                                              # rns.rs:67-70 EmptyBase
    with pytest.raises(O.OracleError) as e:
        O.RNSBase([21, 35])
    assert e.value.code == 7                                              # rns.rs:74-77 CoPrimeError
    base = O.RNSBase([3, 5, 7])                                           # rns.rs:81-100
    big = base.compose_multiple_values_to(np.array([2, 3, 2], dtype=np.uint64), 1)
    assert int(big[0]) == 23 and base.big_uint_value_len() == 1
    assert list(base.decompose_big_uint_values_to(big, 1)) == [2, 3, 2]
    moduli = [Q50, Q50B]                                                  # rns.rs:107-147
    base = O.RNSBase(moduli)
    by_value = [[0, 0], [1, 2], [97, 131], [Q50 - 1, Q50B - 2], [123_456_789, 987_654_321]]
    packed = np.array([[r[i] for r in by_value] for i in range(2)], dtype=np.uint64).reshape(-1)   # modulus-major
    big = base.compose_multiple_values_to(packed, len(by_value))
    vl = base.big_uint_value_len()
    assert vl == 2 and base.moduli_product() == Q50 * Q50B
    for vi, res in enumerate(by_value):
        val = sum(int(big[vi * vl + k]) << (64 * k) for k in range(vl))
        assert val == M.crt_compose(res, moduli) and val % Q50 == res[0] and val % Q50B == res[1]
    assert np.array_equal(base.decompose_big_uint_values_to(big, len(by_value)), packed)
    # centred lift rule (rns.rs expected_wrapping)
    for small_modulus in (2, 5, 128):
        small = np.arange(small_modulus, dtype=np.uint64)
        got = base.wrapping_decompose_small_values_to(small, small_modulus).reshape(2, -1)
        for li, m in enumerate(moduli):
            for v in range(small_modulus):
                exp = v if (small_modulus == 2 or v < (small_modulus + 1) // 2) else m - small_modulus + v
                assert int(got[li, v]) == exp
    # fused centred lift * scale + accumulate vs formula (rns.rs fused test)
    acc = np.array([5, 6, 7, 8, 9, 10], dtype=np.uint64)
    small = np.array([0, 3, 4], dtype=np.uint64)
    scal = [12345, 67890]
    out = base.wrapping_decompose_small_values_scaled_add_to(small, acc.copy(), 5, scal).reshape(2, -1)
    for li, m in enumerate(moduli):
        for vi, v in enumerate([0, 3, 4]):
            c = v if v < 3 else m - 5 + v
            assert int(out[li, vi]) == (int(acc[li * 3 + vi]) + scal[li] * c) % m


def test_bigbasis_matches_bigint_model():
    # the case of primus_decompose/tests/big_uint.rs:21-28: two 27-bit primes, log_basis 7 (u32 words)
    for bits, moduli, beta, rev in [(32, [134215681, 134176769], 7, None), (64, [Q50, Q50B], 7, None), (64, [Q50, Q50B, Q49], 7, 5),
                                    (64, [Q50], 7, None), (32, [Q27], 7, None), (64, [Q50, Q50B], 16, None)]:
        rns = O.RNSBase(moduli, bits)
        bb = O.BigUintApproxSignedBasis(rns, beta, rev)
        Q = rns.moduli_product()
        g = M.Gadget(Q, beta, rev)
        assert bb.decompose_length() == g.levels and bb.drop_bits() == g.drop
        rng = np.random.default_rng(9)
        n = 400
        dt = np.uint64 if bits == 64 else np.uint32
        res = np.stack([rng.integers(0, m, n, dtype=np.uint64).astype(dt) for m in moduli]).reshape(-1)
        big = rns.compose_multiple_values_to(res, n)
        vl = rns.big_uint_value_len()
        vals = [sum(int(big[i * vl + k]) << (bits * k) for k in range(vl)) for i in range(n)]
        for i in range(0, n, 37):
            assert vals[i] == M.crt_compose([int(res[li * n + i]) for li in range(len(moduli))], moduli)
        car = bb.init_value_carry_slice_inplace(big)
        digs = [bb.unsigned_decompose_slice_to(l, big, car) for l in range(g.levels)]
        for i in range(n):
            assert [int(d[i]) for d in digs] == g.unsigned_digits(vals[i])
        if len(moduli) == 1:   # L = 1: unsigned digit + centred lift == single-word signed digit (big_uint.rs:325)
            sb = O.ApproxSignedBasis(moduli[0], beta, rev, bits)
            sd = sb.decompose_slice(res)
            for l in range(g.levels):
                assert np.array_equal(rns.wrapping_decompose_small_values_to(digs[l], 1 << beta), sd[l])


# ---- external product: schoolbook identity --------------------------------------------------------
@pytest.mark.parametrize("bits,moduli,log_n,beta,rev,k", [(64, [Q50], 4, 7, None, 1), (32, [Q27], 5, 7, None, 1), (64, [Q50, Q50B], 4, 9, 4, 1),
                                                          (64, [Q50, Q50B, Q49], 3, 12, 3, 2), (32, [134215681, 134176769], 4, 7, None, 1)])
def test_external_product_schoolbook_identity(bits, moduli, log_n, beta, rev, k):
    n, L, dt = 1 << log_n, len(moduli), (np.uint64 if bits == 64 else np.uint32)
    dcrt = O.DcrtTable(log_n, moduli, bits); rns = O.RNSBase(moduli, bits); bb = O.BigUintApproxSignedBasis(rns, beta, rev)
    levels = bb.decompose_length(); Q = rns.moduli_product(); g = M.Gadget(Q, beta, rev)
    rng = np.random.default_rng(17)
    rnd = lambda: np.stack([rng.integers(0, m, n, dtype=np.uint64).astype(dt) for m in moduli])
    ggsw = np.stack([rnd() for _ in range((k + 1) * levels * (k + 1))]).reshape(k + 1, levels, k + 1, L, n)
    cin = np.stack([rnd() for _ in range(k + 1)])
    out = O.external_product(dcrt, rns, bb, k, ggsw, cin, to_coeff=True).reshape(k + 1, L, n)
    # expected: out_c = sum_{r,l} digit_{r,l} (*) INTT(key_{r,l,c})  per limb, digits as centred integers
    for li, m in enumerate(moduli):
        t = dcrt.tables[li]
        for c in range(k + 1):
            acc = [0] * n
            for r in range(k + 1):
                vals = [M.crt_compose([int(cin[r, lj, i]) for lj in range(L)], moduli) for i in range(n)]
                sd = [g.signed_digits(v) for v in vals]
                for l in range(levels):
                    keyc = ggsw[r, l, c, li].copy(); t.inverse_transform_slice(keyc)
                    dpoly = [sd[i][l] % m for i in range(n)]
                    prod = M.negacyclic_mul(dpoly, keyc, m)
                    acc = [(x + y) % m for x, y in zip(acc, prod)]
            assert [int(v) for v in out[c, li]] == acc
    if L == 1:   # single-word basis path is the same function
        sb = O.ApproxSignedBasis(moduli[0], beta, rev, bits)
        out1 = O.external_product_single(dcrt.tables[0], sb, k, ggsw.reshape(-1), cin.reshape(1, -1), to_coeff=True)
        assert np.array_equal(out1.reshape(-1), out.reshape(-1))


def test_blind_rotate_small_against_model():
    """With a 'trivial' key that makes the external product return D itself (key = gadget matrix), the
    accumulator becomes tv * X^(-b + sum a_i) up to the decomposition error; here only exact algebra is
    checked: zero mask leaves tv * X^(2N-b), and the composed loop equals its step-by-step definition."""
    q, log_n, beta = Q27, 4, 7
    t = O.U32NttTable(log_n, q); n = 1 << log_n
    sb = O.ApproxSignedBasis(q, beta, None, 32); lv = sb.decompose_length()
    rng = np.random.default_rng(1)
    n_lwe = 5
    bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    lwe = np.zeros((1, n_lwe + 1), dtype=np.uint32); lwe[0, -1] = 3
    acc = O.blind_rotate(t, sb, bsk, n_lwe, lwe, tv)[0]
    assert not acc[:n].any() and [int(v) for v in acc[n:]] == M.mul_monomial(tv, 2 * n - 3, q)   # a_i = 0: D = 0
    lwe = rng.integers(0, 2 * n, (1, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    got = O.blind_rotate(t, sb, bsk, n_lwe, lwe, tv)[0]
    acc = np.concatenate([np.zeros(n, dtype=np.uint32), np.array(M.mul_monomial(tv, (2 * n - int(lwe[0, -1])) % (2 * n), q), dtype=np.uint32)])
    bm = O.BarrettModulus(q, 32)
    for i in range(n_lwe):
        d = np.concatenate([bm.reduce_sub_slice_to(np.array(M.mul_monomial(acc[c * n:(c + 1) * n], int(lwe[0, i]), q), dtype=np.uint32),
                                                   acc[c * n:(c + 1) * n]) for c in range(2)])
        e = O.external_product_single(t, sb, 1, bsk[i * 2 * lv * 2 * n:(i + 1) * 2 * lv * 2 * n], d.reshape(1, -1), to_coeff=True)[0]
        acc = bm.reduce_add_slice_to(acc, e)
    assert np.array_equal(got, acc)


def test_golden_fixtures_match_oracle():
    """tests/golden/*.json were written by tests/golden/make_golden.py; the oracle must reproduce them."""
    path = os.path.join(GOLDEN, "golden_small.json")
    data = json.load(open(path))
    for case in data["ntt"]:
        cls = O.U64NttTable if case["bits"] == 64 else O.U32NttTable
        dt = np.uint64 if case["bits"] == 64 else np.uint32
        t = cls(case["log_n"], case["q"])
        assert t.root() == case["root"]
        x = np.array(case["input"], dtype=dt)
        y = x.copy(); t.transform_slice(y)
        assert [int(v) for v in y] == case["forward"]
    for case in data["decompose"]:
        dt = np.uint64 if case["bits"] == 64 else np.uint32
        b = O.ApproxSignedBasis(case["q"], case["log_basis"], case["levels"], case["bits"])
        d = b.decompose_slice(np.array(case["values"], dtype=dt))
        assert [[int(v) for v in row] for row in d] == case["digits"]
    for case in data["external_product"]:
        cls = O.U64NttTable if case["bits"] == 64 else O.U32NttTable
        dt = np.uint64 if case["bits"] == 64 else np.uint32
        t = cls(case["log_n"], case["q"]); sb = O.ApproxSignedBasis(case["q"], case["log_basis"], None, case["bits"])
        out = O.external_product_single(t, sb, 1, np.array(case["key"], dtype=dt), np.array(case["input"], dtype=dt).reshape(1, -1))
        assert [int(v) for v in out.reshape(-1)] == case["output"]


# ---- exactness budget of the lazy FP64 butterflies (product code: primus_fhe_b200/csrc/ntt_core.cuh, F64LazyField) ----------
def test_f64_lazy_fold_exactness_budget():
    """Walks every stage schedule the kernels use with exact rationals: quotient range of the magic-constant rounding and
    integer exactness (< 2^53) of every sum, for the largest admitted modulus (q <= 2^50 - 2^10)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("f64_bounds", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "f64_bounds.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    res = mod.check_all()
    assert len(res) == 46 and all(f < 8.001 and i < 8.001 for f, i in res.values())   # the hard checks (<= 2^53) are the script's asserts
    assert (0, 0, 0) in res                                                           # base conversion / reduce_mul products (round 2)
    assert mod.QMAX == (1 << 50) - 1024


# ---- BaseConverter: the reference's deterministic case (primus_rns/tests/rns.rs:281-345) + big-int model ---------------
def test_base_converter_reference_case_and_model():
    in_m, out_m = [17, 19, 23], [29, 31]
    by_value = [[0, 0, 0], [1, 2, 3], [16, 18, 22], [7, 11, 13], [4, 0, 19]]
    n = len(by_value)
    crt_in = np.array([[r[i] for r in by_value] for i in range(3)], dtype=np.uint64).reshape(-1)
    conv = O.BaseConverter(in_m, out_m)
    Q = 17 * 19 * 23
    assert [[int(v) for v in row] for row in conv.matrix()] == [[(Q // q) % p for q in in_m] for p in out_m]   # converter.rs:57-66
    out = conv.fast_convert_array(crt_in, n).reshape(2, n)
    for vi, res in enumerate(by_value):      # scalar fast_convert formula (converter.rs:111-137)
        y = [res[i] * pow(Q // in_m[i], -1, in_m[i]) % in_m[i] for i in range(3)]
        for k, p in enumerate(out_m):
            assert int(out[k, vi]) == sum(y[i] * ((Q // in_m[i]) % p) for i in range(3)) % p
    # exact conversion on small canonical values == trivial modulo reduction (rns.rs:327-344)
    ex = O.BaseConverter(in_m, [37])
    vals = [0, 1, 2, 7, 16]
    exact_in = np.array([[v] * len(vals) for v in [0]], dtype=np.uint64)  # placeholder, rebuilt below
    exact_in = np.array([[v for v in vals] for _ in range(3)], dtype=np.uint64).reshape(-1)
    assert [int(v) for v in ex.exact_convert_array(exact_in, len(vals))] == [v % 37 for v in vals]
    with pytest.raises(ValueError):
        conv.exact_convert_array(crt_in, n)
    # larger bases: fast conversion equals x + alpha*Q (0 <= alpha < L) mod p; exact conversion equals the centred lift mod p
    rng = np.random.default_rng(77)
    for bits, im, om in [(64, [Q50, Q50B, Q49], [Q60, 1125899904679937]), (32, [134215681, 134176769], [Q27, Q28])]:
        dt = np.uint64 if bits == 64 else np.uint32
        c = O.BaseConverter(im, om, bits)
        Qb = int(np.prod([int(m) for m in im], dtype=object))
        n = 300
        xs = [int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) % Qb for _ in range(n)]
        xs[0], xs[1], xs[2] = 0, 1, Qb - 1
        cin = np.array([[x % m for x in xs] for m in im], dtype=np.uint64).astype(dt).reshape(-1)
        out = c.fast_convert_array(cin, n).reshape(len(om), n)
        for j, x in enumerate(xs):
            y = [(x % m) * pow(Qb // m, -1, m) % m for m in im]
            full = sum(yi * (Qb // m) for yi, m in zip(y, im))
            assert full % Qb == x and 0 <= (full - x) // Qb < len(im)
            for k, p in enumerate(om):
                assert int(out[k, j]) == full % p
        c1 = O.BaseConverter(im, om[:1], bits)
        ex = c1.exact_convert_array(cin, n)
        for j, x in enumerate(xs):
            if Qb // 8 < x < 3 * Qb // 8 or 5 * Qb // 8 < x < 7 * Qb // 8:   # away from the rounding boundary Q/2 (float correction term)
                centred = x if x < Qb // 2 else x - Qb
                assert int(ex[j]) == centred % om[0]


def test_blind_rotation_schedule_model_matches_oracle():
    """tools/emulate_br32.py re-executes the re-scheduled blind-rotation kernel of lattice32.cu at thread / register / shared-memory
    level in numpy (index maps, twiddle indices, padded exchange addresses, carry-free digits, Montgomery reduction, 2^32
    compensation) and compares with the oracle: the design check that runs without a GPU."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "emulate_br32.py"), "2", "3"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("== oracle") >= 6
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "emulate_ep32_2048.py"), "5"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr        # N = 2048 schedule of lattice32_ep.cu (3 + 3 + 3 + 2 passes)
    assert p.stdout.count("== oracle") >= 3


def test_ternary_blind_rotation_oracle_reduces_to_rotation():
    """Sanity of the oracle's ternary composition: with BSK+ = RGSW(1) (trivial, noiseless: the gadget rows) and BSK- = RGSW(0) = 0 the
    step multiplies the accumulator by X^a up to the decomposition's rounding error, and with both keys zero it is the identity."""
    import numpy as np
    from oracle import oracle as O
    q, n = 132120577, 1024
    ot = O.U32NttTable(10, q); ob = O.ApproxSignedBasis(q, 7, None, 32); lv = ob.decompose_length()
    rng = np.random.default_rng(9)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    zero = np.zeros(2 * 2 * lv * 2 * n, dtype=np.uint32)
    lwe = np.array([[5, 7, 3]], dtype=np.uint32)
    acc = O.blind_rotate_ternary(ot, ob, zero, zero, 2, lwe, tv, batch=1)[0]
    assert np.array_equal(acc[:n], np.zeros(n, np.uint32)) and np.array_equal(acc[n:], O.mul_monomial(tv, 2 * n - 3, q, 32))
    # trivial RGSW(1): row r, level l, component c = NTT(g_l) if c == r else 0, g_l = 2^(drop + l*beta)
    plus = np.zeros((2, 2, lv, 2, n), dtype=np.uint32)
    for i in range(2):
        for r in range(2):
            for l, g in enumerate(ob.scalars()):
                plus[i, r, l, r, :] = g % q      # NTT of the constant polynomial g is the constant vector g
    acc = O.blind_rotate_ternary(ot, ob, plus.reshape(-1), zero, 2, lwe, tv, batch=1)[0]
    want = O.mul_monomial(tv, (2 * n - 3 + 5 + 7) % (2 * n), q, 32).astype(np.int64)
    err = (acc[n:].astype(np.int64) - want + q // 2) % q - q // 2
    assert np.abs(err).max() <= 2 * (1 << ob.drop_bits())      # two steps, each within the gadget's rounding error


@pytest.mark.parametrize("bits,q,log_n", [(32, Q27, 10), (32, Q27, 11), (64, Q50, 11)])
def test_external_product_schoolbook_identity_at_baseline_shapes(bits, q, log_n):
    """The same identity as above at the BASELINE degrees (C4: N = 2048, C5: N = 1024):
    out_c = sum_{r,l} digit_{r,l} (*) INTT(key_{r,l,c}) in Z_q[X]/(X^N + 1), with the digits taken from the big-int model and the negacyclic
    products computed exactly by integer convolution (np.convolve on int64 with the key split into 25-bit halves, no NTT involved)."""
    n, beta, dt = 1 << log_n, 7, (np.uint64 if bits == 64 else np.uint32)
    t = (O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q)
    sb = O.ApproxSignedBasis(q, beta, None, bits); levels = sb.decompose_length()
    g = M.Gadget(q, beta, None)
    rng = np.random.default_rng(19)
    key = rng.integers(0, q, (2, levels, 2, n), dtype=np.uint64).astype(dt)
    cin = rng.integers(0, q, (2, n), dtype=np.uint64).astype(dt)
    cin[0, :4] = (0, q - 1, q // 2, q // 2 + 1)
    out = O.external_product_single(t, sb, 1, key.reshape(-1), cin.reshape(1, -1), to_coeff=True).reshape(2, n)

    def negacyclic(d, kc):   # d: small signed int64 digits, kc: canonical words (python ints allowed) -> exact product mod q
        lo = np.array([int(v) & ((1 << 25) - 1) for v in kc], dtype=np.int64)
        hi = np.array([int(v) >> 25 for v in kc], dtype=np.int64)
        full = [int(a) + (int(b) << 25) for a, b in zip(np.convolve(d, lo), np.convolve(d, hi))]   # |terms| < 64 * 2^25 * 2048 = 2^42
        return [(full[i] - (full[i + n] if i + n < len(full) else 0)) % q for i in range(n)]

    digits = np.array([[g.signed_digits(int(v)) for v in cin[r]] for r in range(2)], dtype=np.int64)   # [r][i][l]
    for c in range(2):
        acc = [0] * n
        for r in range(2):
            for l in range(levels):
                kc = key[r, l, c].copy(); t.inverse_transform_slice(kc)
                prod = negacyclic(digits[r, :, l], kc)
                acc = [(x + y) % q for x, y in zip(acc, prod)]
        assert [int(v) for v in out[c]] == acc


def test_blind_rotation_oracle_rotates_at_the_c5_degree():
    """Semantics of the composed (binary-secret) blind rotation at N = 1024: with noiseless trivial keys BSK_i = RGSW(s_i) (gadget rows for
    s_i = 1, zero for s_i = 0) the accumulator ends as tv * X^(-b + sum a_i s_i) up to the gadget's rounding error per step -- i.e. the
    CMux form defined in SURVEY App. A.6 / include/pfhe.h really evaluates the LWE phase in the exponent."""
    q, n, n_lwe = Q27, 1024, 48
    ot = O.U32NttTable(10, q); ob = O.ApproxSignedBasis(q, 7, None, 32); lv = ob.decompose_length()
    rng = np.random.default_rng(23)
    s = rng.integers(0, 2, n_lwe)
    bsk = np.zeros((n_lwe, 2, lv, 2, n), dtype=np.uint32)
    for i in range(n_lwe):
        if s[i]:
            for r in range(2):
                for l, gl in enumerate(ob.scalars()):
                    bsk[i, r, l, r, :] = gl % q          # NTT of the constant polynomial g_l
    lwe = rng.integers(0, 2 * n, (1, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    acc = O.blind_rotate(ot, ob, bsk.reshape(-1), n_lwe, lwe, tv, batch=1)[0]
    shift = (-int(lwe[0, n_lwe]) + int(sum(int(a) * int(b) for a, b in zip(lwe[0, :n_lwe], s)))) % (2 * n)
    want = O.mul_monomial(tv, shift, q, 32).astype(np.int64)
    err = (acc[n:].astype(np.int64) - want + q // 2) % q - q // 2
    assert np.abs(err).max() <= n_lwe * (1 << ob.drop_bits())
    erra = (acc[:n].astype(np.int64) + q // 2) % q - q // 2
    assert np.abs(erra).max() <= n_lwe * (1 << ob.drop_bits())


def test_ifma_restatement_matches_scalar_oracle():
    """oracle/pfhe_oracle_avx512.c (the CPU baseline of bench.py on AVX-512 IFMA hosts) is bit-identical to the scalar Harvey transform for every
    degree it accepts, including the q-1 / 0 edge rows; skipped where the CPU lacks IFMA (the baseline then falls back to the scalar port)."""
    t0 = O.U64NttTable(4, Q50)
    if not t0.simd_supported():
        pytest.skip("no AVX-512 IFMA on this host")
    for q in (Q50, Q49, 1073692673, 132120577):
        for log_n in (4, 5, 6, 9, 11, 12, 13):
            if (q - 1) % (2 << log_n):
                continue
            t = O.U64NttTable(log_n, q)
            rng = np.random.default_rng(log_n)
            x = rng.integers(0, q, (5, 1 << log_n), dtype=np.uint64)
            x[0, :] = q - 1
            x[1, :] = 0
            a = x.copy(); t.forward_batch(a, 1)
            b = x.copy(); t.forward_batch_simd(b, 2)
            assert np.array_equal(a, b), (q, log_n)
    assert not O.U64NttTable(10, Q60).simd_supported()      # 60-bit primes are outside the BIT_SHIFT = 52 back-end


def test_oracle_programmable_bootstrap_recovers_the_lookup_table_at_c5_parameters():
    """Functional pin of the composed bootstrap (modulus switch -> blind rotation -> extract_lwe) at the BASELINE config-5 parameters: real
    LWE / RGSW encryptions WITH noise, decrypted afterwards -- every message m must come back as LUT[m] with noise far below the decoding
    margin.  The reference has no bootstrapping to compare with; this is the property its users rely on.  The GPU runs the same
    flow in tests/test_gpu_bootstrap_functional.py."""
    import bootstrap_common as B
    t = O.U32NttTable(B.LOG_N, B.Q)
    basis = O.ApproxSignedBasis(B.Q, B.LOG_B, None, 32)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    rng = np.random.default_rng(20261018)
    z, s = B.secrets(rng)
    a1, e1 = B.rgsw_rows(rng, lv)
    zrep = lambda rows: np.ascontiguousarray(np.broadcast_to(z.astype(np.uint32), rows.shape))
    az = t.polymul_batch(a1.copy(), zrep(a1)).reshape(a1.shape)
    key = np.ascontiguousarray(B.assemble_key(a1, az, e1, s, lv, drop).reshape(-1, B.N))
    t.forward_batch(key)
    batch = 6
    msgs, lwe_q = B.lwe_inputs(rng, s, batch)
    lwe_2n = O.modulus_switch(lwe_q, B.Q, B.LOG_N + 1)
    ph2n = B.check_switched(lwe_2n, s, msgs)
    acc = O.blind_rotate(t, basis, key.reshape(-1), B.N_LWE, lwe_2n, B.test_vector(), batch=batch).reshape(batch, 2, B.N)
    out = np.stack([O.extract_lwe(acc[i].reshape(-1), B.Q, 32) for i in range(batch)])
    a0 = np.ascontiguousarray(acc[:, 0])
    worst = B.check_outputs(out, acc, t.polymul_batch(a0.copy(), zrep(a0)).reshape(a0.shape), z, msgs, ph2n)
    assert worst > 0          # the keys really carried noise


def test_oracle_ternary_bootstrap_recovers_the_lookup_table_at_c5_parameters():
    """The same functional pin for the ternary-secret rotation by monomial combination: BSK+_i = RGSW([s_i = 1]),
    BSK-_i = RGSW([s_i = -1]) with noise; the bootstrap must return LUT[m]."""
    import bootstrap_common as B
    t = O.U32NttTable(B.LOG_N, B.Q)
    basis = O.ApproxSignedBasis(B.Q, B.LOG_B, None, 32)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    rng = np.random.default_rng(77)
    z, s = B.secrets(rng, ternary=True)
    zrep = lambda rows: np.ascontiguousarray(np.broadcast_to(z.astype(np.uint32), rows.shape))
    keys = []
    for sign in (1, -1):
        a1, e1 = B.rgsw_rows(rng, lv)
        az = t.polymul_batch(a1.copy(), zrep(a1)).reshape(a1.shape)
        k = np.ascontiguousarray(B.assemble_key(a1, az, e1, (s == sign).astype(np.int64), lv, drop).reshape(-1, B.N))
        t.forward_batch(k)
        keys.append(k.reshape(-1))
    batch = 2
    msgs, lwe_q = B.lwe_inputs(rng, s, batch)
    lwe_2n = O.modulus_switch(lwe_q, B.Q, B.LOG_N + 1)
    ph2n = B.check_switched(lwe_2n, s, msgs)
    acc = O.blind_rotate_ternary(t, basis, keys[0], keys[1], B.N_LWE, lwe_2n, B.test_vector(), batch=batch).reshape(batch, 2, B.N)
    out = np.stack([O.extract_lwe(acc[i].reshape(-1), B.Q, 32) for i in range(batch)])
    a0 = np.ascontiguousarray(acc[:, 0])
    B.check_outputs(out, acc, t.polymul_batch(a0.copy(), zrep(a0)).reshape(a0.shape), z, msgs, ph2n)


@pytest.mark.parametrize("bits,moduli,log_n", [(32, [Q27], 11), (64, [Q50], 11), (64, [Q50, Q50B], 11)])
def test_oracle_external_product_decrypts_to_the_product_at_c4_shapes(bits, moduli, log_n):
    """Functional pin of the GGSW external product (a19 / a20) at the BASELINE config-4 degree: RLWE(m) [x] RGSW(mu) with real noisy
    encryptions must decrypt to m * mu (mu = 1 + X^5 - X^100) -- single-word u32 / u64 gadget and the two-limb multi-word gadget."""
    import extprod_common as X
    dt = np.uint64 if bits == 64 else np.uint32
    n, L = 1 << log_n, len(moduli)
    tables = [(O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q) for q in moduli]
    def ring_mul(i, rows, z):
        zz = np.ascontiguousarray(np.broadcast_to(z.astype(dt), rows.shape))
        return tables[i].polymul_batch(np.ascontiguousarray(rows).copy(), zz).reshape(rows.shape)
    rng = np.random.default_rng(5)
    if L == 1:
        basis = O.ApproxSignedBasis(moduli[0], 7, None, bits)
    else:
        rns = O.RNSBase(moduli, bits); basis = O.BigUintApproxSignedBasis(rns, 7, None)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    batch = 2 if L == 1 else 1
    key, glwe, z, msg = X.rgsw_and_inputs(rng, moduli, n, lv, drop, 7, batch, ring_mul, dt)
    for i in range(L):                                       # NTT form, limb by limb
        rows = np.ascontiguousarray(key[:, :, :, i, :].reshape(-1, n))
        tables[i].forward_batch(rows)
        key[:, :, :, i, :] = rows.reshape(2, lv, 2, n)
    if L == 1:
        out = O.external_product_single(tables[0], basis, 1, key.reshape(-1), glwe.reshape(-1), to_coeff=True, batch=batch)
    else:
        out = O.external_product(O.DcrtTable(log_n, moduli, bits), rns, basis, 1, key.reshape(-1), glwe.reshape(-1), to_coeff=True, batch=batch)
    X.check(np.asarray(out).reshape(batch, 2, L, n), moduli, z, msg, ring_mul)


def test_oracle_c3_rns_product_is_the_big_integer_product():
    """BASELINE config 3 (8 limbs of ~50 bits, N = 16384), independent of any NTT: the per-limb products, CRT-composed, equal the schoolbook
    negacyclic product of the composed operands mod Q at sampled coefficients (exact big-integer arithmetic)."""
    from bench import _c3_primes
    mods = _c3_primes(); n = 16384
    Q = 1
    for m in mods:
        Q *= m
    rng = np.random.default_rng(81)
    a = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
    b = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
    c = np.stack([O.U64NttTable(14, m).polymul_batch(a[i:i + 1].copy(), b[i:i + 1].copy(), 1).reshape(-1) for i, m in enumerate(mods)])
    orns = O.RNSBase(mods, 64); vl = orns.big_uint_value_len()

    def compose(res):
        w = orns.compose_multiple_values_to(np.ascontiguousarray(res).reshape(-1), n).reshape(n, vl)
        return [sum(int(w[i, k]) << (64 * k) for k in range(vl)) for i in range(n)]

    A, B_, C_ = compose(a), compose(b), compose(c)
    for j in (0, 1, 8191, n - 1):
        acc = 0
        for i in range(n):
            k = j - i
            acc += A[i] * B_[k] if k >= 0 else -A[i] * B_[k + n]
        assert C_[j] == acc % Q, j
