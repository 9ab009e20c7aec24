"""GPU parity for the pointwise slice operators, gadget decomposition, RNS lift and LWE extraction.

Mirrors primus_modulus/tests/barrett_modulus.rs:25-303 (slice ops vs a second implementation, odd
lengths 0..65 for vector tails), primus_factor/tests/shoup_factor.rs:19-165 (Shoup == Barrett),
primus_decompose/tests/non_pow_of_2.rs:13-232 (slice == scalar, digits), primus_rns/tests/rns.rs (lift rule).
"""
import numpy as np
import pytest

from conftest import Q27, Q50, Q50B, Q60

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.from_numpy(x.view(np.int64 if x.dtype == np.uint64 else np.int32)).cuda()


def _host(t, dt):
    return t.cpu().numpy().view(dt)


@pytest.mark.parametrize("bits,q", [(64, Q50), (64, Q60), (32, Q27), (32, 1073692673), (64, 3), (32, 7)])
@pytest.mark.parametrize("n", [0, 1, 3, 17, 64, 65, 4096, 100003])
def test_slice_ops_match_oracle(bits, q, n):
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(n + bits)
    a, b, c = (rng.integers(0, q, n, dtype=np.uint64).astype(dt) for _ in range(3))
    if n > 2:
        a[0], b[0], c[0] = q - 1, q - 1, q - 1
        a[1], b[1] = 0, q - 1
    s = int(rng.integers(0, q))
    om, gm = O.BarrettModulus(q, bits), P.BarrettModulus(q, bits)
    of = O.ShoupFactor(s, q, bits)
    if n == 0:
        return
    da, db, dc = _dev(a), _dev(b), _dev(c)
    out = _dev(np.zeros(n, dtype=dt))
    assert np.array_equal(_host(gm.reduce_mul_slice_to(da, db, out), dt), om.reduce_mul_slice_to(a, b))
    assert np.array_equal(_host(gm.reduce_mul_add_slice_to(da, db, dc, out), dt), om.reduce_mul_add_slice_to(a, b, c))
    assert np.array_equal(_host(gm.reduce_add_slice_to(da, db, out), dt), om.reduce_add_slice_to(a, b))
    assert np.array_equal(_host(gm.reduce_sub_slice_to(da, db, out), dt), om.reduce_sub_slice_to(a, b))
    assert np.array_equal(_host(gm.reduce_neg_slice_to(da, out), dt), om.reduce_neg_slice_to(a))
    assert np.array_equal(_host(gm.reduce_mul_scalar_slice_to(da, s, out), dt), om.reduce_mul_scalar_slice_to(a, s))
    acc = _dev(c.copy()); gm.reduce_add_mul_slice_assign(acc, da, db)
    assert np.array_equal(_host(acc, dt), om.reduce_add_mul_slice_assign(c.copy(), a, b))
    acc = _dev(c.copy()); gm.reduce_sub_mul_slice_assign(acc, da, db)
    assert np.array_equal(_host(acc, dt), om.reduce_sub_mul_slice_assign(c.copy(), a, b))
    acc = _dev(c.copy()); gm.reduce_add_mul_scalar_slice_assign(acc, da, s)
    assert np.array_equal(_host(acc, dt), om.reduce_add_mul_scalar_slice_assign(c.copy(), a, s))
    # Shoup factor ops == Barrett (shoup_factor.rs:19-60)
    assert np.array_equal(_host(gm.factor_mul_slice_to(s, da, out), dt), of.factor_mul_slice_to(a))
    assert np.array_equal(of.factor_mul_slice_to(a), om.reduce_mul_scalar_slice_to(a, s))
    acc = _dev(c.copy()); gm.add_factor_mul_slice_assign(s, acc, da)
    assert np.array_equal(_host(acc, dt), of.add_factor_mul_slice_assign(c.copy(), a))
    acc = _dev(c.copy()); gm.sub_factor_mul_slice_assign(s, acc, da)
    assert np.array_equal(_host(acc, dt), of.sub_factor_mul_slice_assign(c.copy(), a))
    # a word that is not a canonical residue in an otherwise canonical vector: the double-word Barrett reduction still applies
    # (barrett/mod.rs:99-139); the FP64 product of the u64 q < 2^50 path must hand that vector to the integer product
    if n > 8 and q < (1 << (bits - 2)):
        a2 = a.copy(); a2[5] = dt(q + 5); a2[6] = dt(2 * q - 1)
        assert np.array_equal(_host(gm.reduce_mul_slice_to(_dev(a2), db, out), dt), om.reduce_mul_slice_to(a2, b))
    # in-place alias (mul_assign)
    ia = _dev(a.copy()); gm.reduce_mul_slice_assign(ia, db)
    assert np.array_equal(_host(ia, dt), om.reduce_mul_slice_to(a, b))
    # host-slice shims
    ho = np.zeros(n, dtype=dt)
    assert np.array_equal(gm.reduce_mul_slice_to(a, b, ho), om.reduce_mul_slice_to(a, b))
    hacc = c.copy(); gm.reduce_add_mul_slice_assign(hacc, a, b)
    assert np.array_equal(hacc, om.reduce_add_mul_slice_assign(c.copy(), a, b))


def test_dcrt_per_limb_slice_ops():
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli, n, rows = [Q50, Q50B, Q60], 1024, 3
    rng = np.random.default_rng(3)
    a = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in moduli]) for _ in range(rows)])
    b = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in moduli]) for _ in range(rows)])
    acc = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in moduli]) for _ in range(rows)])
    gm = P.BarrettModulus(moduli, 64, n=n)
    want = acc.copy()
    for li, m in enumerate(moduli):
        for r in range(rows):
            want[r, li] = O.BarrettModulus(m).reduce_add_mul_slice_assign(acc[r, li].copy(), a[r, li], b[r, li])
    d = _dev(acc.copy())
    gm.reduce_add_mul_slice_assign(d, _dev(a), _dev(b))
    assert np.array_equal(_host(d, np.uint64), want)
    # Shoup scalar MAC with one factor per limb (DcrtPolynomial::add_mul_factor_assign, dcrt/mul.rs:142-161)
    factors = [int(rng.integers(0, m)) for m in moduli]
    want = acc.copy()
    for li, m in enumerate(moduli):
        for r in range(rows):
            want[r, li] = O.ShoupFactor(factors[li], m).add_factor_mul_slice_assign(acc[r, li].copy(), a[r, li])
    d = _dev(acc.copy())
    gm.add_factor_mul_slice_assign(factors, d, _dev(a))
    assert np.array_equal(_host(d, np.uint64), want)


@pytest.mark.parametrize("bits,q,log_basis,rev", [(32, Q27, 7, None), (32, Q27, 7, 2), (64, Q50, 7, None), (64, Q50, 7, 3),
                                                  (32, 0b111000110, 3, 2), (64, 536813569, 4, 7), (64, Q60, 1, None),
                                                  (64, Q60, 16, None), (32, Q27, 1, 5), (64, Q50, 25, None)])
def test_decompose_matches_oracle(bits, q, log_basis, rev):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    ob = O.ApproxSignedBasis(q, log_basis, rev, bits)
    gb = P.ApproxSignedBasis(q, log_basis, rev, bits)
    assert gb.decompose_length() == ob.decompose_length() and gb.drop_bits() == ob.drop_bits()
    rng = np.random.default_rng(11)
    for n in (10007, 10008):   # scalar kernel (ragged length) / 16-byte vector kernel
        v = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
        v[:4] = [0, q - 1, q // 2, min(q - 1, (ob.threshold() or 1))]
        want = ob.decompose_slice(v)
        dig = torch.empty((gb.decompose_length(), n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
        gb.decompose_batch(_dev(v), dig)
        assert np.array_equal(_host(dig, dt), want)


def test_rns_lift_and_extract_lwe():
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = [Q50, Q50B, Q60]
    rng = np.random.default_rng(4)
    for B in (2, 128, 5):
        small = rng.integers(0, B, 1000, dtype=np.uint64)
        want = O.RNSBase(moduli).wrapping_decompose_small_values_to(small, B)
        out = torch.empty(3 * 1000, dtype=torch.int64, device="cuda")
        P.RNSBase(moduli).wrapping_decompose_small_values_to(_dev(small), out, B)
        assert np.array_equal(_host(out, np.uint64), want)
    n, batch = 1024, 5
    rl = rng.integers(0, Q27, (batch, 2 * n), dtype=np.uint64).astype(np.uint32)
    rl[0, 1] = 0
    out = torch.empty((batch, n + 1), dtype=torch.int32, device="cuda")
    P.extract_lwe_batch(Q27, _dev(rl), out, n, 32)
    got = _host(out, np.uint32)
    for i in range(batch):
        assert np.array_equal(got[i], O.extract_lwe(rl[i], Q27, 32))


@pytest.mark.parametrize("bits,moduli,n", [(64, [1125899906826241], 1024), (64, [1125899906826241, 1125899906629633, 562949953392641], 256),
                                           (32, [132120577, 134176769], 2048), (64, [1152921504606830593], 7)])
def test_butterfly_mul_factor_matches_oracle(bits, moduli, n):
    """(a, out) = (a + s, (a - s) * w): DcrtPolynomial::butterfly_mul_factor_to (primus_poly/src/dcrt/mul.rs:189-222)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(6)
    rows, L = 3, len(moduli)
    mk = lambda r: np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64).astype(dt) for m in moduli]) for _ in range(r)])
    a, s, w = mk(rows), mk(rows), mk(1)[0]
    a[0, :, 0] = 0; s[0, :, 0] = np.array([m - 1 for m in moduli], dtype=dt)
    want_a, want_out = O.butterfly_mul_factor(a, s, w, moduli, n)
    da, ds, dw = _dev(a.copy()), _dev(s), _dev(w)
    out = torch.empty_like(da)
    P.butterfly_mul_factor_batch(moduli, da, ds, dw, out, n, bits)
    assert np.array_equal(da.cpu().numpy().view(dt).reshape(a.shape), want_a)
    assert np.array_equal(out.cpu().numpy().view(dt).reshape(a.shape), want_out)


@pytest.mark.parametrize("bits,q", [(64, 1125899906826241), (64, 1152921504606830593), (32, 132120577), (32, 1073692673)])
def test_inv_slice_matches_bigint(bits, q):
    """NttPolynomial::inv_to / reduce_inv_slice_to (primus_poly/src/ntt/inv.rs:1-58): the modular inverse is unique, so the
    check is exact big-int arithmetic (pow(a, -1, q)) plus a * a^-1 == 1 through the GPU's own product."""
    import torch
    import primus_fhe_b200 as P
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(3)
    n = 500
    a = rng.integers(1, q, n, dtype=np.uint64).astype(dt)
    a[0], a[1] = 1, q - 1
    da = _dev(a); out = torch.empty_like(da)
    assert P.inv_slice_batch(q, da, out, bits) is None
    got = out.cpu().numpy().view(dt)
    assert [int(v) for v in got] == [pow(int(v), -1, q) for v in a]
    a[7] = 0; a[300] = 0                                   # try_reduce_inv_slice_to -> NoInverseAtIndex { index: 7 }
    assert P.inv_slice_batch(q, _dev(a), out, bits) == 7
    assert int(out.cpu().numpy().view(dt)[7]) == 0


@pytest.mark.parametrize("bits,q", [(64, 1125899906826241), (64, 1152921504606830593), (32, 132120577)])
def test_remaining_slice_ops_match_bigint(bits, q):
    """reduce_double_slice_to, reduce_sub_slice_rev_assign, reduce_mul_scalar_add_slice_to (primus_reduce/src/slice_ops.rs:91-133,
    :229) and FactorSliceOps::factor_mul_add_slice_to (primus_factor/src/ops.rs:117): exact big-int check, device and host shims."""
    import torch
    import primus_fhe_b200 as P
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(12)
    n = 1000
    a = rng.integers(0, q, n, dtype=np.uint64).astype(dt); b = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    c = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    a[0], b[0], c[0] = q - 1, q - 1, q - 1
    s = int(rng.integers(1, q))
    m = P.BarrettModulus(q, bits)
    A, B, Cc = [int(v) for v in a], [int(v) for v in b], [int(v) for v in c]
    for host in (False, True):
        conv = (lambda x: x.copy()) if host else (lambda x: _dev(x.copy()))
        back = (lambda t: t) if host else (lambda t: t.cpu().numpy().view(dt))
        out = conv(np.zeros(n, dtype=dt))
        m.reduce_double_slice_to(conv(a), out)
        assert [int(v) for v in back(out)] == [2 * x % q for x in A]
        bb = conv(b); m.reduce_sub_slice_rev_assign(conv(a), bb)
        assert [int(v) for v in back(bb)] == [(x - y) % q for x, y in zip(A, B)]
        m.reduce_mul_scalar_add_slice_to(conv(a), s, conv(c), out)
        assert [int(v) for v in back(out)] == [(x * s + z) % q for x, z in zip(A, Cc)]
        m.factor_mul_add_slice_to(s, conv(a), conv(c), out)
        assert [int(v) for v in back(out)] == [(x * s + z) % q for x, z in zip(A, Cc)]


@pytest.mark.parametrize("bits,q,n", [(64, 1125899906826241, 1024), (32, 132120577, 512)])
def test_extract_lwe_variants_match_oracle(bits, q, n):
    """extract_lwe_with_index / extract_first_few_lwe (primus_lattice/src/rlwe/coeff.rs:194-261)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(15)
    batch = 4
    rlwe = rng.integers(0, q, (batch, 2 * n), dtype=np.uint64).astype(dt)
    rlwe[0, :5] = 0
    d = _dev(rlwe)
    for index in (0, 1, 17, n - 1):
        out = torch.empty((batch, n + 1), dtype=d.dtype, device="cuda")
        P.extract_lwe_ex_batch(q, d, out, n, index=index, count=1, bits=bits)
        want = np.stack([O.extract_lwe_with_index(rlwe[b], index, q) for b in range(batch)])
        assert np.array_equal(out.cpu().numpy().view(dt), want)
    assert np.array_equal(np.stack([O.extract_lwe_with_index(rlwe[b], 0, q) for b in range(batch)]),
                          np.stack([O.extract_lwe(rlwe[b], q, bits) for b in range(batch)]))
    for count in (1, 3, n):
        out = torch.empty((batch, n + count), dtype=d.dtype, device="cuda")
        P.extract_lwe_ex_batch(q, d, out, n, index=0, count=count, bits=bits)
        want = np.stack([O.extract_first_few_lwe(rlwe[b], count, q) for b in range(batch)])
        assert np.array_equal(out.cpu().numpy().view(dt), want)


@pytest.mark.parametrize("bits,moduli,n,group", [(64, [1125899906826241], 1024, 2), (32, [132120577], 2048, 2),
                                                 (64, [1125899906826241, 1125899906629633], 256, 3)])
def test_ciphertext_times_polynomial_broadcast(bits, moduli, n, group):
    """NttRlwe::mul_ntt_polynomial_to / add_ntt_rlwe_mul_ntt_polynomial_assign (primus_lattice/src/rlwe/ntt.rs:78-152): every
    component of ciphertext i times polynomial i; checked against the oracle's per-polynomial Barrett slice ops."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(19)
    cts, L = 3, len(moduli)
    mk = lambda r: np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64).astype(dt) for m in moduli]) for _ in range(r)])
    a, b, acc = mk(cts * group), mk(cts), mk(cts * group)
    want_mul, want_acc = np.empty_like(a), acc.copy()
    for r in range(cts * group):
        for li, m in enumerate(moduli):
            om = O.BarrettModulus(m, bits)
            want_mul[r, li] = om.reduce_mul_slice_to(a[r, li], b[r // group, li])
            want_acc[r, li] = om.reduce_add_mul_slice_assign(want_acc[r, li].copy(), a[r, li], b[r // group, li])
    out = torch.empty_like(_dev(a))
    P.slice_op_bcast(P.api.OP_MUL, moduli, _dev(a), _dev(b), out, n, group, bits)
    assert np.array_equal(out.cpu().numpy().view(dt).reshape(a.shape), want_mul)
    dacc = _dev(acc.copy())
    P.slice_op_bcast(P.api.OP_ADD_MUL, moduli, _dev(a), _dev(b), dacc, n, group, bits)
    assert np.array_equal(dacc.cpu().numpy().view(dt).reshape(a.shape), want_acc)
