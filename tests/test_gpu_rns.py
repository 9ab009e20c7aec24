"""GPU parity for the RNS limb handling, the multi-word gadget basis and the multi-limb (L > 1) external product vs
the CPU oracle (oracle/, itself pinned in tests/test_oracle.py by the reference's deterministic RNS cases
primus_rns/tests/rns.rs:65-343 and the big-int model).  Everything here is bit-exact.
"""
import numpy as np
import pytest

from conftest import Q27, Q49, Q50, Q50B, Q60

pytestmark = pytest.mark.gpu

P27A, P27B = 134215681, 134176769      # primus_decompose/tests/big_uint.rs:21-28
# 8 limbs just below 2^50, each = 1 mod 2^15 (config C3)
C3_PRIMES = None


def _c3_primes():
    global C3_PRIMES
    if C3_PRIMES is None:
        def is_prime(n):
            if n % 2 == 0: return False
            d, s = n - 1, 0
            while d % 2 == 0: d //= 2; s += 1
            for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
                x = pow(a, d, n)
                if x in (1, n - 1): continue
                for _ in range(s - 1):
                    x = x * x % n
                    if x == n - 1: break
                else:
                    return False
            return True
        out, c = [], (1 << 50) - (1 << 15) + 1
        while len(out) < 8:
            if is_prime(c): out.append(c)
            c -= 1 << 15
        C3_PRIMES = out
    return C3_PRIMES


def _dev(x):
    import torch
    return torch.from_numpy(x.view(np.int64 if x.dtype == np.uint64 else np.int32)).cuda()


def _host(t, dt):
    return t.cpu().numpy().view(dt)


def _rand_res(rng, moduli, n, dt):
    r = np.stack([rng.integers(0, m, n, dtype=np.uint64).astype(dt) for m in moduli])
    r[:, 0] = 0
    r[:, 1] = np.array([m - 1 for m in moduli], dtype=dt)
    return r


CASES = [(32, [P27A, P27B]), (64, [Q50, Q50B]), (64, [Q50, Q50B, Q49]), (64, [Q50]), (32, [Q27]), (64, "c3"),
         (64, [17, 19, 23]), (32, [3, 5, 7, 11, 13]), (64, [Q60, Q50, Q49, Q50B])]   # value_len < limbs; a 60-bit limb


@pytest.mark.parametrize("bits,moduli", CASES)
def test_rns_constructor_compose_decompose(bits, moduli):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = _c3_primes() if moduli == "c3" else moduli
    dt = np.uint64 if bits == 64 else np.uint32
    g, o = P.RNSBase(moduli, bits), O.RNSBase(moduli, bits)
    assert g.moduli_count() == o.moduli_count() and g.big_uint_value_len() == o.big_uint_value_len()
    assert g.moduli_product() == o.moduli_product()
    rng = np.random.default_rng(5)
    n = 1000
    res = _rand_res(rng, moduli, n, dt)
    want = o.compose_multiple_values_to(res.reshape(-1), n)
    big = torch.empty(n * g.big_uint_value_len(), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    g.compose_multiple_values_to(_dev(res), big)
    assert np.array_equal(_host(big, dt), want)
    back = torch.empty_like(_dev(res))
    g.decompose_big_uint_values_to(big, back)
    assert np.array_equal(_host(back, dt).reshape(res.shape), res)
    # values >= Q are legal decompose inputs (base.rs:457-481 reduces word by word)
    junk = rng.integers(0, 1 << 64, n * g.big_uint_value_len(), dtype=np.uint64).astype(dt)
    g.decompose_big_uint_values_to(_dev(junk), back)
    assert np.array_equal(_host(back, dt).reshape(-1), o.decompose_big_uint_values_to(junk, n))


def test_rns_constructor_errors():
    import primus_fhe_b200 as P
    with pytest.raises(P.PfheError) as e:
        P.RNSBase([])
    assert e.value.name == "EmptyBase"                       # primus_rns/tests/rns.rs:67-70
    with pytest.raises(P.PfheError) as e:
        P.RNSBase([21, 35])
    assert e.value.name == "CoPrimeError"                    # rns.rs:74-77
    assert P.RNSBase([3, 5, 7]).moduli_product() == 105      # rns.rs:81-100


@pytest.mark.parametrize("bits,moduli,small_modulus", [(64, [Q50, Q50B], 5), (64, [Q50, Q50B, Q49], 128), (32, [P27A, P27B], 2),
                                                       (32, [P27A, P27B], 127), (64, "c3", 128)])
def test_lift_scaled_add_matches_oracle(bits, moduli, small_modulus):
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = _c3_primes() if moduli == "c3" else moduli
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(8)
    for n in (777, 776):   # scalar kernel (ragged length) / 16-byte vector kernel
        small = rng.integers(0, small_modulus, n, dtype=np.uint64).astype(dt)
        small[:small_modulus] = np.arange(small_modulus, dtype=dt)[:n]
        acc = _rand_res(rng, moduli, n, dt)
        scal = [int(rng.integers(0, m)) for m in moduli]
        scal[0] = 0
        want = O.RNSBase(moduli, bits).wrapping_decompose_small_values_scaled_add_to(small, acc.reshape(-1).copy(), small_modulus, scal)
        d = _dev(acc.copy())
        P.RNSBase(moduli, bits).wrapping_decompose_small_values_scaled_add_to(_dev(small), d, small_modulus, scal)
        assert np.array_equal(_host(d, dt).reshape(-1), want)


@pytest.mark.parametrize("bits,moduli,beta,rev", [(32, [P27A, P27B], 7, None), (64, [Q50, Q50B], 7, None), (64, [Q50, Q50B, Q49], 7, 5),
                                                  (64, [Q50], 7, None), (32, [Q27], 7, None), (64, [Q50, Q50B], 16, None),
                                                  (64, [Q50, Q50B], 1, None), (64, [Q50, Q50B], 1, 99), (64, "c3", 7, None),
                                                  (64, "c3", 30, 4), (32, [P27A, P27B], 31, None)])
def test_gadget_digits_match_oracle(bits, moduli, beta, rev):
    """compose -> init_value_carry -> unsigned digit per level -> centred lift, fused on the GPU, against the oracle's
    step-by-step restatement (big_integer/basis.rs:326-367, common.rs:275-325, base.rs:279-315)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = _c3_primes() if moduli == "c3" else moduli
    dt = np.uint64 if bits == 64 else np.uint32
    L = len(moduli)
    orns = O.RNSBase(moduli, bits); obb = O.BigUintApproxSignedBasis(orns, beta, rev)
    grns = P.RNSBase(moduli, bits); gbb = P.BigUintApproxSignedBasis(grns, beta, rev)
    assert gbb.decompose_length() == obb.decompose_length() and gbb.drop_bits() == obb.drop_bits()
    levels = obb.decompose_length()
    rng = np.random.default_rng(13)
    n, polys = 64, 3
    res = np.stack([_rand_res(rng, moduli, n, dt) for _ in range(polys)])          # [polys][L][n]
    want = np.empty((polys, levels, L, n), dtype=dt)
    for p in range(polys):
        big = orns.compose_multiple_values_to(res[p].reshape(-1), n)
        car = obb.init_value_carry_slice_inplace(big)
        for l in range(levels):
            dig = obb.unsigned_decompose_slice_to(l, big, car)
            want[p, l] = orns.wrapping_decompose_small_values_to(dig, 1 << beta).reshape(L, n)
    out = torch.empty(polys * levels * L * n, dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    gbb.gadget_decompose_batch(_dev(res), out, n)
    assert np.array_equal(_host(out, dt).reshape(want.shape), want)


@pytest.mark.parametrize("bits,moduli,log_n,beta,rev,k", [(64, [Q50, Q50B], 10, 9, 4, 1), (64, [Q50, Q50B, Q49], 10, 12, 3, 2),
                                                          (32, [P27A, P27B], 10, 7, None, 1), (64, [Q50, Q50B], 11, 7, None, 1),
                                                          (64, [Q50], 10, 7, None, 1), (64, "c3", 10, 25, None, 1),
                                                          (64, [Q50, Q50B], 4, 9, 4, 1),
                                                          # composed values of three / four words in the single fused kernel (k = 1)
                                                          (64, [Q50, Q50B, Q49], 10, 7, None, 1), (64, [Q50, Q50B, Q49], 11, 13, 3, 1),
                                                          (64, [Q50, Q50B, Q49, Q60], 10, 11, None, 1), (64, [Q50, Q50B, Q49, 1125899904679937], 10, 16, None, 1),
                                                          (32, [P27A, P27B, Q27], 10, 7, None, 1), (32, [P27A, P27B, Q27, 268369921], 10, 7, 11, 1)])
@pytest.mark.parametrize("to_coeff", [True, False])
def test_dcrt_external_product_matches_oracle(bits, moduli, log_n, beta, rev, k, to_coeff):
    """CrtGlwe::mul_dcrt_ggsw_to for L >= 1 limbs (primus_lattice/src/glwe/crt.rs:200-227); parity unpinned by any
    reference test -- the oracle is pinned by the schoolbook identity in tests/test_oracle.py."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = _c3_primes() if moduli == "c3" else moduli
    dt = np.uint64 if bits == 64 else np.uint32
    n, L = 1 << log_n, len(moduli)
    odc = O.DcrtTable(log_n, moduli, bits); orns = O.RNSBase(moduli, bits); obb = O.BigUintApproxSignedBasis(orns, beta, rev)
    gdc = (P.U64DcrtTable if bits == 64 else P.U32DcrtTable)(log_n, moduli)
    gbb = P.BigUintApproxSignedBasis(P.RNSBase(moduli, bits), beta, rev)
    levels = obb.decompose_length()
    rng = np.random.default_rng(33)
    batch = 5
    key = np.stack([_rand_res(rng, moduli, n, dt) for _ in range((k + 1) * levels * (k + 1))]).reshape(-1)
    cin = np.stack([_rand_res(rng, moduli, n, dt) for _ in range(batch * (k + 1))]).reshape(batch, -1)
    want = O.external_product(odc, orns, obb, k, key, cin, to_coeff=to_coeff, batch=batch)
    out = torch.empty_like(_dev(cin))
    P.dcrt_external_product_batch(gdc, gbb, k, _dev(key), _dev(cin), out, to_coeff)
    assert np.array_equal(_host(out, dt).reshape(want.shape), want)
    # a scratch buffer that only fits ONE ciphertext forces the chunked path
    one = torch.empty(2 * levels * L * n * (k + 1) * (bits // 8) // 2, dtype=torch.uint8, device="cuda")
    out2 = torch.zeros_like(out)
    P.dcrt_external_product_batch(gdc, gbb, k, _dev(key), _dev(cin), out2, to_coeff, scratch=one)
    assert torch.equal(out, out2)
    if L == 1 and log_n >= 10:   # the fused single-modulus kernel computes the same function
        gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, moduli[0])
        out3 = torch.empty_like(out)
        gt.external_product_batch(k, beta, rev, _dev(key), _dev(cin), out3, to_coeff)
        assert torch.equal(out, out3)


@pytest.mark.parametrize("bits,moduli,log_n", [(64, [Q50], 10), (32, [Q27], 4), (64, [Q50, Q50B, Q49], 6), (32, [P27A, P27B], 11)])
def test_mul_monomial_matches_oracle(bits, moduli, log_n):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    n, L = 1 << log_n, len(moduli)
    rng = np.random.default_rng(2)
    degs = np.array([0, 1, n - 1, n, n + 1, 2 * n - 1, 2 * n, 2 * n + 5] + [int(rng.integers(0, 2 * n)) for _ in range(8)], dtype=np.uint32)
    batch = len(degs)
    polys = np.stack([_rand_res(rng, moduli, n, dt) for _ in range(batch)])        # [batch][L][n]
    want = np.empty_like(polys)
    for b in range(batch):
        for li, m in enumerate(moduli):
            want[b, li] = O.mul_monomial(polys[b, li], int(degs[b]) % (2 * n), m, bits)
    out = torch.empty_like(_dev(polys))
    P.mul_monomial_batch(moduli, _dev(degs), _dev(polys), out, log_n, bits)
    assert np.array_equal(_host(out, dt).reshape(want.shape), want)


@pytest.mark.parametrize("bits,q,n", [(64, Q50, 4096), (32, Q27, 1024), (64, Q50, 7), (32, Q27, 100), (64, 1152921504606830593, 2048)])
def test_dot_product_matches_oracle(bits, q, n):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(4)
    rows = 6
    a = rng.integers(0, q, (rows, n), dtype=np.uint64).astype(dt)
    b = rng.integers(0, q, (rows, n), dtype=np.uint64).astype(dt)
    a[0, :] = q - 1; b[0, :] = q - 1
    om = O.BarrettModulus(q, bits)
    want = np.array([om.reduce_dot_product(a[r], b[r]) for r in range(rows)], dtype=dt)
    out = torch.empty(rows, dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    P.dot_product_batch(q, _dev(a), _dev(b), out, n, bits)
    assert np.array_equal(_host(out, dt), want)


@pytest.mark.parametrize("bits,in_m,out_m", [(64, [17, 19, 23], [29, 31]), (64, [Q50, Q50B, Q49], [1152921504606830593, 1125899904679937]),
                                             (32, [P27A, P27B], [Q27, 268369921]), (64, "c3", [Q50B, Q49, Q50]), (64, [Q50], [Q50B, Q49]),
                                             (64, [Q50, Q50B], [Q49])])
def test_base_converter_matches_oracle(bits, in_m, out_m):
    """BaseConverter::fast_convert_array / exact_convert_array (primus_rns/src/converter.rs:186-365); the oracle is pinned on
    the reference's deterministic case primus_rns/tests/rns.rs:281-345 (tests/test_oracle.py)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    in_m = _c3_primes() if in_m == "c3" else in_m
    dt = np.uint64 if bits == 64 else np.uint32
    tdt = torch.int64 if bits == 64 else torch.int32
    rng = np.random.default_rng(44)
    n, polys = 256, 3
    cin = np.stack([_rand_res(rng, in_m, n, dt) for _ in range(polys)])                 # [polys][L_in][n]
    oc, gc = O.BaseConverter(in_m, out_m, bits), P.BaseConverter(in_m, out_m, bits)
    want = np.stack([oc.fast_convert_array(cin[p].reshape(-1), n).reshape(len(out_m), n) for p in range(polys)])
    out = torch.empty(polys * len(out_m) * n, dtype=tdt, device="cuda")
    gc.fast_convert_array(_dev(cin), out, n)
    assert np.array_equal(_host(out, dt).reshape(want.shape), want)
    o1, g1 = O.BaseConverter(in_m, out_m[:1], bits), P.BaseConverter(in_m, out_m[:1], bits)
    want1 = np.stack([o1.exact_convert_array(cin[p].reshape(-1), n) for p in range(polys)])
    out1 = torch.empty(polys * n, dtype=tdt, device="cuda")
    g1.exact_convert_array(_dev(cin), out1, n)
    assert np.array_equal(_host(out1, dt).reshape(want1.shape), want1)
    if len(out_m) > 1:
        with pytest.raises(P.PfheError):
            gc.exact_convert_array(_dev(cin), out1, n)
    # words that are not canonical residues (any word value: the reference reduces them with the Shoup / Barrett product), the extreme
    # residues 0 and q_i - 1, and a polynomial length that is not a power of two (index split by division)
    n2 = 100
    wild = rng.integers(0, 1 << bits, (polys, len(in_m), n2), dtype=np.uint64).astype(dt)
    wild[0, :, :4] = 0
    for i, m in enumerate(in_m):
        wild[0, i, 4:8] = m - 1
    wild[1] = np.stack([rng.integers(0, m, n2, dtype=np.uint64).astype(dt) for m in in_m])
    want = np.stack([oc.fast_convert_array(wild[p].reshape(-1), n2).reshape(len(out_m), n2) for p in range(polys)])
    out = torch.empty(polys * len(out_m) * n2, dtype=tdt, device="cuda")
    gc.fast_convert_array(_dev(wild), out, n2)
    assert np.array_equal(_host(out, dt).reshape(want.shape), want)
    want1 = np.stack([o1.exact_convert_array(wild[p].reshape(-1), n2) for p in range(polys)])
    out1 = torch.empty(polys * n2, dtype=tdt, device="cuda")
    g1.exact_convert_array(_dev(wild), out1, n2)
    assert np.array_equal(_host(out1, dt).reshape(want1.shape), want1)


def test_base_converter_reference_case():
    """The reference's own deterministic vectors (primus_rns/tests/rns.rs:281-345), straight through the C-ABI."""
    import torch
    import primus_fhe_b200 as P
    in_m, out_m = [17, 19, 23], [29, 31]
    by_value = [[0, 0, 0], [1, 2, 3], [16, 18, 22], [7, 11, 13], [4, 0, 19]]
    n = len(by_value)
    crt_in = np.array([[r[i] for r in by_value] for i in range(3)], dtype=np.uint64)
    Q = 17 * 19 * 23
    out = torch.empty(2 * n, dtype=torch.int64, device="cuda")
    P.BaseConverter(in_m, out_m).fast_convert_array(_dev(crt_in), out, n)
    got = _host(out, np.uint64).reshape(2, n)
    for vi, res in enumerate(by_value):
        y = [res[i] * pow(Q // in_m[i], -1, in_m[i]) % in_m[i] for i in range(3)]
        for k, p in enumerate(out_m):
            assert int(got[k, vi]) == sum(y[i] * ((Q // in_m[i]) % p) for i in range(3)) % p
    vals = [0, 1, 2, 7, 16]
    ex_in = np.array([vals] * 3, dtype=np.uint64)
    ex = torch.empty(len(vals), dtype=torch.int64, device="cuda")
    P.BaseConverter(in_m, [37]).exact_convert_array(_dev(ex_in), ex, len(vals))
    assert [int(v) for v in _host(ex, np.uint64)] == [v % 37 for v in vals]
