"""GPU parity: CUDA NTT path (through the C-ABI) vs the CPU oracle, bit-exact.

Mirrors the reference's own NTT tests: canonical forward/inverse equality against the second
implementation (primus_ntt/src/ntt/prime64/tests.rs:78-237, prime32/tests.rs:137-236), round trips for
N = 8..1024 (prime64/tests.rs), the integration primes of primus_ntt/tests/ntt.rs:16-127, monomial
transforms, and the lazy-range contracts (outputs congruent mod q and inside [0,4q) / [0,2q)).
"""
import numpy as np
import pytest

from conftest import Q27, Q28, Q29, Q30, Q49, Q50, Q60

pytestmark = pytest.mark.gpu


def _rand(rng, q, shape, dt):
    return rng.integers(0, q, size=shape, dtype=np.uint64).astype(dt)


def _tables(bits, log_n, q):
    import primus_fhe_b200 as P
    from oracle import oracle as O
    g = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    o = (O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q)
    return g, o


CASES = [(64, ln, Q50) for ln in range(1, 14)] + [(64, 14, 1125899904679937)] + [(64, ln, Q60) for ln in (3, 10, 11, 12, 13)] + \
        [(64, 11, Q29), (64, 11, Q49), (64, 12, Q30), (64, 10, Q27)] + \
        [(32, ln, Q27) for ln in range(1, 16)] + [(32, 12, Q28), (32, 11, Q29), (32, 12, Q30)]


@pytest.mark.parametrize("bits,log_n,q", CASES)
def test_forward_inverse_match_oracle(bits, log_n, q):
    import torch
    g, o = _tables(bits, log_n, q)
    assert g.root() == o.root() and g.inv_root() == o.inv_root() and g.inv_n() == o.inv_n()
    n, dt = 1 << log_n, (np.uint64 if bits == 64 else np.uint32)
    rng = np.random.default_rng(1000 * bits + log_n)
    batch = 5 if log_n >= 12 else 37
    x = _rand(rng, q, (batch, n), dt)
    x[0, :] = q - 1  # extreme values
    x[1, :] = 0
    want = x.copy()
    o.forward_batch(want)
    tdt = torch.int64 if bits == 64 else torch.int32
    d = torch.from_numpy(x.view(np.int64 if bits == 64 else np.int32)).cuda()
    g.forward_batch(d)
    got = d.cpu().numpy().view(dt)
    assert np.array_equal(got, want)
    g.inverse_batch(d)
    assert np.array_equal(d.cpu().numpy().view(dt), x)
    # out-of-place variants
    src = torch.from_numpy(x.view(np.int64 if bits == 64 else np.int32)).cuda()
    dst = torch.empty_like(src)
    g.forward_batch_to(src, dst)
    assert np.array_equal(dst.cpu().numpy().view(dt), want) and np.array_equal(src.cpu().numpy().view(dt), x)
    g.inverse_batch_to(dst, src)
    assert np.array_equal(src.cpu().numpy().view(dt), x)
    assert tdt is not None


@pytest.mark.parametrize("bits,log_n,q", [(64, 12, Q50), (64, 11, Q60), (32, 10, Q27), (64, 4, Q50), (32, 5, Q27)])
def test_host_slice_trait_methods(bits, log_n, q):
    g, o = _tables(bits, log_n, q)
    n, dt = 1 << log_n, (np.uint64 if bits == 64 else np.uint32)
    rng = np.random.default_rng(7)
    x = _rand(rng, q, n, dt)
    want = x.copy(); o.transform_slice(want)
    y = x.copy(); g.transform_slice(y)
    assert np.array_equal(y, want)
    g.inverse_transform_slice(y)
    assert np.array_equal(y, x)
    # lazy contracts: inputs in [0,4q) forward / [0,2q) inverse; outputs congruent and in range
    # (prime64/tests.rs:97-106, :133-142)
    lz = (x.astype(object) + rng.integers(0, 4, n) * q).astype(dt)
    yl = lz.copy(); g.lazy_transform_slice(yl)
    assert (yl < 4 * q).all() and np.array_equal(yl.astype(object) % q, want.astype(object))
    iz = (want.astype(object) + rng.integers(0, 2, n) * q).astype(dt)
    g.lazy_inverse_transform_slice(iz)
    assert (iz < 2 * q).all() and np.array_equal(iz.astype(object) % q, x.astype(object))
    # many slices in one call
    xs = _rand(rng, q, (19, n), dt)
    ws = xs.copy(); o.forward_batch(ws)
    ys = xs.copy(); g.transform_slices(ys)
    assert np.array_equal(ys, ws)
    g.inverse_transform_slices(ys)
    assert np.array_equal(ys, xs)


@pytest.mark.parametrize("bits,log_n,q", [(64, 12, Q50), (64, 5, Q50), (32, 10, Q27), (32, 12, Q28)])
def test_monomial_transforms(bits, log_n, q):
    import torch
    g, o = _tables(bits, log_n, q)
    n, dt = 1 << log_n, (np.uint64 if bits == 64 else np.uint32)
    for coeff in (0, 1, q - 1, 123457 % q):
        for degree in (0, 1, 3, n // 2, n - 1):
            assert np.array_equal(g.transform_monomial(coeff, degree), o.transform_monomial(coeff, degree))
    for degree in (0, 2, n - 1):
        assert np.array_equal(g.transform_coeff_one_monomial(degree), o.transform_coeff_one_monomial(degree))
        assert np.array_equal(g.transform_coeff_minus_one_monomial(degree), o.transform_coeff_minus_one_monomial(degree))
    # device batch, degrees up to 2N-1 (X^(N+d) = -X^d)
    degs = np.array([0, 1, n - 1, n, n + 5, 2 * n - 1], dtype=np.uint32)
    out = torch.empty((len(degs), n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    g.monomial_batch(1, torch.from_numpy(degs.view(np.int32)).cuda(), out)
    got = out.cpu().numpy().view(dt)
    for i, d in enumerate(degs):
        if d < n:
            want = o.transform_monomial(1, int(d))
        else:
            want = o.transform_monomial(q - 1, int(d - n))
        assert np.array_equal(got[i], want)


@pytest.mark.parametrize("bits,log_n,q", [(64, 12, Q50), (64, 13, Q50), (64, 10, Q60), (64, 6, Q50), (32, 10, Q27), (32, 11, Q27), (32, 4, Q27), (64, 14, 1125899904679937), (32, 13, Q27), (64, 13, Q60)])
def test_polymul_matches_oracle_and_schoolbook(bits, log_n, q):
    import torch
    from oracle import oracle as O
    g, o = _tables(bits, log_n, q)
    n, dt = 1 << log_n, (np.uint64 if bits == 64 else np.uint32)
    rng = np.random.default_rng(99 + log_n)
    batch = 3 if log_n >= 12 else 9
    a, b = _rand(rng, q, (batch, n), dt), _rand(rng, q, (batch, n), dt)
    want = o.polymul_batch(a, b)
    if log_n <= 10:
        assert np.array_equal(want[0], O.naive_mul(a[0], b[0], q, bits))
    sdt = np.int64 if bits == 64 else np.int32
    da, db = torch.from_numpy(a.view(sdt)).cuda(), torch.from_numpy(b.view(sdt)).cuda()
    dc = torch.empty_like(da)
    g.polymul_batch(da, db, dc)
    assert np.array_equal(dc.cpu().numpy().view(dt), want)
    # aliasing c = a and c = b is allowed; a = b is a squaring
    g.polymul_batch(da, db, da)
    assert np.array_equal(da.cpu().numpy().view(dt), want)
    da = torch.from_numpy(a.view(sdt)).cuda()
    g.polymul_batch(da, db, db)
    assert np.array_equal(db.cpu().numpy().view(dt), want)
    sq = o.polymul_batch(a, a)
    g.polymul_batch(da, da, dc)
    assert np.array_equal(dc.cpu().numpy().view(dt), sq)
    g.polymul_batch(da, da, da)
    assert np.array_equal(da.cpu().numpy().view(dt), sq)
    # host shim
    assert np.array_equal(g.polymul_slices(a, b), want)


def test_constructor_errors():
    import primus_fhe_b200 as P
    with pytest.raises(P.PfheError) as e:
        P.U64NttTable(12, 1125899906842597)  # prime, but 2N does not divide q-1  (root.rs:72-81)
    assert e.value.name == "NoPrimitiveRoot"
    with pytest.raises(P.PfheError) as e:
        P.U32NttTable(10, 2013265921)  # 15*2^27+1 >= 2^30  (prime32/table.rs:195)
    assert e.value.name == "ModulusTooLarge"
    with pytest.raises(P.PfheError) as e:
        P.U64NttTable(10, (1 << 62) + 2049 * 0 + 0x1000000000001 * 0 + 4611686018427394049 - (1 << 62))  # q >= 2^62
    assert e.value.name in ("ModulusTooLarge", "NoPrimitiveRoot")


def test_dcrt_tables_match_per_limb_oracle():
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    moduli = [Q50, 1125899906629633, Q49]
    log_n, n = 11, 2048
    g = P.U64DcrtTable(log_n, moduli)
    o = O.DcrtTable(log_n, moduli)
    assert g.poly_length() == n and g.moduli_count() == 3 and g.crt_poly_length() == 3 * n
    rng = np.random.default_rng(5)
    batch = 4
    x = np.stack([np.stack([_rand(rng, m, n, np.uint64) for m in moduli]) for _ in range(batch)])
    want = x.copy()
    for bi in range(batch):
        o.transform_slice(want[bi].reshape(-1))
    d = torch.from_numpy(x.view(np.int64)).cuda()
    g.forward_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), want)
    g.inverse_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), x)
    # host trait method on one CRT polynomial
    y = x[0].copy().reshape(-1)
    g.transform_slice(y)
    assert np.array_equal(y, want[0].reshape(-1))
    # RNS polymul per limb
    a = x
    b = np.stack([np.stack([_rand(rng, m, n, np.uint64) for m in moduli]) for _ in range(batch)])
    wantc = np.empty_like(a)
    for li, t in enumerate(o.tables):
        wantc[:, li, :] = t.polymul_batch(np.ascontiguousarray(a[:, li, :]), np.ascontiguousarray(b[:, li, :]))
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    dc = torch.empty_like(da)
    g.polymul_batch(da, db, dc)
    assert np.array_equal(dc.cpu().numpy().view(np.uint64), wantc)
