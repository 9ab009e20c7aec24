"""GPU parity at the BASELINE.json shapes themselves (VERDICT r01 "weak" #1): full bootstrapping depth, the 8-limb N=16384 RNS
shape, multi-wave batches with odd sizes, the batch-4096 external product, mixed-width DCRT limbs, and host-slice calls that span
several pipeline chunks of pageable memory.  Where the oracle would take minutes the comparison is on sampled units: every unit
(polynomial, ciphertext) is independent, so the oracle result of a sampled unit does not depend on the rest of the batch.
"""
import numpy as np
import pytest

from conftest import Q27, Q50, Q50B, Q60

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.from_numpy(x.view(np.int64 if x.dtype == np.uint64 else np.int32)).cuda()


def c3_primes():
    """The eight largest primes below 2^50 with q = 1 mod 2^15."""
    want = [1125899904679937, 1125899903991809, 1125899903827969, 1125899903795201, 1125899903500289, 1125899903107073,
            1125899902124033, 1125899901665281]
    return want


# ---- C5: blind rotation at full depth --------------------------------------------------------------------------------
@pytest.mark.parametrize("q,log_basis,levels,n_lwe,batch", [
    (Q27, 7, None, 512, 6),     # BASELINE config 5: n = 512, N = 1024, base 2^7 (l = 3)
    (Q27, 4, 5, 16, 4),         # truncated basis (reverse_length), drop_bits 7
    (Q27, 9, None, 16, 4),      # drop_bits 0 (no rounding bit)
    (Q27, 1, 6, 16, 3),         # binary basis (unsigned digits)
    (Q27, 2, 7, 16, 3),         # 14 accumulated terms
    (134215681, 7, None, 16, 4),  # the other 27-bit prime of primus_decompose/tests/big_uint.rs:21
])
def test_blind_rotate_full_depth_u32(q, log_basis, levels, n_lwe, batch):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    n = 1024
    gt, ot = P.U32NttTable(10, q), O.U32NttTable(10, q)
    ob = O.ApproxSignedBasis(q, log_basis, levels, 32)
    lv = ob.decompose_length()
    rng = np.random.default_rng(77)
    bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    lwe[0, :] = 0
    lwe[1, :] = 2 * n - 1
    lwe[2, :] = n
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    tv[:3] = (0, q - 1, 1)
    want = O.blind_rotate(ot, ob, bsk, n_lwe, lwe, tv, batch=batch)
    out = torch.empty((batch, 2 * n), dtype=torch.int32, device="cuda")
    gt.blind_rotate_batch(log_basis, levels, _dev(bsk), n_lwe, _dev(lwe), _dev(tv), out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)


def test_blind_rotate_full_depth_u64():
    """n = 512 on the FP64-pipe path (q < 2^50) and a short run on the integer path (60-bit prime)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    for q, n_lwe, batch in ((Q50, 512, 3), (Q60, 24, 3)):
        n = 1024
        gt, ot = P.U64NttTable(10, q), O.U64NttTable(10, q)
        ob = O.ApproxSignedBasis(q, 7, None, 64)
        lv = ob.decompose_length()
        rng = np.random.default_rng(78)
        bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64)
        lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
        lwe[0, :] = 2 * n - 1
        tv = rng.integers(0, q, n, dtype=np.uint64)
        want = O.blind_rotate(ot, ob, bsk, n_lwe, lwe, tv, batch=batch)
        out = torch.empty((batch, 2 * n), dtype=torch.int64, device="cuda")
        gt.blind_rotate_batch(7, None, _dev(bsk), n_lwe, _dev(lwe), _dev(tv), out)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want), q


def test_blind_rotate_multi_wave_batch_sampled():
    """1250 ciphertexts (the per-GPU share of config 5: 2.1 waves of resident CTAs), 32 CMux steps, sampled rows vs the oracle."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    q, n, n_lwe, batch = Q27, 1024, 32, 1250
    gt, ot = P.U32NttTable(10, q), O.U32NttTable(10, q)
    ob = O.ApproxSignedBasis(q, 7, None, 32)
    lv = ob.decompose_length()
    rng = np.random.default_rng(79)
    bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    out = torch.empty((batch, 2 * n), dtype=torch.int32, device="cuda")
    gt.blind_rotate_batch(7, None, _dev(bsk), n_lwe, _dev(lwe), _dev(tv), out)
    rows = np.unique(np.concatenate([[0, 1, 591, 592, 593, 1183, 1184, 1249], rng.integers(0, batch, 24)]))
    want = O.blind_rotate(ot, ob, bsk, n_lwe, np.ascontiguousarray(lwe[rows]), tv, batch=len(rows))
    assert np.array_equal(out.cpu().numpy().view(np.uint32)[rows], want)


# ---- C3: 8 limbs x N = 16384 -------------------------------------------------------------------------------------------
def test_c3_dcrt_8_limbs_n16384():
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    mods = c3_primes()
    n, L, batch = 16384, 8, 3
    gt = P.U64DcrtTable(14, mods)
    ots = [O.U64NttTable(14, m) for m in mods]
    rng = np.random.default_rng(80)
    a = np.stack([rng.integers(0, m, (batch, n), dtype=np.uint64) for m in mods], axis=1)  # [batch][L][n]
    b = np.stack([rng.integers(0, m, (batch, n), dtype=np.uint64) for m in mods], axis=1)
    a[0, :, :4] = 0
    for i, m in enumerate(mods):
        a[1, i, :] = m - 1
    fwd = a.copy()
    inv = a.copy()
    mul = np.empty_like(a)
    for i, ot in enumerate(ots):
        x = np.ascontiguousarray(fwd[:, i]); ot.forward_batch(x); fwd[:, i] = x
        x = np.ascontiguousarray(inv[:, i]); ot.inverse_batch(x); inv[:, i] = x
        mul[:, i] = ot.polymul_batch(np.ascontiguousarray(a[:, i]), np.ascontiguousarray(b[:, i]))
    d = _dev(np.ascontiguousarray(a)); gt.forward_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), fwd)
    gt.inverse_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), a)
    d = _dev(np.ascontiguousarray(a)); gt.inverse_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), inv)
    c = torch.empty_like(d)
    gt.polymul_batch(_dev(np.ascontiguousarray(a)), _dev(np.ascontiguousarray(b)), c)
    assert np.array_equal(c.cpu().numpy().view(np.uint64), mul)


def test_c3_rns_product_is_the_big_integer_product():
    """Independent of the oracle: the 8-limb N = 16384 RNS product, CRT-composed (RNSBase::compose on the GPU), equals the schoolbook negacyclic
    product of the composed operands mod Q = prod q_i at sampled output coefficients (exact big-integer arithmetic)."""
    import torch
    import primus_fhe_b200 as P
    mods = c3_primes()
    n, L = 16384, 8
    Q = 1
    for m in mods:
        Q *= m
    rng = np.random.default_rng(81)
    a = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])      # [L][n]
    b = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
    gt = P.U64DcrtTable(14, mods)
    c = torch.empty((1, L, n), dtype=torch.int64, device="cuda")
    gt.polymul_batch(_dev(a[None].copy()), _dev(b[None].copy()), c)
    rns = P.RNSBase(mods, 64)
    vl = rns.big_uint_value_len()

    def compose(res):                                                          # GPU compose -> python integers
        big = torch.empty((n, vl), dtype=torch.int64, device="cuda")
        rns.compose_multiple_values_to(_dev(np.ascontiguousarray(res)), big)
        w = big.cpu().numpy().view(np.uint64)
        return [sum(int(w[i, k]) << (64 * k) for k in range(vl)) for i in range(n)]

    A, B_, C_ = compose(a), compose(b), compose(c.cpu().numpy().view(np.uint64)[0])
    for i in range(L):                                                         # the composed operands really are the CRT lifts
        assert A[5] % mods[i] == int(a[i, 5]) and B_[n - 1] % mods[i] == int(b[i, n - 1])
    for j in (0, 1, 8191, n - 1):
        acc = 0
        for i in range(n):
            k = j - i
            acc += A[i] * B_[k] if k >= 0 else -A[i] * B_[k + n]
        assert C_[j] == acc % Q, j


def test_dcrt_mixed_50_and_60_bit_limbs():
    """A 50-bit limb next to a 60-bit limb: the whole launch runs on the integer pipe (one field policy per launch)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    mods = [Q50, Q60, Q50B]
    for log_n in (11, 12):
        n, batch = 1 << log_n, 5
        gt = P.U64DcrtTable(log_n, mods)
        ots = [O.U64NttTable(log_n, m) for m in mods]
        rng = np.random.default_rng(81 + log_n)
        a = np.stack([rng.integers(0, m, (batch, n), dtype=np.uint64) for m in mods], axis=1)
        b = np.stack([rng.integers(0, m, (batch, n), dtype=np.uint64) for m in mods], axis=1)
        fwd, mul = a.copy(), np.empty_like(a)
        for i, ot in enumerate(ots):
            x = np.ascontiguousarray(fwd[:, i]); ot.forward_batch(x); fwd[:, i] = x
            mul[:, i] = ot.polymul_batch(np.ascontiguousarray(a[:, i]), np.ascontiguousarray(b[:, i]))
        d = _dev(np.ascontiguousarray(a)); gt.forward_batch(d)
        assert np.array_equal(d.cpu().numpy().view(np.uint64), fwd)
        gt.inverse_batch(d)
        assert np.array_equal(d.cpu().numpy().view(np.uint64), a)
        c = torch.empty_like(d)
        gt.polymul_batch(_dev(np.ascontiguousarray(a)), _dev(np.ascontiguousarray(b)), c)
        assert np.array_equal(c.cpu().numpy().view(np.uint64), mul)


def test_dcrt_lazy_contract_50_bit():
    """DcrtTable::lazy_transform_slice takes [0, 4 q_i) inputs, lazy_inverse_transform_slice [0, 2 q_i)
    (primus_ntt/src/dcrt/mod.rs:77-103); results must be congruent and in range (ours are canonical)."""
    import primus_fhe_b200 as P
    from oracle import oracle as O
    mods = [Q50, Q50B]
    n = 4096
    gt = P.U64DcrtTable(12, mods)
    rng = np.random.default_rng(83)
    canon = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
    for i, m in enumerate(mods):
        canon[i, 0] = m - 1
    for factor, fwd in ((4, True), (2, False)):
        lazy_in = canon.copy()
        for i, m in enumerate(mods):
            lazy_in[i] += np.uint64(m) * rng.integers(0, factor, n, dtype=np.uint64)
            lazy_in[i, 0] = factor * m - 1   # the largest value the lazy contract admits
        want = canon.copy()
        for i, m in enumerate(mods):
            ot = O.U64NttTable(12, m)
            (ot.transform_slice if fwd else ot.inverse_transform_slice)(want[i])
        got = np.ascontiguousarray(lazy_in)
        (gt.lazy_transform_slice if fwd else gt.lazy_inverse_transform_slice)(got)
        assert np.array_equal(got, want)


# ---- C2: batches of several waves, odd sizes, both field policies -------------------------------------------------------
@pytest.mark.parametrize("log_n,q,batch", [(12, Q50, 4099), (13, Q50, 4097), (12, Q60, 4099), (13, Q60, 4097), (11, Q50, 4101), (10, Q50, 4103)])
def test_c2_multi_wave_batches_sampled(log_n, q, batch):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    n = 1 << log_n
    gt, ot = P.U64NttTable(log_n, q), O.U64NttTable(log_n, q)
    rng = np.random.default_rng(84)
    x = rng.integers(0, q, (batch, n), dtype=np.uint64)
    y = rng.integers(0, q, (batch, n), dtype=np.uint64)
    rows = np.unique(np.concatenate([[0, 1, batch // 2, batch - 2, batch - 1], rng.integers(0, batch, 40)]))
    d = _dev(x.copy()); gt.forward_batch(d)
    want = np.ascontiguousarray(x[rows]); ot.forward_batch(want)
    got = d.cpu().numpy().view(np.uint64)
    assert np.array_equal(got[rows], want)
    gt.inverse_batch(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint64), x)
    c = torch.empty_like(d)
    gt.polymul_batch(_dev(x), _dev(y), c)
    want = ot.polymul_batch(np.ascontiguousarray(x[rows]), np.ascontiguousarray(y[rows]))
    assert np.array_equal(c.cpu().numpy().view(np.uint64)[rows], want)


# ---- C4: external product, batch 4096 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits,q", [(32, Q27), (64, Q50)])
def test_c4_external_product_batch_4096_sampled(bits, q):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    n, k, batch = 2048, 1, 4096
    gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(11, q)
    ot = (O.U64NttTable if bits == 64 else O.U32NttTable)(11, q)
    ob = O.ApproxSignedBasis(q, 7, None, bits)
    lv = ob.decompose_length()
    rng = np.random.default_rng(85)
    key = rng.integers(0, q, 2 * lv * 2 * n, dtype=np.uint64).astype(dt)
    cin = rng.integers(0, q, (batch, 2 * n), dtype=np.uint64).astype(dt)
    out = torch.empty((batch, 2 * n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    gt.external_product_batch(k, 7, None, _dev(key), _dev(cin), out, True)
    rows = np.unique(np.concatenate([[0, 1, 2047, 2048, 4095], rng.integers(0, batch, 27)]))
    want = O.external_product_single(ot, ob, k, key, np.ascontiguousarray(cin[rows]), to_coeff=True, batch=len(rows))
    assert np.array_equal(out.cpu().numpy().view(dt)[rows], want)


@pytest.mark.parametrize("bits,q,log_basis,levels", [(32, Q27, 1, None), (32, Q27, 2, None), (64, Q60, 3, None), (64, Q50, 2, None)])
def test_external_product_more_than_16_terms(bits, q, log_basis, levels):
    """comps * levels > 16: the lazy double-word sums are renormalised mid-way (reduce_dot_product's inner chunk of 16)."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    n, k, batch = 1024, 1, 3
    gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(10, q)
    ot = (O.U64NttTable if bits == 64 else O.U32NttTable)(10, q)
    ob = O.ApproxSignedBasis(q, log_basis, levels, bits)
    lv = ob.decompose_length()
    assert 2 * lv > 16
    rng = np.random.default_rng(86)
    key = rng.integers(0, q, 2 * lv * 2 * n, dtype=np.uint64).astype(dt)
    key[: 2 * n] = q - 1
    cin = rng.integers(0, q, (batch, 2 * n), dtype=np.uint64).astype(dt)
    cin[0, :] = q - 1
    want = O.external_product_single(ot, ob, k, key, cin, to_coeff=True, batch=batch)
    out = torch.empty((batch, 2 * n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    gt.external_product_batch(k, log_basis, levels, _dev(key), _dev(cin), out, True)
    assert np.array_equal(out.cpu().numpy().view(dt), want)


# ---- host-slice shim: pageable memory, several pipeline chunks ------------------------------------------------------------
def test_transform_slices_pageable_multi_chunk():
    """160 MB of pageable host memory through pfhe_ntt64_transform_slices (64 MiB pipeline chunks): what a Rust
    `&mut [u64]` (Vec) caller hands to the trait shim."""
    import primus_fhe_b200 as P
    from oracle import oracle as O
    q, n, batch = Q50, 4096, 5003
    gt, ot = P.U64NttTable(12, q), O.U64NttTable(12, q)
    rng = np.random.default_rng(87)
    x = rng.integers(0, q, (batch, n), dtype=np.uint64)
    orig = x.copy()
    gt.transform_slices(x)
    rows = np.unique(np.concatenate([[0, 2047, 2048, 4095, 4096, batch - 1], rng.integers(0, batch, 30)]))
    want = np.ascontiguousarray(orig[rows]); ot.forward_batch(want)
    assert np.array_equal(x[rows], want)
    gt.inverse_transform_slices(x)
    assert np.array_equal(x, orig)
