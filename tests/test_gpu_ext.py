"""GPU parity for the round-2 C-ABI additions: bootstrapping-key handles + host-slice bootstrap, the multi-device drivers
(one process, several devices -- on a one-GPU box the same device twice), named whole-ciphertext transforms, the byte layout,
the modulus switch and UintNttTable<T>."""
import numpy as np
import pytest

from conftest import Q27, Q50, Q60

pytestmark = pytest.mark.gpu


def _tables(bits, log_n, q):
    import primus_fhe_b200 as P
    from oracle import oracle as O
    return (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q), (O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q)


def _devices():
    import primus_fhe_b200 as P
    return [0, 1] if P.device_count() >= 2 else [0, 0]


@pytest.mark.parametrize("bits,q,n_lwe,batch", [(32, Q27, 24, 9), (64, Q50, 6, 5)])
def test_bootstrap_slices_match_oracle(bits, q, n_lwe, batch):
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    gt, ot = _tables(bits, 10, q)
    n = 1024
    ob = O.ApproxSignedBasis(q, 7, None, bits); lv = ob.decompose_length()
    rng = np.random.default_rng(201)
    bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(dt)
    lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    acc = O.blind_rotate(ot, ob, bsk, n_lwe, lwe, tv, batch=batch)
    want = np.stack([O.extract_lwe(a, q, bits) for a in acc])
    key = P.BootstrappingKey(gt, 7, None, n_lwe, bsk)
    assert key.levels == lv
    got = key.bootstrap_slices(lwe, tv)
    assert np.array_equal(got, want)
    assert np.array_equal(key.bootstrap_slices(lwe, tv, extract=False), acc)
    # the serialised form of the key (to_bytes) uploads to the same key
    key2 = P.BootstrappingKey(gt, 7, None, n_lwe, bsk.tobytes())
    assert np.array_equal(key2.bootstrap_slices(lwe, tv), want)
    with pytest.raises(P.PfheError):
        P.BootstrappingKey(gt, 7, None, n_lwe, bsk[:-1].tobytes())


def test_multi_device_drivers_match_oracle():
    import primus_fhe_b200 as P
    from oracle import oracle as O
    devs = _devices()
    rng = np.random.default_rng(202)
    # transforms + fused product, u64 N = 4096, ragged shards (batch 7 over 2 parts)
    q, log_n = Q50, 12
    n = 1 << log_n
    mt = P.MultiNttTable(log_n, q, devs, 64)
    ot = O.U64NttTable(log_n, q)
    x = rng.integers(0, q, (7, n), dtype=np.uint64); y = rng.integers(0, q, (7, n), dtype=np.uint64)
    want = x.copy(); ot.forward_batch(want)
    got = x.copy(); mt.transform_slices(got)
    assert np.array_equal(got, want)
    mt.inverse_transform_slices(got)
    assert np.array_equal(got, x)
    c = np.empty_like(x); mt.polymul_slices(x, y, c)
    assert np.array_equal(c, ot.polymul_batch(x, y))
    one = x[:1].copy(); mt.transform_slices(one)          # fewer units than devices: the empty shard is a no-op
    assert np.array_equal(one, want[:1])
    # external product + bootstrap, u32 N = 1024
    q32, n32 = Q27, 1024
    m32 = P.MultiNttTable(10, q32, devs, 32)
    o32 = O.U32NttTable(10, q32)
    ob = O.ApproxSignedBasis(q32, 7, None, 32); lv = ob.decompose_length()
    key = rng.integers(0, q32, 2 * lv * 2 * n32, dtype=np.uint64).astype(np.uint32)
    cin = rng.integers(0, q32, (5, 2 * n32), dtype=np.uint64).astype(np.uint32)
    out = np.empty_like(cin)
    m32.external_product_slices(1, 7, None, key, cin, out, True)
    assert np.array_equal(out, O.external_product_single(o32, ob, 1, key, cin, to_coeff=True, batch=5))
    n_lwe = 10
    bsk = rng.integers(0, q32, n_lwe * 2 * lv * 2 * n32, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n32, (5, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q32, n32, dtype=np.uint64).astype(np.uint32)
    keys = m32.bootstrapping_keys(7, None, n_lwe, bsk)
    acc = O.blind_rotate(o32, ob, bsk, n_lwe, lwe, tv, batch=5)
    assert np.array_equal(m32.bootstrap_slices(keys, lwe, tv, extract=False), acc)
    assert np.array_equal(m32.bootstrap_slices(keys, lwe, tv), np.stack([O.extract_lwe(a, q32, 32) for a in acc]))


def test_named_ciphertext_transforms_and_bytes():
    import primus_fhe_b200 as P
    rng = np.random.default_rng(203)
    for bits, q in ((32, Q27), (64, Q50)):
        dt = np.uint64 if bits == 64 else np.uint32
        gt, ot = _tables(bits, 10, q)
        n = 1024
        for shape, dims, words in (("rlwe", (), 2 * n), ("rlev", (3,), 3 * 2 * n), ("rgsw", (3,), 2 * 3 * 2 * n), ("glwe", (2,), 3 * n),
                                   ("glev", (2, 4), 4 * 3 * n), ("ggsw", (2, 2), 3 * 2 * 3 * n)):
            assert P.cipher_words(gt, shape, *dims) == words
            data = rng.integers(0, q, words, dtype=np.uint64).astype(dt)
            want = data.copy().reshape(-1, n); ot.forward_batch(want)
            got = data.copy()
            P.into_ntt_form(gt, shape, got, *dims)
            assert np.array_equal(got.reshape(-1, n), want), shape
            dst = np.zeros_like(data)
            P.write_ntt_form(gt, data, dst)
            assert np.array_equal(dst, got)
            P.into_coeff_form(gt, shape, got, *dims)
            assert np.array_equal(got, data)
            back = np.zeros_like(data); P.write_coeff_form(gt, dst, back)
            assert np.array_equal(back, data)
            # byte layout = raw little-endian words (bytemuck::cast_slice)
            b = P.to_bytes(data, bits)
            assert b == data.astype(data.dtype.newbyteorder("<")).tobytes()
            assert np.array_equal(P.from_bytes(b, bits), data)
        with pytest.raises(P.PfheError):
            P.from_bytes(b"\x00" * 7, bits)
    # CRT containers through the DCRT table
    from conftest import Q50B
    from oracle import oracle as O
    mods = [Q50, Q50B]
    dc = P.U64DcrtTable(11, mods)
    data = np.stack([rng.integers(0, m, (3, 2048), dtype=np.uint64) for m in mods], axis=1).copy()   # CrtGlwe k = 2: [3][L][N]
    want = data.copy()
    for i, m in enumerate(mods):
        x = np.ascontiguousarray(want[:, i]); O.U64NttTable(11, m).forward_batch(x); want[:, i] = x
    got = data.copy(); P.dcrt_into_ntt_form(dc, got)
    assert np.array_equal(got, want)
    P.dcrt_into_coeff_form(dc, got)
    assert np.array_equal(got, data)


def test_modulus_switch_matches_convention():
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    rng = np.random.default_rng(204)
    for bits, q, log_2n in ((32, Q27, 11), (64, Q50, 11), (64, Q60, 12), (32, 1073692673, 13)):
        dt = np.uint64 if bits == 64 else np.uint32
        v = rng.integers(0, q, 5000, dtype=np.uint64).astype(dt)
        v[:6] = (0, 1, q - 1, q // 2, q // 2 + 1, (q + (1 << log_2n)) // (1 << (log_2n + 1)))
        d = torch.from_numpy(v.view(np.int64 if bits == 64 else np.int32)).cuda()
        out = torch.empty(v.size, dtype=torch.int32, device="cuda")
        P.modulus_switch_batch(q, log_2n, d, out, bits)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), O.modulus_switch(v, q, log_2n))


def test_uint_ntt_table_words_and_rules():
    import primus_fhe_b200 as P
    from oracle import oracle as O
    rng = np.random.default_rng(205)
    # canonical results equal the fast tables' (prime64/tests.rs:78-237 cross-implementation check), all three word types
    for bits, q, log_n, obits in ((16, 12289, 10, 32), (16, 257, 5, 32), (32, Q27, 10, 32), (32, 1073692673, 12, 32), (64, Q50, 11, 64), (64, Q60, 10, 64)):
        dt = {16: np.uint16, 32: np.uint32, 64: np.uint64}[bits]
        n = 1 << log_n
        ut = P.UintNttTable(log_n, q, bits)
        ot = (O.U64NttTable if obits == 64 else O.U32NttTable)(log_n, q)
        assert ut.poly_length() == n and ut.root() == ot.root() and ut.inv_root() == ot.inv_root()
        x = rng.integers(0, q, (3, n), dtype=np.uint64)
        x[0, :] = q - 1
        want = x.astype(np.uint64 if obits == 64 else np.uint32); ot.forward_batch(want)
        got = x.astype(dt); ut.transform_slices(got)
        assert np.array_equal(got.astype(np.uint64), want.astype(np.uint64)), (bits, q)
        ut.inverse_transform_slices(got)
        assert np.array_equal(got.astype(np.uint64), x)
        if bits != 16:
            lazy = (x + np.uint64(q) * rng.integers(0, 4, (3, n), dtype=np.uint64)).astype(dt); ut.lazy_transform_slice(lazy)
            assert np.array_equal(lazy.astype(np.uint64), want.astype(np.uint64))
    for bits, q, log_n, name in ((16, 12289, 13, "NoPrimitiveRoot"), (32, Q27, 21, "NoPrimitiveRoot"), (16, 40961, 10, "ModulusTooLarge"),
                                 (32, 3221225473, 10, "ModulusTooLarge"), (64, 97, 7, "NoPrimitiveRoot")):
        with pytest.raises(P.PfheError) as e:
            P.UintNttTable(log_n, q, bits)
        assert e.value.name == name, (bits, q, log_n, e.value.name)


@pytest.mark.parametrize("log_basis,levels,n_lwe,batch", [(7, None, 40, 5), (4, 5, 6, 3), (9, None, 6, 3)])
def test_ternary_blind_rotation_matches_oracle(log_basis, levels, n_lwe, batch):
    """pfhe_blind_rotate_ternary32_batch == the oracle's composition (monomial NTTs + slice ops + external product), bit for bit."""
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    q, n = Q27, 1024
    gt, ot = P.U32NttTable(10, q), O.U32NttTable(10, q)
    ob = O.ApproxSignedBasis(q, log_basis, levels, 32); lv = ob.decompose_length()
    rng = np.random.default_rng(206)
    bp = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    bm = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    lwe[0, :] = 0
    lwe[1, :] = 2 * n - 1
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    want = O.blind_rotate_ternary(ot, ob, bp, bm, n_lwe, lwe, tv, batch=batch)
    d = lambda x: torch.from_numpy(x.view(np.int32)).cuda()
    out = torch.empty((batch, 2 * n), dtype=torch.int32, device="cuda")
    gt.blind_rotate_ternary_batch(log_basis, levels, d(bp), d(bm), n_lwe, d(lwe), d(tv), out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    with pytest.raises(P.PfheError):   # other shapes: Unsupported, stated in pfhe.h
        t11 = P.U32NttTable(11, q)
        t11.blind_rotate_ternary_batch(log_basis, levels, d(bp), d(bm), 1, d(lwe[:1, :2].copy()), d(np.zeros(2048, np.uint32)),
                                       torch.empty((1, 4096), dtype=torch.int32, device="cuda"))


def test_registered_host_buffer_takes_the_direct_path():
    """pfhe_host_register: a caller-owned pageable buffer page-locked in place is no longer staged; results are identical."""
    import primus_fhe_b200 as P
    from oracle import oracle as O
    q, log_n, batch = 1125899906826241, 12, 600      # 19 MiB: above the staging threshold
    rng = np.random.default_rng(21)
    a = rng.integers(0, q, (batch, 1 << log_n), dtype=np.uint64)
    want = a[:4].copy()
    O.U64NttTable(log_n, q).forward_batch(want)
    t = P.U64NttTable(log_n, q)
    staged = a.copy()
    assert P.host_is_pageable(staged)
    t.transform_slices(staged)
    direct = a.copy()
    with P.registered_host_buffer(direct):
        assert not P.host_is_pageable(direct)
        t.transform_slices(direct)
    assert P.host_is_pageable(direct)
    assert np.array_equal(staged, direct) and np.array_equal(direct[:4], want)
    with pytest.raises(P.PfheError):
        P.registered_host_buffer(np.empty(0, dtype=np.uint64)).__enter__()


def test_c_program_drives_the_library_without_python(tmp_path):
    """examples/c_abi_smoke.c: plain C against include/pfhe.h -- host slices round trip and a fused product, no torch in the process."""
    import subprocess
    from test_cabi import _build_c_example
    exe = _build_c_example(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "c-abi smoke ok" in p.stdout
