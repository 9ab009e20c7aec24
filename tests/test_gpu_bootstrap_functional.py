"""Functional (size-independent) check of the bootstrapping path at the BASELINE config-5 parameters (n = 512, N = 1024, q = 132120577,
base 2^7): real LWE / RLWE / RGSW encryptions with noise, modulus switch -> blind rotation -> sample extraction on the GPU, then decryption.
A programmable bootstrap must return an encryption of LUT[m] for every input message m, with small noise.  The same flow runs against the
oracle in tests/test_oracle.py (CPU)."""
import numpy as np
import pytest

import bootstrap_common as B

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint32)


def _times_secret(t, rows, z):
    import torch
    da = _dev(rows.astype(np.uint32))
    dz = _dev(np.broadcast_to(z.astype(np.uint32), rows.shape))
    dc = torch.empty_like(da)
    t.polymul_batch(da, dz, dc)
    return _host(dc)


def test_programmable_bootstrap_recovers_the_lookup_table_at_c5_parameters():
    import torch
    import primus_fhe_b200 as P
    t = P.U32NttTable(B.LOG_N, B.Q)
    basis = P.ApproxSignedBasis(B.Q, B.LOG_B, None, 32)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    rng = np.random.default_rng(20261018)
    z, s = B.secrets(rng)
    a1, e1 = B.rgsw_rows(rng, lv)
    key = B.assemble_key(a1, _times_secret(t, a1, z), e1, s, lv, drop)
    dkey = _dev(key.reshape(-1, B.N))
    t.forward_batch(dkey)                                            # NttRgsw form
    batch = 96
    msgs, lwe_q = B.lwe_inputs(rng, s, batch)
    lwe_2n = torch.empty((batch, B.N_LWE + 1), dtype=torch.int32, device="cuda")
    P.modulus_switch_batch(B.Q, B.LOG_N + 1, _dev(lwe_q), lwe_2n, 32)
    ph2n = B.check_switched(_host(lwe_2n), s, msgs)
    tv = B.test_vector()
    acc = torch.empty((batch, 2 * B.N), dtype=torch.int32, device="cuda")
    t.blind_rotate_batch(B.LOG_B, None, dkey.view(-1), B.N_LWE, lwe_2n, _dev(tv), acc)
    out = torch.empty((batch, B.N + 1), dtype=torch.int32, device="cuda")
    P.extract_lwe_batch(B.Q, acc, out, B.N, 32)
    accs = _host(acc).reshape(batch, 2, B.N)
    B.check_outputs(_host(out), accs, _times_secret(t, accs[:, 0], z), z, msgs, ph2n)
    # the host-slice shim with a resident key handle returns the same LWE samples bit for bit
    handle = P.BootstrappingKey(t, B.LOG_B, None, B.N_LWE, _host(dkey).reshape(-1))
    assert np.array_equal(handle.bootstrap_slices(_host(lwe_2n).copy(), tv), _host(out))


def test_ternary_bootstrap_recovers_the_lookup_table_at_c5_parameters():
    """Ternary LWE secret: BSK+ / BSK- = RGSW([s_i = +1]) / RGSW([s_i = -1]) with noise, rotation by monomial combination."""
    import torch
    import primus_fhe_b200 as P
    t = P.U32NttTable(B.LOG_N, B.Q)
    basis = P.ApproxSignedBasis(B.Q, B.LOG_B, None, 32)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    rng = np.random.default_rng(77)
    z, s = B.secrets(rng, ternary=True)
    dkeys = []
    for sign in (1, -1):
        a1, e1 = B.rgsw_rows(rng, lv)
        key = B.assemble_key(a1, _times_secret(t, a1, z), e1, (s == sign).astype(np.int64), lv, drop)
        dk = _dev(key.reshape(-1, B.N))
        t.forward_batch(dk)
        dkeys.append(dk.view(-1))
    batch = 96
    msgs, lwe_q = B.lwe_inputs(rng, s, batch)
    lwe_2n = torch.empty((batch, B.N_LWE + 1), dtype=torch.int32, device="cuda")
    P.modulus_switch_batch(B.Q, B.LOG_N + 1, _dev(lwe_q), lwe_2n, 32)
    ph2n = B.check_switched(_host(lwe_2n), s, msgs)
    acc = torch.empty((batch, 2 * B.N), dtype=torch.int32, device="cuda")
    t.blind_rotate_ternary_batch(B.LOG_B, None, dkeys[0], dkeys[1], B.N_LWE, lwe_2n, _dev(B.test_vector()), acc)
    out = torch.empty((batch, B.N + 1), dtype=torch.int32, device="cuda")
    P.extract_lwe_batch(B.Q, acc, out, B.N, 32)
    accs = _host(acc).reshape(batch, 2, B.N)
    B.check_outputs(_host(out), accs, _times_secret(t, accs[:, 0], z), z, msgs, ph2n)
