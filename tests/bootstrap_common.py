"""Shared pieces of the functional bootstrapping tests (CPU oracle and GPU): real LWE / RGSW encryptions with noise at the BASELINE
config-5 parameters, and the decryption checks.  Conventions as fixed in include/pfhe.h and the oracle (the reference has no bootstrapping):
  LWE phase = b - <a, s>;  RLWE phase = b - a*z;  BSK_i[r][l] = RLWE_z(0) + s_i * g_l in component r,  g_l = 2^(drop_bits + l * log_basis);
  blind rotation leaves tv * X^(-(b - sum a_i s_i)), so coefficient j of the body is (-1)^floor((j+p)/N) * tv[(j + p) mod N], p = phase mod 2N."""
import numpy as np

Q, LOG_N, N_LWE, LOG_B = 132120577, 10, 512, 7
N = 1 << LOG_N
LUT = np.array([2, 0, 3, 1])          # not the identity: the bootstrap is programmable


def secrets(rng, ternary=False):
    """(RLWE secret z, binary; LWE secret s, binary or ternary)"""
    return rng.integers(0, 2, N).astype(np.int64), (rng.integers(-1, 2, N_LWE) if ternary else rng.integers(0, 2, N_LWE)).astype(np.int64)


def rgsw_rows(rng, levels):
    """Fresh uniform `a` parts and noise in [-2, 2] for the n_lwe * 2 * levels RLWE(0) rows of the key."""
    rows = N_LWE * 2 * levels
    return rng.integers(0, Q, (rows, N), dtype=np.uint64).astype(np.uint32), rng.integers(-2, 3, (rows, N))


def assemble_key(a1, az, e1, s, levels, drop):
    """key[i][r][l][c][N] in coefficient form from a (uniform), a*z, the noise, the LWE secret and the gadget."""
    b1 = ((az.astype(np.int64) + e1) % Q).astype(np.uint32)
    key = np.empty((N_LWE, 2, levels, 2, N), dtype=np.uint32)
    key[:, :, :, 0, :] = a1.reshape(N_LWE, 2, levels, N)
    key[:, :, :, 1, :] = b1.reshape(N_LWE, 2, levels, N)
    for l in range(levels):
        g = (1 << (drop + l * LOG_B)) % Q
        for r in (0, 1):
            key[:, r, l, r, 0] = ((key[:, r, l, r, 0].astype(np.int64) + s * g) % Q).astype(np.uint32)
    return key


def lwe_inputs(rng, s, batch):
    """Messages m in {0..3} with one padding bit: phase = (2m + 1) * q/16 + noise."""
    msgs = rng.integers(0, 4, batch)
    msgs[:min(4, batch)] = [0, 1, 2, 3][:min(4, batch)]
    a = rng.integers(0, Q, (batch, N_LWE), dtype=np.uint64).astype(np.int64)
    mu = np.round(Q * (msgs * 256 + 128) / 2048).astype(np.int64)
    b = (a @ s + mu + rng.integers(-8, 9, batch)) % Q
    return msgs, np.concatenate([a, b[:, None]], axis=1).astype(np.uint32)


def test_vector():
    return (LUT[np.arange(N) // 256] * (Q // 8)).astype(np.uint32)


def check_switched(lwe_2n, s, msgs):
    l = lwe_2n.astype(np.int64)
    ph = (l[:, N_LWE] - l[:, :N_LWE] @ s) % (2 * N)
    assert np.all(ph // 256 == msgs) and np.all(np.abs(ph % 256 - 128) < 64)
    return ph


def check_outputs(lwe_out, acc, az_of_acc, z, msgs, ph2n):
    """lwe_out [batch][N+1] (extracted), acc [batch][2][N], az_of_acc = acc[:,0] * z (ring product)."""
    o = lwe_out.astype(np.int64)
    phase = (o[:, N] - o[:, :N] @ z) % Q
    want = LUT[msgs] * (Q // 8)
    err = (phase - want + Q // 2) % Q - Q // 2
    assert np.all(np.abs(err) < Q // 64), int(np.abs(err).max())       # noise far below the decoding margin q/16
    assert np.array_equal(np.round(phase * 8 / Q).astype(np.int64) % 8, LUT[msgs])
    tv = test_vector().astype(np.int64)
    body = (acc[:, 1].astype(np.int64) - az_of_acc.astype(np.int64)) % Q
    for i in range(min(4, len(msgs))):                                  # the accumulator encrypts the whole rotated test vector
        for j in (0, 5, N - 1):
            src = int(ph2n[i]) + j
            expect = int(tv[src % N]) * (1 if (src // N) % 2 == 0 else -1)
            d = (int(body[i, j]) - expect + Q // 2) % Q - Q // 2
            assert abs(d) < Q // 64
    return int(np.abs(err).max())
