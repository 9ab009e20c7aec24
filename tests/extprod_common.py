"""Shared pieces of the functional external-product tests (CPU oracle and GPU): RLWE(m) [x] RGSW(mu) must decrypt to m * mu.
Real encryptions with noise at the BASELINE config-4 shapes; conventions as in bootstrap_common.py:
  RLWE phase = b - a*z;  RGSW row (r, l) = RLWE_z(0) + mu * g_l in component r,  g_l = 2^(drop_bits + l * log_basis) (mod every limb)."""
import numpy as np

MU_TERMS = ((0, 1), (5, 1), (100, -1))      # mu = 1 + X^5 - X^100
T = 16                                      # message space; Delta = floor(Q / 16)


def negacyclic_by_mu(msg):
    """msg * mu in Z[X]/(X^N + 1), msg integer array [..., N]."""
    n = msg.shape[-1]
    out = np.zeros_like(msg)
    for k, sgn in MU_TERMS:
        rolled = np.roll(msg, k, axis=-1)
        rolled[..., :k] *= -1
        out += sgn * rolled
    return out


def add_mu_times(poly_row, g, q):
    """poly_row (uint array [..., N]) += mu * g  (mod q), in place on a copy."""
    r = poly_row.astype(object)
    for k, sgn in MU_TERMS:
        r[..., k] = (r[..., k] + sgn * g) % q
    return (r % q).astype(poly_row.dtype)


def rgsw_and_inputs(rng, moduli, n, levels, drop, log_b, batch, ring_mul, dt):
    """Returns (key [2][levels][2][L][N] coefficient form, glwe_in [batch][2][L][N], z, msg [batch][N]).
    ring_mul(limb, rows[rows, N]) -> rows * z mod q_limb."""
    L = len(moduli)
    Q = 1
    for m in moduli:
        Q *= m
    z = rng.integers(0, 2, n).astype(np.int64)
    key = np.empty((2, levels, 2, L, n), dtype=dt)
    e = rng.integers(-2, 3, (2, levels, n))
    for i, q in enumerate(moduli):
        a = rng.integers(0, q, (2 * levels, n), dtype=np.uint64).astype(dt)
        az = ring_mul(i, a, z).astype(np.int64)
        b = ((az + e.reshape(2 * levels, n)) % q).astype(dt)
        key[:, :, 0, i, :] = a.reshape(2, levels, n)
        key[:, :, 1, i, :] = b.reshape(2, levels, n)
        for l in range(levels):
            g = (1 << (drop + l * log_b)) % q
            for r in (0, 1):
                key[r, l, r, i, :] = add_mu_times(key[r, l, r, i, :], g, q)
    delta = Q // T
    msg = rng.integers(0, T, (batch, n)).astype(np.int64)
    e_in = rng.integers(-4, 5, (batch, n))
    glwe = np.empty((batch, 2, L, n), dtype=dt)
    for i, q in enumerate(moduli):
        a = rng.integers(0, q, (batch, n), dtype=np.uint64).astype(dt)
        az = ring_mul(i, a, z).astype(np.int64)
        scaled = np.array([[(delta * int(v)) % q for v in row] for row in msg], dtype=np.int64)
        glwe[:, 0, i, :] = a
        glwe[:, 1, i, :] = ((az + scaled + e_in) % q).astype(dt)
    return key, glwe, z, msg


def check(out, moduli, z, msg, ring_mul):
    """out [batch][2][L][N] coefficient form: phase (CRT-composed, centred) must be Delta * (msg * mu) + small noise."""
    L = len(moduli)
    Q = 1
    for m in moduli:
        Q *= m
    delta = Q // T
    batch, n = msg.shape
    want = negacyclic_by_mu(msg)
    res = []
    for i, q in enumerate(moduli):
        az = ring_mul(i, np.ascontiguousarray(out[:, 0, i, :]), z).astype(np.int64)
        res.append((out[:, 1, i, :].astype(np.int64) - az) % q)
    worst = 0
    for bi in range(batch):
        for j in range(n):
            x = 0
            for i, q in enumerate(moduli):     # CRT
                Mi = Q // q
                x += int(res[i][bi, j]) * Mi * pow(Mi, -1, q)
            x %= Q
            d = (x - delta * int(want[bi, j])) % Q
            d = d - Q if d > Q // 2 else d
            worst = max(worst, abs(d))
    assert 0 < worst < delta // 64, (worst, delta)
    return worst
