"""Functional check of the GGSW external product on the GPU at the BASELINE config-4 degree (N = 2048, base 2^7): RLWE(m) [x] RGSW(mu)
with real noisy encryptions must decrypt to m * mu -- the re-scheduled u32 kernel, the FP64 u64 kernel, and the single-kernel two-limb
product.  The same flow runs against the oracle in tests/test_oracle.py (CPU)."""
import numpy as np
import pytest

import extprod_common as X
from conftest import Q27, Q50, Q50B

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,moduli,log_n", [(32, [Q27], 11), (64, [Q50], 11), (64, [Q50, Q50B], 11), (32, [Q27], 10)])
def test_external_product_decrypts_to_the_product(bits, moduli, log_n):
    import torch
    import primus_fhe_b200 as P
    dt = np.uint64 if bits == 64 else np.uint32
    tdt = torch.int64 if bits == 64 else torch.int32
    n, L = 1 << log_n, len(moduli)
    tables = [(P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q) for q in moduli]

    def dev(x):
        return torch.from_numpy(np.ascontiguousarray(x).view(np.int64 if bits == 64 else np.int32)).cuda()

    def host(t):
        return t.cpu().numpy().view(dt)

    def ring_mul(i, rows, z):
        da = dev(rows.astype(dt))
        dz = dev(np.broadcast_to(z.astype(dt), rows.shape))
        dc = torch.empty_like(da)
        tables[i].polymul_batch(da, dz, dc)
        return host(dc)

    rng = np.random.default_rng(5)
    if L == 1:
        basis = P.ApproxSignedBasis(moduli[0], 7, None, bits)
    else:
        basis = P.BigUintApproxSignedBasis(P.RNSBase(moduli, bits), 7, None)
    lv, drop = basis.decompose_length(), basis.drop_bits()
    batch = 4
    key, glwe, z, msg = X.rgsw_and_inputs(rng, moduli, n, lv, drop, 7, batch, ring_mul, dt)
    for i in range(L):
        rows = dev(key[:, :, :, i, :].reshape(-1, n))
        tables[i].forward_batch(rows)
        key[:, :, :, i, :] = host(rows).reshape(2, lv, 2, n)
    dkey, din = dev(key.reshape(-1)), dev(glwe.reshape(batch, -1))
    out = torch.empty_like(din)
    if L == 1:
        tables[0].external_product_batch(1, 7, None, dkey, din, out, True)
    else:
        P.dcrt_external_product_batch((P.U64DcrtTable if bits == 64 else P.U32DcrtTable)(log_n, moduli), basis, 1, dkey, din, out, True)
    X.check(host(out).reshape(batch, 2, L, n), moduli, z, msg, ring_mul)
