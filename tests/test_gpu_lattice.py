"""GPU parity for the fused external product and blind rotation vs the CPU oracle.

The reference has no test for the NTT external product and no blind rotation at all (SURVEY 8c:
"parity unpinned" at the composed level); the oracle follows primus_lattice/src/glwe/crt.rs:200-227 /
glwe/dcrt.rs:178-255 line by line and is itself checked against the schoolbook identity in
tests/test_oracle.py.  Shapes: C4-A (u32, N=2048, B=2^7, l=3), C4-B (u64, N=2048, l=7), C5 (u32, N=1024).
"""
import numpy as np
import pytest

from conftest import Q27, Q50

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.from_numpy(x.view(np.int64 if x.dtype == np.uint64 else np.int32)).cuda()


@pytest.mark.parametrize("bits,q,log_n,log_basis,rev,k", [
    (32, Q27, 11, 7, None, 1), (64, Q50, 11, 7, None, 1), (32, Q27, 10, 7, None, 1), (64, Q50, 10, 7, 3, 1),
    (64, Q50, 12, 10, None, 1), (32, Q27, 10, 4, None, 2), (64, Q50, 10, 2, None, 1), (64, Q50, 10, 7, None, 2), (64, Q50, 11, 3, None, 1),
    (64, 1152921504606830593, 10, 7, None, 1), (64, 562949953392641, 11, 7, None, 1), (32, 1073692673, 10, 7, None, 1)])
@pytest.mark.parametrize("to_coeff", [True, False])
def test_external_product_matches_oracle(bits, q, log_n, log_basis, rev, k, to_coeff):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    n = 1 << log_n
    gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    ot = (O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q)
    ob = O.ApproxSignedBasis(q, log_basis, rev, bits)
    levels = ob.decompose_length()
    rng = np.random.default_rng(21)
    batch = 5
    key = rng.integers(0, q, ((k + 1) * levels * (k + 1) * n), dtype=np.uint64).astype(dt)
    cin = rng.integers(0, q, (batch, (k + 1) * n), dtype=np.uint64).astype(dt)
    cin[0, :] = q - 1
    cin[1, :] = 0
    want = O.external_product_single(ot, ob, k, key, cin, to_coeff=to_coeff, batch=batch)
    out = torch.empty((batch, (k + 1) * n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    gt.external_product_batch(k, log_basis, rev, _dev(key), _dev(cin), out, to_coeff=to_coeff)
    assert np.array_equal(out.cpu().numpy().view(dt), want)


@pytest.mark.parametrize("bits,q,log_n,log_basis,n_lwe", [(32, Q27, 10, 7, 12), (64, Q50, 10, 7, 5), (32, Q27, 11, 7, 4), (64, Q50, 10, 2, 3),
                                                          (64, 1152921504606830593, 10, 7, 3), (32, 1073692673, 10, 7, 4), (64, Q50, 11, 7, 3)])
def test_blind_rotate_matches_oracle(bits, q, log_n, log_basis, n_lwe):
    import torch
    import primus_fhe_b200 as P
    from oracle import oracle as O
    dt = np.uint64 if bits == 64 else np.uint32
    n = 1 << log_n
    gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    ot = (O.U64NttTable if bits == 64 else O.U32NttTable)(log_n, q)
    ob = O.ApproxSignedBasis(q, log_basis, None, bits)
    levels = ob.decompose_length()
    rng = np.random.default_rng(33)
    batch = 6
    bsk = rng.integers(0, q, (n_lwe * 2 * levels * 2 * n), dtype=np.uint64).astype(dt)
    lwe = rng.integers(0, 2 * n, (batch, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    lwe[0, :] = 0          # identity rotations
    lwe[1, :] = 2 * n - 1
    lwe[2, :] = n
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    want = O.blind_rotate(ot, ob, bsk, n_lwe, lwe, tv, batch=batch)
    out = torch.empty((batch, 2 * n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    gt.blind_rotate_batch(log_basis, None, _dev(bsk), n_lwe, _dev(lwe), _dev(tv), out)
    assert np.array_equal(out.cpu().numpy().view(dt), want)


def test_external_product_host_slices_match_device():
    """pfhe_ggsw*_external_product_slices: host buffers in, host buffers out (what a host-resident `mul_dcrt_ggsw_to` binds to)."""
    import primus_fhe_b200 as P
    from oracle import oracle as O
    for bits, q in ((32, Q27), (64, Q50)):
        dt = np.uint64 if bits == 64 else np.uint32
        n, k = 1024, 1
        gt = (P.U64NttTable if bits == 64 else P.U32NttTable)(10, q)
        ot = (O.U64NttTable if bits == 64 else O.U32NttTable)(10, q)
        ob = O.ApproxSignedBasis(q, 7, None, bits); levels = ob.decompose_length()
        rng = np.random.default_rng(5)
        key = rng.integers(0, q, 2 * levels * 2 * n, dtype=np.uint64).astype(dt)
        cin = rng.integers(0, q, (7, 2 * n), dtype=np.uint64).astype(dt)
        want = O.external_product_single(ot, ob, k, key, cin, to_coeff=True, batch=7)
        out = np.empty_like(cin)
        gt.external_product_slices(k, 7, None, key, cin, out, True)
        assert np.array_equal(out, want)
