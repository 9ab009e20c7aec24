"""Every tuning hook selects a different kernel for the same mathematical function: all of them must give the same bits.
(PFHE_NTT_TMA=0 -> LSU copy-out/copy-in kernels; PFHE_F64_LAZY=0 -> per-stage-fold FP64 butterflies; PFHE_DISABLE_F64=1 ->
integer-pipe butterflies; PFHE_DISABLE_WIDE32=1 -> Harvey forward butterflies in the lattice kernels.)  The hooks are read once
per process, so each setting runs in a subprocess that prints digests of its outputs; the default setting is additionally
checked against the CPU oracle in the other test modules."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import primus_fhe_b200 as P
out = {}
def dig(t): return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()
rng = np.random.default_rng(123)
for bits, log_n, q, batch in [(64, 12, 1125899906826241, 5), (64, 11, 1125899906826241, 7), (64, 13, 1125899906826241, 3),
                              (64, 10, 1125899906826241, 9), (32, 10, 132120577, 9), (32, 12, 268369921, 3), (32, 13, 132120577, 2),
                              (64, 14, 1125899904679937, 3)]:
    n = 1 << log_n
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    x = rng.integers(0, q, (batch, n), dtype=np.uint64)
    x[0, :] = q - 1; x[-1, :] = 0
    tdt = torch.int64 if bits == 64 else torch.int32
    d = torch.from_numpy(x.astype(np.int64)).to(tdt).cuda()
    e = d.flip(0).contiguous()
    key = f"u{bits}_n{n}"
    f = d.clone(); t.forward_batch(f); out[key + "_fwd"] = dig(f)
    g = torch.empty_like(d); t.forward_batch_to(d, g); assert torch.equal(f, g)
    i = d.clone(); t.inverse_batch(i); out[key + "_inv"] = dig(i)
    c = torch.empty_like(d); t.polymul_batch(d, e, c); out[key + "_mul"] = dig(c)
    if log_n == 14:   # in-place forms (the output parks fwd(a): aliasing rules of the stash / cluster kernels)
        c1 = d.clone(); t.polymul_batch(c1, e, c1); assert torch.equal(c1, c)
        c2 = e.clone(); t.polymul_batch(d, c2, c2); assert torch.equal(c2, c)
        sq = torch.empty_like(d); t.polymul_batch(d, d, sq); c3 = d.clone(); t.polymul_batch(c3, c3, c3); assert torch.equal(c3, sq)
        out[key + "_sq"] = dig(sq)
# DCRT (per-limb tables) through the same kernels
mods = [1125899906826241, 1125899906629633, 1125899904679937]
dc = P.U64DcrtTable(12, mods)
x = np.stack([rng.integers(0, m, (3, 4096), dtype=np.uint64) for m in mods], axis=1)
d = torch.from_numpy(x.astype(np.int64)).cuda().contiguous()
f = d.clone(); dc.forward_batch(f); out["dcrt_fwd"] = dig(f)
i = d.clone(); dc.inverse_batch(i); out["dcrt_inv"] = dig(i)
c = torch.empty_like(d); dc.polymul_batch(d, d.flip(0).contiguous(), c); out["dcrt_mul"] = dig(c)
# lattice kernels (u32 wide forward / Harvey forward, u64 FP64 / integer)
for bits, q in ((32, 132120577), (64, 1125899906826241)):
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(10, q)
    lv = P.ApproxSignedBasis(q, 7, None, bits).decompose_length()
    tdt = torch.int64 if bits == 64 else torch.int32
    key_ = torch.from_numpy(rng.integers(0, q, 2 * lv * 2 * 1024, dtype=np.uint64).astype(np.int64)).to(tdt).cuda()
    cin = torch.from_numpy(rng.integers(0, q, (4, 2048), dtype=np.uint64).astype(np.int64)).to(tdt).cuda()
    o = torch.empty_like(cin); t.external_product_batch(1, 7, None, key_, cin, o, True); out[f"ep{bits}"] = dig(o)
    o2 = torch.empty_like(cin); t.external_product_batch(1, 7, None, key_, cin, o2, False); out[f"ep{bits}_ntt"] = dig(o2)
# u32 N = 2048 external product: re-scheduled kernel (lattice32_ep.cu) vs the generic one (PFHE_EP_FAST=0)
t = P.U32NttTable(11, 132120577)
for lb, lvl in ((7, None), (4, 5)):
    lv = P.ApproxSignedBasis(132120577, lb, lvl, 32).decompose_length()
    key_ = torch.from_numpy(rng.integers(0, 132120577, 2 * lv * 2 * 2048, dtype=np.uint64).astype(np.int64)).to(torch.int32).cuda()
    cin = torch.from_numpy(rng.integers(0, 132120577, (5, 4096), dtype=np.uint64).astype(np.int64)).to(torch.int32).cuda()
    for tc in (True, False):
        o = torch.empty_like(cin); t.external_product_batch(1, lb, lvl, key_, cin, o, tc); out[f"ep32_n2048_b{lb}_{int(tc)}"] = dig(o)
# multi-limb external product: single fused kernel vs gadget kernel + per-limb kernel (PFHE_DCRT_EP_TWO_KERNEL=1)
for bits, m2, lb, lvl in ((64, [1125899906826241, 1125899906629633], 7, None), (32, [134215681, 134176769], 7, None), (64, [1125899906826241, 1152921504606830593], 9, 5),
                          (64, [1125899906826241, 1125899906629633, 562949953392641], 7, None), (32, [134215681, 134176769, 132120577, 268369921], 7, 11),
                          (64, [1125899906826241, 1125899906629633, 562949953392641, 1152921504606830593], 11, None)):
    tdt = torch.int64 if bits == 64 else torch.int32
    dcx = (P.U64DcrtTable if bits == 64 else P.U32DcrtTable)(10, m2)
    bbx = P.BigUintApproxSignedBasis(P.RNSBase(m2, bits), lb, lvl)
    lvx = bbx.decompose_length()
    keyx = torch.stack([torch.from_numpy(rng.integers(0, m, (2 * lvx * 2, 1024), dtype=np.uint64).astype(np.int64)).to(tdt) for m in m2], dim=1).contiguous().cuda()
    cinx = torch.stack([torch.from_numpy(rng.integers(0, m, (3 * 2, 1024), dtype=np.uint64).astype(np.int64)).to(tdt) for m in m2], dim=1).contiguous().cuda()
    for tc in (True, False):
        ox = torch.empty_like(cinx); P.dcrt_external_product_batch(dcx, bbx, 1, keyx, cinx, ox, tc); out[f"dcrt_ep{bits}_L{len(m2)}_{lb}_{int(tc)}"] = dig(ox)
# blind rotation: re-scheduled u32 kernel (lattice32.cu) vs the generic lattice kernel (PFHE_BR_FAST=0)
t = P.U32NttTable(10, 132120577)
for lb, lvl, nl in ((7, None, 24), (4, 5, 8), (1, 6, 8)):
    lv = P.ApproxSignedBasis(132120577, lb, lvl, 32).decompose_length()
    bsk = torch.from_numpy(rng.integers(0, 132120577, nl * 2 * lv * 2 * 1024, dtype=np.uint64).astype(np.int64)).to(torch.int32).cuda()
    lwe = torch.from_numpy(rng.integers(0, 2048, (5, nl + 1), dtype=np.uint64).astype(np.int64)).to(torch.int32).cuda()
    tv = torch.from_numpy(rng.integers(0, 132120577, 1024, dtype=np.uint64).astype(np.int64)).to(torch.int32).cuda()
    acc = torch.empty((5, 2048), dtype=torch.int32, device="cuda")
    t.blind_rotate_batch(lb, lvl, bsk, nl, lwe, tv, acc); out[f"br32_b{lb}"] = dig(acc)
# pointwise product and base conversion: FP64-pipe formulations (u64 moduli below 2^50) vs the integer ones (PFHE_DISABLE_F64=1)
Q50_, Q50B_, Q49_ = 1125899906826241, 1125899906629633, 562949953392641
bm = P.BarrettModulus(Q50_, 64)
xa = torch.from_numpy(rng.integers(0, Q50_, 4096, dtype=np.uint64).astype(np.int64)).cuda()
xb = torch.from_numpy(rng.integers(0, Q50_, 4096, dtype=np.uint64).astype(np.int64)).cuda()
xa[0] = Q50_ - 1; xb[0] = Q50_ - 1; xa[7] = Q50_ + 3      # one non-canonical word
xo = torch.empty_like(xa); bm.reduce_mul_slice_to(xa, xb, xo); out["reduce_mul_q50"] = dig(xo)
bcv = P.BaseConverter([Q50_, Q50B_, Q49_], [1125899904679937, 1125899905744897], 64)
ci = torch.stack([torch.from_numpy(rng.integers(0, m, (3, 256), dtype=np.uint64).astype(np.int64)) for m in bcv.in_moduli], dim=1).contiguous().cuda()
ci[0, 0, 0] = (1 << 63) - 1                                  # arbitrary word: reduced before the product
co = torch.empty((3, 2, 256), dtype=torch.int64, device="cuda"); bcv.fast_convert_array(ci, co, 256); out["baseconv_fast"] = dig(co)
bc1 = P.BaseConverter([Q50_, Q50B_, Q49_], [1125899904679937], 64)
ce = torch.empty((3, 256), dtype=torch.int64, device="cuda"); bc1.exact_convert_array(ci, ce, 256); out["baseconv_exact"] = dig(ce)
print("DIGESTS " + json.dumps(out))
''' % ROOT


def _run(env):
    p = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("DIGESTS ")]
    assert lines, p.stderr[-3000:]
    return json.loads(lines[0][8:])


def test_all_kernel_variants_agree_bit_for_bit():
    base = _run({})
    for env in ({"PFHE_NTT_TMA": "0"}, {"PFHE_F64_LAZY": "0"}, {"PFHE_DISABLE_F64": "1"}, {"PFHE_DISABLE_WIDE32": "1"},
                {"PFHE_NTT_TMA": "0", "PFHE_DISABLE_F64": "1"}, {"PFHE_BR_FAST": "0"}, {"PFHE_BR_MINB": "5"}, {"PFHE_EP_FAST": "0"},
                {"PFHE_POLYMUL_STASH": "0"}, {"PFHE_STAGE": "0"}, {"PFHE_DCRT_EP_TWO_KERNEL": "1"}, {"PFHE_EP_KEY_PREFETCH": "1"}, {"PFHE_EP_KEY_PREFETCH": "2"},
                {"PFHE_DCRT_EP_FUSED_WIDE": "1"}, {"PFHE_NTT_CLUSTER": "0"}, {"PFHE_NTT_CLUSTER": "1"}, {"PFHE_NTT_CLUSTER_STAGGER_NS": "0"}, {"PFHE_NTT_CLUSTER_ASYNC": "0"}, {"PFHE_NTT_CLUSTER13": "2"}):
        other = _run(env)
        diff = [k for k in base if base[k] != other[k]]
        assert not diff, (env, diff)


def test_empty_batches_are_noops():
    import torch
    import primus_fhe_b200 as P
    t = P.U64NttTable(12, 1125899906826241)
    e = torch.empty((0, 4096), dtype=torch.int64, device="cuda")
    t.forward_batch(e); t.inverse_batch(e); t.polymul_batch(e, e, e)
    dc = P.U64DcrtTable(10, [1125899906826241, 1125899906629633])
    e2 = torch.empty((0, 2, 1024), dtype=torch.int64, device="cuda")
    dc.forward_batch(e2); dc.inverse_batch(e2)
    torch.cuda.synchronize()
