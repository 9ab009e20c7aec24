"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import primus_fhe_b200 as P
rng = np.random.default_rng(1)
def dev(a, bits): return torch.from_numpy(a.astype(np.int64)).to(torch.int64 if bits == 64 else torch.int32).cuda()
for bits, log_n, q in [(64, 12, 1125899906826241), (64, 11, 1125899906826241), (64, 13, 1125899906826241), (64, 10, 1125899906826241),
                       (64, 12, 1152921504606830593), (32, 10, 132120577), (32, 11, 132120577), (32, 13, 132120577), (64, 6, 1125899906826241)]:
    n = 1 << log_n
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(log_n, q)
    x = dev(rng.integers(0, q, (3, n), dtype=np.uint64), bits); x0 = x.clone()
    t.forward_batch(x); t.inverse_batch(x)
    assert torch.equal(x, x0), (bits, log_n)
    y = torch.empty_like(x); t.polymul_batch(x, x0, y)
dc = P.U64DcrtTable(12, [1125899906826241, 1125899906629633])
x = dev(np.stack([rng.integers(0, m, (2, 4096), dtype=np.uint64) for m in dc.moduli], axis=1), 64).contiguous(); x0 = x.clone()
dc.forward_batch(x); dc.inverse_batch(x); assert torch.equal(x, x0)
for bits, q in ((32, 132120577), (64, 1125899906826241)):
    t = (P.U64NttTable if bits == 64 else P.U32NttTable)(10, q)
    lv = P.ApproxSignedBasis(q, 7, None, bits).decompose_length()
    key = dev(rng.integers(0, q, 2 * lv * 2 * 1024, dtype=np.uint64), bits)
    cin = dev(rng.integers(0, q, (3, 2048), dtype=np.uint64), bits); out = torch.empty_like(cin)
    t.external_product_batch(1, 7, None, key, cin, out, True)
    t.external_product_batch(1, 7, None, key, cin, out, False)
    nl = 3
    bsk = dev(rng.integers(0, q, nl * 2 * lv * 2 * 1024, dtype=np.uint64), bits)
    lwe = torch.from_numpy(rng.integers(0, 2048, (2, nl + 1), dtype=np.uint64).astype(np.int32)).cuda()
    tv = dev(rng.integers(0, q, 1024, dtype=np.uint64), bits); acc = torch.empty((2, 2048), dtype=cin.dtype, device="cuda")
    t.blind_rotate_batch(7, None, bsk, nl, lwe, tv, acc)
mods = [1125899906826241, 1125899906629633]
bb = P.BigUintApproxSignedBasis(P.RNSBase(mods, 64), 9, 4)
dc2 = P.U64DcrtTable(10, mods)
key = dev(np.stack([rng.integers(0, m, (2 * 4 * 2, 1024), dtype=np.uint64) for m in mods], axis=1), 64).contiguous()
cin = dev(np.stack([rng.integers(0, m, (2 * 2, 1024), dtype=np.uint64) for m in mods], axis=1), 64).contiguous()
out = torch.empty_like(cin)
P.dcrt_external_product_batch(dc2, bb, 1, key, cin, out, True)
# round 2: N = 16384 fused product (stash variant), bootstrap shim (lattice32.cu + extract), multi-device driver, modulus switch, UintNttTable
t14 = P.U64NttTable(14, 1125899904679937)
x = dev(rng.integers(0, 1125899904679937, (2, 16384), dtype=np.uint64), 64); y = torch.empty_like(x)
t14.polymul_batch(x, x.flip(0).contiguous(), y)
t10 = P.U32NttTable(10, 132120577)
lv = P.ApproxSignedBasis(132120577, 7, None, 32).decompose_length()
bskh = rng.integers(0, 132120577, 3 * 2 * lv * 2 * 1024, dtype=np.uint64).astype(np.uint32)
lweh = rng.integers(0, 2048, (3, 4), dtype=np.uint64).astype(np.uint32)
tvh = rng.integers(0, 132120577, 1024, dtype=np.uint64).astype(np.uint32)
key = P.BootstrappingKey(t10, 7, None, 3, bskh)
key.bootstrap_slices(lweh, tvh)
mt = P.MultiNttTable(10, 132120577, [0, 0], 32)
mt.bootstrap_slices(mt.bootstrapping_keys(7, None, 3, bskh), lweh, tvh)
h = rng.integers(0, 132120577, (5, 1024), dtype=np.uint64).astype(np.uint32); mt.transform_slices(h)
v = dev(rng.integers(0, 132120577, 4096, dtype=np.uint64), 32); o = torch.empty(4096, dtype=torch.int32, device="cuda")
P.modulus_switch_batch(132120577, 11, v, o, 32)
u = P.UintNttTable(10, 12289, 16); a = rng.integers(0, 12289, (2, 1024), dtype=np.uint64).astype(np.uint16); u.transform_slices(a); u.inverse_transform_slices(a)
torch.cuda.synchronize()
print("sanitize workload ok")
