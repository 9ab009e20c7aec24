import sys; sys.path.insert(0, "/root/repo")
import torch, primus_fhe_b200 as P
Q = 1125899906826241
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
m2 = [Q, 1125899906629633]
dc2 = P.U64DcrtTable(11, m2)
bb = P.BigUintApproxSignedBasis(P.RNSBase(m2, 64), 7, None)
lv2 = bb.decompose_length()
g = torch.Generator(device="cuda"); g.manual_seed(1)
key2 = torch.stack([torch.randint(0, m, (2 * lv2 * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
cin2 = torch.stack([torch.randint(0, m, (1024 * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
cout2 = torch.empty_like(cin2)
scratch = P.dcrt_external_product_batch(dc2, bb, 1, key2, cin2, cout2, True)
ms = timeit(lambda: P.dcrt_external_product_batch(dc2, bb, 1, key2, cin2, cout2, True, scratch=scratch))
print(f"dcrt ext product N=2048 L=2 l={lv2} batch 1024: {1024/ms*1e3:.3e} /s ({ms:.3f} ms)")
# three and four limbs (composed value of 3 / 4 words): single fused kernel vs the gadget kernel + per-limb kernel pair (PFHE_DCRT_EP_TWO_KERNEL=1)
for mods in ([Q, 1125899906629633, 562949953392641], [Q, 1125899906629633, 562949953392641, 1125899905744897]):
    L = len(mods)
    dc = P.U64DcrtTable(11, mods)
    bbL = P.BigUintApproxSignedBasis(P.RNSBase(mods, 64), 7, None)
    lv = bbL.decompose_length()
    nb = 512
    keyL = torch.stack([torch.randint(0, m, (2 * lv * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in mods], dim=1).contiguous()
    cinL = torch.stack([torch.randint(0, m, (nb * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in mods], dim=1).contiguous()
    coutL = torch.empty_like(cinL)
    sc = P.dcrt_external_product_batch(dc, bbL, 1, keyL, cinL, coutL, True)
    ms = timeit(lambda: P.dcrt_external_product_batch(dc, bbL, 1, keyL, cinL, coutL, True, scratch=sc))
    print(f"dcrt ext product N=2048 L={L} l={lv} batch {nb}: {nb/ms*1e3:.3e} /s ({ms:.3f} ms)")
