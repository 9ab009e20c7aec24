"""Streaming kernels (slice ops, multi-word gadget, compose, base conversion) at >= 256 MiB per operand -- used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
Q, QB = 1125899906826241, 1125899906629633
g = torch.Generator(device="cuda"); g.manual_seed(3)
n = 1 << 25
x = torch.randint(0, Q, (n,), dtype=torch.int64, device="cuda", generator=g); y = torch.randint(0, Q, (n,), dtype=torch.int64, device="cuda", generator=g)
o = torch.empty_like(x)
bm = P.BarrettModulus(Q, 64)
for _ in range(5):
    bm.reduce_mul_slice_to(x, y, o)
m2 = [Q, QB]
rns = P.RNSBase(m2, 64); bb = P.BigUintApproxSignedBasis(rns, 7, None)
polys = 2048
res = torch.stack([torch.randint(0, m, (polys, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
digs = torch.empty((polys, bb.decompose_length(), 2, 2048), dtype=torch.int64, device="cuda")
for _ in range(3):
    bb.gadget_decompose_batch(res, digs, 2048)
big = torch.empty((polys * 2048, rns.big_uint_value_len()), dtype=torch.int64, device="cuda")
flat = res.permute(1, 0, 2).contiguous().view(2, -1)
for _ in range(3):
    rns.compose_multiple_values_to(flat, big)
bc = P.BaseConverter([137438822401, 137438814209, 137438773249], m2, 64)
cin = torch.stack([torch.randint(0, m, (polys, 2048), dtype=torch.int64, device="cuda", generator=g) for m in bc.in_moduli], dim=1).contiguous()
cout = torch.empty((polys, 2, 2048), dtype=torch.int64, device="cuda")
for _ in range(3):
    bc.fast_convert_array(cin, cout, 2048)
torch.cuda.synchronize()
print("stream workload ok")
