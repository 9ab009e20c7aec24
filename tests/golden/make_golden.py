"""Generates tests/golden/golden_small.json from the CPU oracle, cross-checked against the independent
big-int model (oracle/pymodel.py) before writing.  The reference ships no golden vectors for this path
 and cannot be run here (Rust, no toolchain), so these are oracle-minted known answers;
they pin the oracle AND the CUDA path against silent drift.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import pymodel as M  # noqa: E402

rng = np.random.default_rng(0x5EED0000)
out = {"ntt": [], "decompose": [], "external_product": [], "seed": "0x5EED0000"}
for bits, q, log_n in [(32, 132120577, 4), (32, 132120577, 6), (64, 1125899906826241, 5), (64, 1152921504606830593, 4), (64, 562949953392641, 6)]:
    cls = O.U64NttTable if bits == 64 else O.U32NttTable
    dt = np.uint64 if bits == 64 else np.uint32
    t = cls(log_n, q); n = 1 << log_n
    x = rng.integers(0, q, n, dtype=np.uint64).astype(dt)
    y = x.copy(); t.transform_slice(y)
    assert [int(v) for v in y] == M.ntt_forward(x, q, t.root())
    out["ntt"].append({"bits": bits, "q": q, "log_n": log_n, "root": t.root(), "input": [int(v) for v in x], "forward": [int(v) for v in y]})
for bits, q, beta, lv in [(32, 132120577, 7, None), (64, 1125899906826241, 7, None), (32, 132120577, 7, 2), (64, 1125899906826241, 10, 3)]:
    dt = np.uint64 if bits == 64 else np.uint32
    b = O.ApproxSignedBasis(q, beta, lv, bits); g = M.Gadget(q, beta, lv)
    v = rng.integers(0, q, 24, dtype=np.uint64).astype(dt); v[:3] = [0, q - 1, q // 2]
    d = b.decompose_slice(v)
    for i in range(len(v)):
        assert [int(d[l, i]) for l in range(g.levels)] == [s % q for s in g.signed_digits(int(v[i]))]
    out["decompose"].append({"bits": bits, "q": q, "log_basis": beta, "levels": lv, "values": [int(x) for x in v],
                             "digits": [[int(x) for x in row] for row in d]})
for bits, q, log_n, beta in [(32, 132120577, 4, 7), (64, 1125899906826241, 4, 7)]:
    cls = O.U64NttTable if bits == 64 else O.U32NttTable
    dt = np.uint64 if bits == 64 else np.uint32
    t = cls(log_n, q); n = 1 << log_n
    sb = O.ApproxSignedBasis(q, beta, None, bits); lv = sb.decompose_length()
    key = rng.integers(0, q, 2 * lv * 2 * n, dtype=np.uint64).astype(dt)
    cin = rng.integers(0, q, 2 * n, dtype=np.uint64).astype(dt)
    res = O.external_product_single(t, sb, 1, key, cin.reshape(1, -1))
    out["external_product"].append({"bits": bits, "q": q, "log_n": log_n, "log_basis": beta, "key": [int(v) for v in key],
                                    "input": [int(v) for v in cin], "output": [int(v) for v in res.reshape(-1)]})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.json"), "w"))
print("written")
