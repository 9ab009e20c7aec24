#!/bin/bash
mkdir -p gpurun_out
{ echo "== parity of the st.async forward against the default (digest over 3 polys + 8-limb batch)"
  timeout 120 python - <<'PY'
import os, subprocess, sys, json
code = r"""
import sys, hashlib, numpy as np, torch
sys.path.insert(0, '.')
import primus_fhe_b200 as P
from bench import _c3_primes
rng = np.random.default_rng(1)
q = 1125899904679937
t = P.U64NttTable(14, q)
x = torch.from_numpy(rng.integers(0, q, (301, 16384), dtype=np.uint64).astype(np.int64)).cuda()
x[0] = q - 1; x[1] = 0
f = x.clone(); t.forward_batch(f)
mods = _c3_primes(); dc = P.U64DcrtTable(14, mods)
y = torch.stack([torch.from_numpy(rng.integers(0, m, (5, 16384), dtype=np.uint64).astype(np.int64)) for m in mods], dim=1).contiguous().cuda()
g = y.clone(); dc.forward_batch(g)
print(hashlib.sha256(f.cpu().numpy().tobytes() + g.cpu().numpy().tobytes()).hexdigest())
"""
outs = []
for env in ({"PFHE_NTT_CLUSTER": "0"}, {}, {"PFHE_NTT_CLUSTER_ASYNC": "1"}):
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=100)
    outs.append(p.stdout.strip() or p.stderr[-500:])
    print(env, outs[-1])
print("IDENTICAL" if len(set(outs)) == 1 else "MISMATCH")
PY
  for r in 1 2 3; do
    echo "== round $r cluster (barrier.cluster)"; timeout 120 python tools/gpu_c3.py 2>&1 | sed -n 7,8p
    echo "== round $r cluster (st.async + mbarrier)"; PFHE_NTT_CLUSTER_ASYNC=1 timeout 120 python tools/gpu_c3.py 2>&1 | sed -n 7,8p
  done; } > gpurun_out/r2ax.log 2>&1
cat gpurun_out/r2ax.log
