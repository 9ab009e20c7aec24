"""Experiment: do the integer-pipe and FP64-pipe NTT kernels overlap when run concurrently?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import primus_fhe_b200 as P
q, log_n, n = 1125899906826241, 12, 4096
tf = P.U64NttTable(log_n, q)
os.environ["PFHE_DISABLE_F64"] = "1"
ti = P.U64NttTable(log_n, q)
os.environ["PFHE_DISABLE_F64"] = "0"
batch = 65536
x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(frac_int):
    bi = int(batch * frac_int)
    xi, xf = x[:bi], x[bi:]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        if bi < batch: tf.forward_batch(xf)
    with torch.cuda.stream(s2):
        if bi > 0: ti.forward_batch(xi)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for f in (0.0, 0.125, 0.25, 0.33, 0.4, 0.5, 1.0):
    run(f); ms = min(run(f) for _ in range(4))
    print(f"int fraction {f:.3f}: {batch/ms*1e3:.3e} NTT/s")
