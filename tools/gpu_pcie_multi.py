"""Concurrent PCIe probe: every rank copies 1 GiB pinned H2D and D2H at the same time on its own GPU, all ranks together
(torchrun, gloo barrier).  Rank 0 prints one JSON line: the slowest rank's per-direction GB/s = what the e2e host-slice path can get
when N ranks stream simultaneously.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_pcie_multi.py
"""
import json, os, time
import torch
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
both(); torch.cuda.synchronize()
best = 1e9
for _ in range(4):
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); both(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
t = torch.tensor([best], dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"ranks": world, "per_direction_gbs_slowest_rank": n / float(t.item()) / 1e9}))
if world > 1:
    dist.destroy_process_group()
