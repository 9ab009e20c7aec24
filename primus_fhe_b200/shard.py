"""Multi-GPU plumbing for the hot path: independent work units (polynomials, RNS limbs x polynomials,
ciphertexts) are split contiguously across ranks; tables / keys are replicated; there is NO collective on
the data path.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only to gather
batched results when the caller wants them on every rank."""
from __future__ import annotations


def shard_range(total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `total` units owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world_size: int) -> list[int]:
    return [shard_range(total, world_size, r)[1] - shard_range(total, world_size, r)[0] for r in range(world_size)]


def gather_batches(local, total_units: int, group=None):
    """All-gather per-rank result batches ([units_r, ...] tensors, ragged in dim 0) into one [total_units, ...]
    tensor on every rank.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(total_units, world)
    assert local.shape[0] == sizes[dist.get_rank(group)], "local batch does not match this rank's shard"
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)
