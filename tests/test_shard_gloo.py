"""Host-side multi-rank logic on CPU: world_size-2 gloo process group (the N>1 path of bench.py / shard.py).
The compute inside each rank is stood in by the CPU oracle (tests may use it); what is under test is the
sharding arithmetic, the ragged gather and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from primus_fhe_b200.shard import gather_batches, shard_range, shard_sizes


def test_shard_ranges_cover_exactly():
    for total in (0, 1, 7, 10000, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    assert shard_sizes(10000, 8) == [1250] * 8          # C5: 10k ciphertexts over 8 GPUs
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, q, log_n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    n = 1 << log_n
    rng = np.random.default_rng(123)                      # same data on every rank, each transforms its shard
    full = rng.integers(0, q, (total, n), dtype=np.uint64)
    b, e = shard_range(total, world, rank)
    mine = full[b:e].copy()
    t = O.U64NttTable(log_n, q)
    t.forward_batch(mine, 1)
    got = gather_batches(torch.from_numpy(mine.view(np.int64)), total)
    want = full.copy(); t.forward_batch(want, 1)
    assert np.array_equal(got.numpy().view(np.uint64), want)
    # max-over-ranks timing reduction used by bench.py
    tm = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    assert tm.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, 7, 1125899906826241, 6), nprocs=2, join=True)
