"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding, the gather and
the max-over-ranks timing reduction that bench.py / the data-parallel path use."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flownet2_b200 import sharding
        full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
        mine = sharding.shard_batch(full)
        lo, hi = sharding.shard_range(total, world, rank)
        assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
        # per-sample "layer": every rank processes only its shard; no collective on the data path
        out_local = mine * 2.0 + 1.0
        gathered = sharding.gather_batch(out_local, total)
        assert torch.equal(gathered, full * 2.0 + 1.0)
        # timing reduction: the job is as slow as the slowest rank
        assert sharding.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 2])
def test_shard_gather_world2(total):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total), nprocs=2, join=True)


def test_shard_range_partition():
    from flownet2_b200 import sharding
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
