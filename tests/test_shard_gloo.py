"""N > 1 host logic on CPU: two `gloo` ranks shard a batch, 'step' their slices and gather."""
import os
import socket

import numpy as np
import pytest

from opensc2_b200.shard import shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 8, 4096, 131072, 131075):
        for world in (1, 2, 3, 4, 8):
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    import torch.distributed as dist

    from opensc2_b200.shard import any_rank_failed, gather_to_all, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total, rank, world)
        # stand-in for the per-rank batched step: a result that encodes the global conductor index
        local = np.stack([np.arange(lo, hi, dtype=np.float64), np.arange(lo, hi, dtype=np.float64) ** 2], axis=1)
        full = gather_to_all(local, total)
        status = np.zeros(hi - lo, dtype=np.int32)
        if rank == 1:
            status[0] = -9                      # one singular conductor on rank 1 only
        failed = any_rank_failed(status)
        q.put((rank, full, failed))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_two_gloo_ranks_gather_results(total):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([np.arange(total, dtype=np.float64), np.arange(total, dtype=np.float64) ** 2], axis=1)
    for rank, full, failed in got:
        assert np.array_equal(full, expect), rank
        assert failed is True                   # every rank learns about rank 1's singular conductor
