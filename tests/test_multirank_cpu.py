"""CPU: the N > 1 path of bench.py.  Streams shard across ranks with no data-path collective; the only exchange is the
barrier and the MAX over ranks of the device time (bench.aggregate).  Exercised with world_size 2 over gloo."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r "measures" (r + 1) * 10 ms for 1000 samples per rank: the aggregate uses the slowest rank
    ms_local = (rank + 1) * 10.0
    ms_max, value = bench.aggregate(ms_local, samples_per_rank=1_000_000, world=world, dist=dist, device="cpu")
    shard = bench.shard_streams(10, rank, world)
    results[rank] = (ms_max, value, shard)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregate_and_sharding():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert set(results.keys()) == {0, 1}
    for r in (0, 1):
        ms_max, value, shard = results[r]
        assert ms_max == pytest.approx(20.0)
        assert value == pytest.approx(2 * 1_000_000 / 20e-3 / 1e6)   # whole-job MSamples/s over the slowest rank's time
    # every global stream id lands on exactly one rank
    s0, s1 = results[0][2], results[1][2]
    assert sorted(list(s0) + list(s1)) == list(range(10)) and not set(s0) & set(s1)
