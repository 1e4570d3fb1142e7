"""The N>1 host logic on CPU: two gloo processes shard a batch of recordings with no data-path collective and
agree on the aggregate (sum of samples / max of times)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from noaa_apt_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.rank_recordings(7, rank, world)
    samples = 1000 * len(mine)
    ms = 10.0 * (rank + 1)                      # rank 1 is the slow one
    value, worst = sharding.aggregate_throughput(samples, ms, dist)
    dist.barrier()
    out.put((rank, mine, value, worst))
    dist.destroy_process_group()


def test_two_ranks_shard_and_aggregate():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, v0, w0), (r1, mine1, v1, w1) = res
    assert mine0 == [0, 2, 4, 6] and mine1 == [1, 3, 5]           # disjoint, covering
    assert w0 == w1 == 20.0                                       # MAX over ranks
    assert v0 == v1 == pytest.approx(7000 / 20e-3 / 1e6)          # all samples / slowest rank


def test_device_stream_assignment_matches_batch_rule():
    sys.path.insert(0, ROOT)
    from noaa_apt_b200 import sharding
    # 512 recordings over 8 GPUs x 4 streams (BASELINE configs[4]): 64 per GPU, every slot used equally
    slots = {}
    for i in range(512):
        slots.setdefault(sharding.assign(i, 8, 4), []).append(i)
    assert len(slots) == 32 and all(len(v) == 16 for v in slots.values())
    assert sharding.assign(0, 8, 4) == (0, 0) and sharding.assign(9, 8, 4) == (1, 1)
