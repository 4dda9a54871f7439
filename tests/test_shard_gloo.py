"""Multi-process host logic of the N>1 path on CPU (gloo, world size 2): the stream partition is disjoint and
complete, every rank reaches the timing barrier, and max-over-ranks reduction works as bench.py uses it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libmpegh_b200 import shard


def test_stream_range_properties():
    for n in (0, 1, 7, 256, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = shard.stream_range(r, world, n)
                assert 0 <= b <= e <= n
                seen.extend(range(b, e))
                for s in (b, e - 1):
                    if b < e:
                        assert shard.owner_of(s, world, n) == r
            assert seen == list(range(n))
    assert shard.stream_range(0, 8, 4096) == (0, 512) and shard.stream_range(7, 8, 4096) == (3584, 4096)
    with pytest.raises(ValueError):
        shard.stream_range(2, 2, 10)


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard.stream_range(rank, world, n_streams)
    # each rank "processes" its streams: a per-stream checksum that depends only on the stream index
    local = torch.zeros(n_streams, dtype=torch.int64)
    local[b:e] = torch.arange(b, e, dtype=torch.int64) * 2654435761 % 1000003
    dist.barrier()
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)  # pretend elapsed ms
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cover = torch.zeros(n_streams, dtype=torch.int64)
    cover[b:e] = 1
    dist.all_reduce(cover)           # bookkeeping only: the data path itself has no collective
    dist.all_reduce(local)
    q.put((rank, float(t.item()), cover.numpy().copy(), local.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world, n = 2, 513
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(n, dtype=np.int64) * 2654435761 % 1000003
    for rank, tmax, cover, local in res:
        assert tmax == 11.0                      # max over ranks
        assert (cover == 1).all()                # disjoint and complete
        assert np.array_equal(local, want)       # union of the shards == the unsharded result


def test_numa_helpers():
    from libmpegh_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard._parse_cpulist("") == []
    assert shard.gpu_numa_cpus("00000000:FF:1F.7") is None  # no such device: the caller keeps its affinity
