"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): image sharding needs no
communication and covers the batch exactly once; the data-parallel gradient exchange is one flat
sum all-reduce whose 1/world scale is handed to the fused Adam."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_style_transfer_gan_b200.stylize import shard_range


def test_shard_range_partitions_every_batch():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n)), (n, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_style_transfer_gan_b200.enhanced_train import allreduce_flat_
    from multi_style_transfer_gan_b200.stylize import shard_range as sr
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)        # per-rank "flat gradient"
    scale = allreduce_flat_(g)
    lo, hi = sr(10, rank, world)
    # image sharding: each rank "stylises" its shard, results are gathered (test-only collective)
    out = torch.zeros(10)
    out[lo:hi] = torch.arange(lo, hi, dtype=torch.float32) + 100
    dist.all_reduce(out)
    q.put((rank, g.tolist(), scale, out.tolist()))
    dist.destroy_process_group()


def test_flat_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, g, scale, out in res:
        assert scale == 0.5
        assert g == [float(i) * 3 for i in range(10)]            # (1 + 2) * i on every rank
        assert out == [float(i) + 100 for i in range(10)]        # every image produced exactly once


def test_allreduce_noop_without_process_group():
    from multi_style_transfer_gan_b200.enhanced_train import allreduce_flat_
    g = torch.ones(4)
    assert allreduce_flat_(g) == 1.0 and g.tolist() == [1.0] * 4
