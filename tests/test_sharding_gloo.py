"""The N>1 host logic on CPU: two processes over gloo compute their shards of the 64M instance, agree that the
shards tile the instance, split a spawn request consistently, and reduce a timing with MAX like bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevy_hanabi_b200.sharding import merge_counts, shard_range, split_spawn


def test_shard_range_tiles():
    for total in (0, 1, 7, 64 << 20, (64 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                a, b = shard_range(total, r, world)
                assert a == pos and b >= a
                pos = b
            assert pos == total
            sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_split_spawn():
    assert split_spawn(100, [50, 50]) == [50, 50]
    assert split_spawn(10, [50, 50]) == [5, 5]
    assert split_spawn(11, [50, 50]) == [6, 5]
    assert split_spawn(1000, [3, 0, 7]) == [3, 0, 7]        # capped by the free slots, excess dropped
    assert split_spawn(5, [0, 0, 9]) == [0, 0, 5]
    assert split_spawn(-4, [5, 5]) == [0, 0]
    s = split_spawn(12345, [100000, 1, 50000, 7])
    assert sum(s) == 12345 and all(x <= f for x, f in zip(s, [100000, 1, 50000, 7]))
    assert merge_counts([{"capacity": 4, "alive_count": 1}, {"capacity": 6, "alive_count": 5}])["alive_count"] == 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, end = shard_range(total, rank, world)
        mine = torch.tensor([first, end], dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        pos = 0
        for r in allr:
            assert int(r[0]) == pos
            pos = int(r[1])
        assert pos == total
        # every rank derives the same split from the gathered free-slot counts (8 integers, no data-path collective)
        free = torch.tensor([100 + 10 * rank], dtype=torch.int64)
        frees = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(frees, free)
        split = split_spawn(150, [int(f) for f in frees])
        assert sum(split) == 150
        chk = torch.tensor(split, dtype=torch.int64)
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(chk, ref)
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_process_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 64 << 20), nprocs=2, join=True)


def test_oracle_shard_fill_and_checksum_compose(orc):
    """The oracle side of hnb_slab_fill_c5_ex / hnb_slab_checksum_ex: shards hold the unsharded values, checksums add up."""
    import numpy as np
    from oracle import c_oracle as O
    P, seed = 9000, 31
    ref = np.zeros((P, 8), dtype=np.float32)
    orc.orc_fill_c5(O.ptr(ref), None, 0, P, seed, 0.2, 0.9)
    whole = orc.orc_checksum(O.ptr(ref), 0, P, 8)
    for world in (2, 3, 8):
        acc = 0
        for r in range(world):
            first, end = shard_range(P, r, world)
            shard = np.zeros((end - first, 8), dtype=np.float32)
            ind = np.zeros((end - first, 3), dtype=np.uint32)
            orc.orc_fill_c5_ex(O.ptr(shard), O.ptr(ind), 0, end - first, seed, 0.2, 0.9, first)
            np.testing.assert_array_equal(shard, ref[first:end])
            np.testing.assert_array_equal(ind[:, 0], np.arange(end - first, dtype=np.uint32))
            acc = (acc + orc.orc_checksum_ex(O.ptr(shard), 0, end - first, 8, first)) % 2**64
        assert acc == whole
