"""CPU test of the N>1 host logic with world_size 2 over gloo: parameter broadcast, frame sharding,
max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import checkers
import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 holds the real parameter block, the others start from a different one
        p = checkers.stereomapper(255) if rank == 0 else checkers.middlebury()
        p = sharding.broadcast_params(p, torch.device("cpu"))
        mine = sharding.shard_frames(11, rank, world)
        (tmax,) = sharding.max_over_ranks([10.0 + rank], torch.device("cpu"))
        (total,) = sharding.sum_over_ranks([len(mine)], torch.device("cpu"))
        out[rank] = (bytes(p), mine, tmax, total)
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    want = bytes(checkers.stereomapper(255))
    assert out[0][0] == want and out[1][0] == want            # one broadcast of the parameter block
    assert sorted(out[0][1] + out[1][1]) == list(range(11))   # disjoint cover, frame i -> rank i mod 2
    assert out[0][1] == [0, 2, 4, 6, 8, 10]
    assert out[0][2] == out[1][2] == 11.0                     # max over ranks
    assert out[0][3] == out[1][3] == 11.0


def test_single_process_is_identity():
    p = checkers.stereomapper(128)
    q = sharding.broadcast_params(p, torch.device("cpu"))
    assert bytes(p) == bytes(q)
    assert sharding.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks([3.5], torch.device("cpu")) == [3.5]
