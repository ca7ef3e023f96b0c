"""world_size-2 gloo test of the multi-rank host logic (no GPU)."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("zip-ada_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [(i * 7919) % 65536 + 1 for i in range(1000)]
    mine = sh.shard_entries(sizes, world, rank)
    t = sh.max_over_ranks(1.0 + rank)                 # slowest rank = 2.0
    mbps = sh.aggregate_mbps(sum(sizes[i] for i in mine), 1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, t, mbps, gathered, sh.stream_seed(100, rank)))
    dist.destroy_process_group()


def test_two_ranks_partition_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = [(i * 7919) % 65536 + 1 for i in range(1000)]
    a, b = res[0][1], res[1][1]
    assert sorted(a + b) == list(range(1000)) and not set(a) & set(b)        # disjoint, complete
    la, lb = sum(sizes[i] for i in a), sum(sizes[i] for i in b)
    assert abs(la - lb) <= max(sizes)                                       # balanced
    assert res[0][2] == res[1][2] == 2.0                                    # max over ranks
    assert abs(res[0][3] - sum(sizes) / 1e6 / 2.0) < 1e-9 and res[0][3] == res[1][3]
    assert res[0][4] == res[1][4] == [a, b]                                 # every rank sees the same partition
    assert (res[0][5], res[1][5]) == (100, 101)


def test_single_process_degenerates():
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("zip-ada_b200.sharding")
    assert sh.shard_entries([5, 1, 9], 1, 0) == [0, 1, 2]
    assert sh.max_over_ranks(3.5) == 3.5
