"""CPU tests of the multi-rank host logic (no GPU): the two scalar exchanges of one stream over several ranks
(zip-ada_b200/sharding.py) run over gloo with the oracle standing in for the CUDA engine; the assembled
stream must equal the sequential oracle's, whatever the number of ranks."""
import ctypes as C
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stream():
    import corpus
    # stripes of text / random / sparse: chunks of very different raw sizes, real segmentations
    return corpus.mixed(4_300_000, 11, 1 << 19)


def _worker(rank, world, port, q, bounds):
    import oracle_lib as orc
    sh = importlib.import_module("zip-ada_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = _stream()
    n = data.size
    margin = 9_004_096
    lo, hi = bounds[rank], min(n, bounds[rank + 1] + (margin if rank + 1 < world else 0))
    local = np.ascontiguousarray(data[lo:hi])               # the rank holds only its own bytes
    out = np.zeros(n + 100_000, np.uint8)
    comm = sh.TorchComm(dist, rank, world)
    eng = sh.OracleShardEngine(orc.lib(), 9, threads=2)
    res = sh.encode_sharded(eng, comm, local.ctypes.data, False, n, n, bounds, (lo, hi), out.ctypes.data, False, out.size)
    sizes = [(i * 7919) % 65536 + 1 for i in range(1000)]
    q.put((rank, res, out[:res["length"]].tobytes(), sh.assign_entries(sizes, world)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bounds", [(2, [0, 2_000_000, 4_300_000]), (3, [0, 1_300_000, 1_400_000, 4_300_000])])
def test_one_stream_over_ranks_equals_sequential_oracle(world, bounds):
    import bz2
    import oracle_lib as orc
    b2 = importlib.import_module("zip-ada_b200")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, bounds)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = _stream()
    ref = orc.encode_stream(data, 9, data.size)
    total = res[0][1]["total_length"]
    assert all(r[1]["total_length"] == total for r in res)
    got = b2.assemble_pieces([(r[1]["byte_offset"], r[2]) for r in res], total).tobytes()
    assert got == ref
    assert bz2.decompress(got) == data.tobytes()
    # the chain: every rank starts where the one before handed off; the middle rank of the 3-rank case owns
    # a range smaller than a chunk and may own no chunk at all
    for a, b in zip(res, res[1:]):
        assert a[1]["handoff"] == b[1]["entry"]
    assert res[0][1]["entry"] == 0 and res[-1][1]["handoff"] == data.size
    # independent streams (archive entries): same table on every rank, disjoint, complete, balanced
    sizes = [(i * 7919) % 65536 + 1 for i in range(1000)]
    parts = res[0][3]
    assert all(r[3] == parts for r in res)
    assert sorted(sum(parts, [])) == list(range(1000))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes) + 4096


def test_plan_and_resolve_host_functions():
    """b2_shard_plan / b2_shard_resolve are host-only entry points of libb2gpu.so: they run without a GPU."""
    b2 = importlib.import_module("zip-ada_b200")
    sh = importlib.import_module("zip-ada_b200.sharding")
    n = (4 << 30) + 12345
    for world in (1, 2, 4, 8):
        bounds = b2.shard_plan(n, world, 9, 30)
        assert bounds[0] == 0 and bounds[-1] == n and len(bounds) == world + 1
        assert all(a < b for a, b in zip(bounds, bounds[1:]))
        assert all(b % 4096 == 0 for b in bounds[1:-1])
        shares = [b - a for a, b in zip(bounds, bounds[1:])]
        assert all(x >= y - 4096 for x, y in zip(shares, shares[1:]))            # later ranks start later: smaller shares
        bounds2, spans = sh.plan(n, world, 9, b2.lib(), 30)
        assert bounds2 == bounds and spans[-1][1] == n
    rng = np.random.default_rng(5)
    links, rows = [], []
    for _ in range(5):
        l = b2.ShardLink()
        for k in range(8):
            l.total_bits[k] = int(rng.integers(0, 1 << 40)); l.crc_rot[k] = int(rng.integers(0, 32)); l.crc_fold[k] = int(rng.integers(0, 1 << 32))
        links.append(l)
        rows.append(list(l.total_bits) + list(l.crc_rot) + list(l.crc_fold))
    assert b2.shard_resolve(links) == tuple(sh.resolve(rows))


def test_single_rank_degenerates():
    import oracle_lib as orc
    sh = importlib.import_module("zip-ada_b200.sharding")
    import corpus
    data = corpus.markov_text(1_100_000, 3)
    out = np.zeros(data.size + 100_000, np.uint8)
    eng = sh.OracleShardEngine(orc.lib(), 9, threads=4)
    res = sh.encode_sharded(eng, sh.LocalComm(), data.ctypes.data, False, data.size, data.size, [0, data.size], (0, data.size),
                            out.ctypes.data, False, out.size)
    assert out[:res["length"]].tobytes() == orc.encode_stream(data, 9, data.size)
    assert sh.assign_entries([5, 1, 9], 1) == [[0, 1, 2]]
