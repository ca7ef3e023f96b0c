"""One BZip2 stream over several ranks (one process per GPU): the host side of include/b2gpu.h's b2_shard_*.

SURVEY.md §8e: chunks are independent once their start is known; the reference's own parallelism is four
tasks per chunk (zip_lib/bzip2-encoding.adb:1226-1303).  Here a stream is cut into contiguous byte ranges, one
per rank; a rank owns the chunks that START in its range.  Two scalar exchanges, no bulk data between ranks:

  1. the cutting is a chain (Data_Acquisition, :1160-1208): rank r receives from rank r-1 the stream offset
     where its first chunk starts (8 bytes, point to point) and sends rank r+1 its own;
  2. a chunk's winner depends on the incoming bit offset mod 8 (:1319-1325), the combined CRC is folded block by
     block (:990): every rank publishes, per possible incoming offset, the bits it appends and its CRC fold
     (128 bytes, all-gather); all ranks compose them in rank order (b2_shard_resolve).

`encode_sharded` is what bench.py runs under torchrun; the engine is the CUDA Encoder (zip-ada_b200), or — in
the CPU tests of this protocol only — a stand-in built on the oracle.  `assign_entries` is the other natural
shard of the archive workloads (independent streams, longest first).
"""
import ctypes as C

import numpy as np


# ---- communication: the two exchanges, over torch.distributed (nccl on the GPU box, gloo in the CPU tests) ----
class TorchComm:
    """group: the process group of the exchanges.  On a GPU box pass a gloo group with device=None: the scalars
    then travel through the host and no exchange has to wait for a free SM (a NCCL point-to-point kernel queues
    behind whatever the device is running)."""

    def __init__(self, dist, rank, world, device=None, group=None):
        self.dist, self.rank, self.world, self.device, self.group = dist, rank, world, device, group

    def _t(self, values):
        import torch
        return torch.tensor(values, dtype=torch.int64, device=self.device)

    def send_u64(self, dst, v):
        self.dist.send(self._t([int(v)]), dst, group=self.group)

    def recv_u64(self, src):
        t = self._t([0])
        self.dist.recv(t, src, group=self.group)
        return int(t.item())

    def all_gather_i64(self, values):
        t = self._t([int(v) for v in values])
        outs = [self._t([0] * len(values)) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        return [[int(x) for x in o.tolist()] for o in outs]


class LocalComm:
    """World of one."""
    rank, world = 0, 1

    def all_gather_i64(self, values):
        return [[int(v) for v in values]]


def plan(n, world, level, lib=None, stagger_permille=-1):
    """Byte ranges of the ranks (b2_shard_plan) and the bytes every rank needs: (bounds, [(lo, hi)])."""
    if lib is None:
        from . import lib as _lib
        lib = _lib()
    b = (C.c_uint64 * (world + 1))()
    rc = lib.b2_shard_plan(C.c_uint64(n), world, level, stagger_permille, b)
    assert rc == 0
    bounds = [int(x) for x in b]
    margin = int(lib.b2_shard_margin(level))
    spans = [(bounds[r], min(n, bounds[r + 1] + (margin if r + 1 < world else 0))) for r in range(world)]
    return bounds, spans


def _link_to_list(link):
    return list(link.total_bits) + list(link.crc_rot) + list(link.crc_fold)


def resolve(rows):
    """rows[r] = 24 integers of rank r's link -> (bit_offsets, crcs) with world + 1 entries (b2_shard_resolve,
    restated here so that the CPU tests check the library's version against it)."""
    bit, crc = [32], [0]
    for row in rows:
        ph = bit[-1] & 7
        bit.append(bit[-1] + row[ph])
        k = row[8 + ph] & 31
        c = crc[-1]
        c = ((c << k) | (c >> (32 - k))) & 0xFFFFFFFF if k else c
        crc.append(c ^ row[16 + ph])
    return bit, crc


def encode_sharded(engine, comm, slice_ptr, in_is_device, n, size_hint, bounds, span, out_ptr, out_is_device, out_cap):
    """This rank's part of one stream.  slice_ptr points at stream byte span[0]; the rank holds the bytes
    [span[0], span[1]).  Returns dict(byte_offset, length, first_byte, last_byte, total_length)."""
    import time
    r, w = comm.rank, comm.world
    lo, hi = span
    t0 = time.perf_counter()
    engine.shard_open(slice_ptr, in_is_device, lo, hi - lo, n, size_hint, bounds[r + 1])
    entry = 0 if r == 0 else comm.recv_u64(r - 1)
    t1 = time.perf_counter()
    handoff = engine.shard_cut(entry)
    if r + 1 < w:
        comm.send_u64(r + 1, handoff)
    t2 = time.perf_counter()
    link = engine.shard_encode()
    t3 = time.perf_counter()
    rows = comm.all_gather_i64(_link_to_list(link))
    bit, crc = resolve(rows)
    t4 = time.perf_counter()
    off, ln, fb, lb = engine.shard_finish(bit[r], crc[r], out_ptr, out_is_device, out_cap)
    t5 = time.perf_counter()
    return dict(byte_offset=off, length=ln, first_byte=fb, last_byte=lb, total_length=(bit[w] + 80 + 7) >> 3,
                entry=entry, handoff=handoff,
                phases_ms=dict(open_and_wait_for_entry=1000 * (t1 - t0), cut=1000 * (t2 - t1), encode=1000 * (t3 - t2),
                               wait_for_links=1000 * (t4 - t3), finish=1000 * (t5 - t4)))


class OracleShardEngine:
    """Stand-in for the CUDA Encoder in the CPU tests of the protocol (tests/test_sharding.py): the same four
    calls on top of the oracle (oracle/b2_oracle.cpp, orc_shard_*).  Test infrastructure, never the product."""

    class _Link(C.Structure):
        _fields_ = [("total_bits", C.c_uint64 * 8), ("crc_rot", C.c_uint32 * 8), ("crc_fold", C.c_uint32 * 8)]

    def __init__(self, orc_lib, level=9, threads=2):
        self.lib, self.level, self.threads, self.h = orc_lib, level, threads, None
        self.lib.orc_shard_encode.restype = C.c_void_p

    def shard_open(self, in_ptr, in_is_device, base, n_local, stream_size, size_hint, own_end):
        self.args = (in_ptr, base, n_local, stream_size, size_hint, own_end)

    def shard_cut(self, entry):
        in_ptr, base, n_local, stream_size, size_hint, own_end = self.args
        ho = C.c_uint64(0)
        self.link = self._Link()
        self.h = self.lib.orc_shard_encode(C.c_void_p(in_ptr), C.c_uint64(base), C.c_uint64(n_local), C.c_uint64(stream_size), self.level,
                                           C.c_int64(size_hint), 0, self.threads, C.c_uint64(entry), C.c_uint64(own_end), C.byref(ho),
                                           C.byref(self.link))
        return ho.value

    def shard_encode(self):
        return self.link

    def shard_finish(self, bit_offset, crc_in, out_ptr, out_is_device, out_cap):
        off, ln = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.orc_shard_finish(C.c_void_p(self.h), C.c_uint64(bit_offset), C.c_uint32(crc_in), C.c_void_p(out_ptr), C.c_uint64(out_cap),
                                       C.byref(off), C.byref(ln))
        assert rc == 0, rc
        self.lib.orc_shard_free(C.c_void_p(self.h))
        self.h = None
        buf = (C.c_uint8 * max(1, ln.value)).from_address(out_ptr)
        return off.value, ln.value, (buf[0] if ln.value else 0), (buf[ln.value - 1] if ln.value else 0)


# ---- independent streams (archive entries, BASELINE.json configs[4]) --------------------------------------
def assign_entries(sizes, world):
    """Entry indices of every rank: longest first onto the least loaded rank, no communication (every rank
    computes the same table).  Returns a list of `world` index lists, each in ascending order."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i]) + 4096           # per-entry cost besides its bytes
    return [sorted(x) for x in out]
