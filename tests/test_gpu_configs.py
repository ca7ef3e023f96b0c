"""The BASELINE.json configurations as parity runs on the GPU (`-m gpu`):

  configs[0]  64 MiB text stream, size_hint = size: GPU == oracle (SHA-256 of the committed golden, made by the
              oracle with tools/make_golden_sha.py) and libbz2 decodes it
  configs[2]  mixed corpus (256 MiB here; 4 GiB in bench.py --config mixed): same
  configs[3]  pathological blocks at full block size (SURVEY.md §8d list): every tap of the block == oracle,
              doubling rounds and sort time reported
  one stream over several handles (b2_shard_*, b2_encode_stream_multi): identical bytes for every shard count
  b2_get_segments == the oracle's Segment_by_Entropy cut lists
"""
import bz2
import ctypes as C
import hashlib
import json
import os
import time

import numpy as np
import pytest

import corpus
import datagen
import oracle_lib as orc
from test_gpu_parity import _cmp_block, _cmp_stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_sha.json")))
MiB = 1 << 20


def _gpu_corpus(name, n, seed):
    import torch
    return corpus.workload(name, n, seed, torch, "cuda").cpu().numpy()


@pytest.mark.parametrize("key", ["markov:%d:5eed0001:9" % (64 * MiB), "mixed:%d:5eed0004:9" % (256 * MiB)])
def test_config_stream_equals_golden(enc9, key):
    name, n, seed, level = key.split(":")
    n, seed = int(n), int(seed, 16)
    data = _gpu_corpus(name, n, seed)
    g = GOLDEN[key]
    assert hashlib.sha256(data.tobytes()).hexdigest() == g["input_sha256"], "the corpus generator changed"
    out = enc9.encode(data, n).tobytes()
    assert len(out) == g["bytes"], (len(out), g["bytes"])
    assert hashlib.sha256(out).hexdigest() == g["sha256"]
    assert bz2.decompress(out) == data.tobytes()


def test_config1_stream_equals_oracle_chunk_by_chunk(enc9):
    """16 MiB of the config-1 stream against the oracle run here (chunk-parallel), with the per-chunk trace."""
    data = corpus.markov_text(16 * MiB, 0x5EED0001)
    out = enc9.encode(data, data.size).tobytes()
    ref, otr = orc.encode_stream(data, 9, data.size, 0, want_trace=True, threads=os.cpu_count() or 4)
    gtr = enc9.trace()
    assert len(gtr) == len(otr)
    for k, (a, b) in enumerate(zip(gtr, otr)):
        assert (a.start, a.len, a.dyn_capacity, a.n_seg1, a.n_seg2, a.winner) == (b.start, b.len, b.dyn_capacity, b.n_seg1, b.n_seg2, b.winner), k
        assert list(a.bits) == list(b.bits), k
    assert out == ref


# ---- configs[3]: pathological blocks at full block size ----------------------------------------------------
def _distinct_neighbours(n, seed):
    """n bytes, no two neighbours equal (no RLE1 effect), then a 259-run."""
    r = datagen.random_bytes(n, seed, 0, 254).astype(np.int64)
    same = np.zeros(n, bool)
    same[1:] = r[1:] == r[:-1]
    r[same] = 255                                         # 255 never appears otherwise: breaks every pair
    return r.astype(np.uint8)


PATHOLOGICAL = {
    "zeros_9M_rawcap": lambda: np.zeros(9_000_000, np.uint8),                                    # RLE1 output of period 5
    "abc_x300000": lambda: np.tile(np.frombuffer(b"abc", np.uint8), 300_000),                    # period 3
    "period_half": lambda: np.tile(datagen.random_bytes(449_000, 41, 1, 250), 2),                # period N/2
    "unit100k_x9": lambda: np.tile(datagen.random_bytes(99_990, 42), 9),                         # long exact repeats
    "distinct_then_run259": lambda: np.concatenate([_distinct_neighbours(899_995, 43), np.full(259, 7, np.uint8)]),
    "fibonacci_880k": lambda: np.frombuffer(_fib(880_000), np.uint8),                             # deepest repeats two symbols allow
}


def _fib(n):
    a, b = b"b", b"a"
    while len(b) < n:
        a, b = b, b + a
    return b[:n]


@pytest.mark.parametrize("name", list(PATHOLOGICAL))
def test_pathological_block_full_size(enc9, name, record_property):
    data = PATHOLOGICAL[name]()
    enc9.reset_stats()
    enc9.set_timing(2)
    try:
        _cmp_block(enc9, data, name)
        st = enc9.stats()
    finally:
        enc9.set_timing(0)
    rep = {"case": name, "raw_bytes": int(data.size), "doubling_rounds": int(st.sort_rounds), "sort_ms": round(st.stage_ms[2], 2),
           "rows_round0": int(st.sort_elems_round0), "rows_later_rounds": int(st.sort_elems_later)}
    record_property("sort_report", json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pathological_report.jsonl"), "a") as f:
        f.write(json.dumps(rep) + "\n")
    # the reference stops its comparator at the first difference: a fully periodic block never differs, so
    # the doubling must run until the compared prefix covers the block (and the origin pointer still be right)
    assert st.sort_rounds >= 1


def test_pathological_streams(enc9):
    for name in ("zeros_9M_rawcap", "abc_x300000", "unit100k_x9"):
        data = PATHOLOGICAL[name]()
        _cmp_stream(enc9, data, data.size, name)


# ---- Segment_by_Entropy cut positions (A3) -------------------------------------------------------------------
def test_get_segments_equals_oracle(enc9):
    data = corpus.mixed(6 * MiB, 0x77, 1 << 19)           # stripe changes inside every chunk
    enc9.encode(data, data.size)
    tr = enc9.trace()
    assert len(tr) >= 5
    seen_real = 0
    for c, t in enumerate(tr):
        chunk = data[t.start:t.start + t.len]
        for profile in (0, 1):
            cuts = enc9.segments(c, profile)
            ref = orc.segment(chunk, profile)
            assert np.array_equal(cuts, ref), (c, profile, cuts[:8], ref[:8])
            seen_real += len(ref) > 1
    assert seen_real >= 4


# ---- one stream over several handles ---------------------------------------------------------------------
def _shards_by_hand(b2mod, encs, data, hint, bounds):
    """The protocol of include/b2gpu.h, call by call, with the handles of this process."""
    n = data.size
    ns = len(encs)
    margin = int(b2mod.lib().b2_shard_margin(9))
    spans = [(bounds[r], min(n, bounds[r + 1] + (margin if r + 1 < ns else 0))) for r in range(ns)]
    for r, e in enumerate(encs):
        lo, hi = spans[r]
        e.shard_open(data[lo:].ctypes.data, False, lo, hi - lo, n, hint, bounds[r + 1])
    entry, links = 0, []
    entries = []
    for e in encs:
        entries.append(entry)
        entry = e.shard_cut(entry)
    assert entry == n
    for e in encs:
        links.append(e.shard_encode())
    bit, crc = b2mod.shard_resolve(links)
    total = (bit[-1] + 80 + 7) >> 3
    pieces = []
    for r, e in enumerate(encs):
        buf = np.zeros(n + n // 50 + 200_000, np.uint8)
        off, ln, fb, lb = e.shard_finish(bit[r], crc[r], buf.ctypes.data, False, buf.size)
        assert ln == 0 or (buf[0] == fb and buf[ln - 1] == lb)
        pieces.append((off, buf[:ln].copy()))
    return b2mod.assemble_pieces(pieces, total).tobytes(), entries


SHARD_CASES = {
    "mixed_6M_3": lambda: (corpus.mixed(6_000_000, 5, 1 << 19), 6_000_000, [0, 2_000_000 & ~4095, 4_000_000 & ~4095, 6_000_000]),
    "markov_2p5M_balanced_2": lambda: (corpus.markov_text(2_500_000, 6), 2_500_000, [0, 1_200_000 & ~4095, 2_500_000]),
    "zeros_12M_4": lambda: (np.zeros(12_000_000, np.uint8), 12_000_000, [0, 3_002_368, 6_000_640, 9_003_008, 12_000_000]),   # one chunk spans three ranges
    "nohint_tiny_middle_3": lambda: (corpus.markov_text(3_000_000, 8), -1, [0, 1_503_232, 1_507_328, 3_000_000]),
}


@pytest.mark.parametrize("name", list(SHARD_CASES))
def test_shards_equal_single_handle_and_oracle(b2mod, enc9, name):
    data, hint, bounds = SHARD_CASES[name]()
    ref = orc.encode_stream(data, 9, hint, threads=os.cpu_count() or 4)
    one = enc9.encode(data, hint).tobytes()
    assert one == ref
    encs = [b2mod.Encoder(9, 0) for _ in range(len(bounds) - 1)]
    try:
        got, entries = _shards_by_hand(b2mod, encs, data, hint, bounds)
    finally:
        for e in encs:
            e.close()
    assert got == ref, (name, len(got), len(ref), entries)
    assert bz2.decompress(got) == data.tobytes()


@pytest.mark.parametrize("n_handles", [2, 3])
def test_encode_stream_multi(b2mod, enc9, n_handles, monkeypatch):
    """b2_encode_stream_multi with several handles (all on device 0 here; one per device on a multi-GPU box)."""
    import torch
    ndev = torch.cuda.device_count()
    monkeypatch.setenv("B2GPU_SHARD_MIN_BYTES", str(1 << 20))
    data = corpus.mixed(40 * MiB, 0x99, 1 << 21)
    ref = enc9.encode(data, data.size).tobytes()
    encs = [b2mod.Encoder(9, r % ndev) for r in range(n_handles)]
    try:
        out = b2mod.encode_multi(encs, data, data.size).tobytes()
        out2 = b2mod.encode_multi(encs, data, data.size).tobytes()          # handles are reusable
    finally:
        for e in encs:
            e.close()
    assert out == ref and out2 == ref
    assert bz2.decompress(out) == data.tobytes()


# ---- feedback / abort through the batched calls (zip.ads:301-306, zip-compress-bzip2_e.adb:78-96) --------------
def test_progress_callback_and_abort(b2mod, monkeypatch):
    monkeypatch.setenv("B2GPU_BATCH_POSITIONS", str(4 << 20))          # several device batches for a small input
    data = corpus.markov_text(12 * MiB, 0x31)
    ref = None
    with b2mod.Encoder(9, 0) as e:
        calls = []
        e.set_progress(lambda done, total: calls.append((done, total)) or False)
        out = e.encode(data, data.size).tobytes()
        assert len(calls) >= 3 and calls[-1][0] == calls[-1][1] == data.size
        assert all(a[0] < b[0] for a, b in zip(calls, calls[1:]))
        e.set_progress(None)
        ref = e.encode(data, data.size).tobytes()
        assert out == ref                                              # the callback does not change the stream
        # User_abort: the second call asks to stop
        seen = []
        e.set_progress(lambda done, total: seen.append(done) or len(seen) >= 2)
        with pytest.raises(b2mod.B2Error) as ei:
            e.encode(data, data.size)
        assert "error 12" in str(ei.value) and len(seen) == 2
        # the handle stays usable; archives report progress too
        e.set_progress(None)
        assert e.encode(data, data.size).tobytes() == ref
        zc = []
        e.set_progress(lambda done, total: zc.append((done, total)) or False)
        arch = e.zip_create([("a.txt", data[:3 * MiB]), ("b.bin", datagen.random_bytes(2 * MiB, 5)), ("c.txt", data[3 * MiB:9 * MiB])])
        assert zc and zc[-1][0] == zc[-1][1] == 3 * MiB + 2 * MiB + 6 * MiB
        import io, zipfile
        assert zipfile.ZipFile(io.BytesIO(arch.tobytes())).testzip() is None
