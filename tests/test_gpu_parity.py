"""Parity of the CUDA path (through the C ABI) against the oracle, bit-exact.  `-m gpu`."""
import bz2

import numpy as np
import pytest

import datagen
import oracle_lib as orc

pytestmark = pytest.mark.gpu


def _cmp_block(enc, data, tag):
    g = enc.dbg_block(data)
    o = orc.encode_block(data, 9)
    gi, oi = g["info"], o["info"]
    assert gi.n_rle == oi.n_rle, (tag, "n_rle", gi.n_rle, oi.n_rle)
    assert np.array_equal(g["rle"], o["rle"]), (tag, "rle1 bytes differ at", int(np.argmax(g["rle"] != o["rle"])))
    assert gi.crc == oi.crc, (tag, "crc", hex(gi.crc), hex(oi.crc))
    assert np.array_equal(g["bwt"], o["bwt"]), (tag, "bwt differs at", int(np.argmax(g["bwt"] != o["bwt"])))
    assert gi.origin == oi.origin, (tag, "origin", gi.origin, oi.origin)
    assert gi.n_mtf == oi.n_mtf, (tag, "n_mtf", gi.n_mtf, oi.n_mtf)
    assert np.array_equal(g["mtf"], o["mtf"]), (tag, "mtf differs at", int(np.argmax(g["mtf"] != o["mtf"])))
    assert (gi.max_len, gi.ec_count, gi.sample_width) == (oi.max_len, oi.ec_count, oi.sample_width), \
        (tag, "triple", (gi.max_len, gi.ec_count, gi.sample_width, gi.cost), (oi.max_len, oi.ec_count, oi.sample_width, oi.cost))
    assert gi.cost == oi.cost, (tag, "cost", gi.cost, oi.cost)
    assert np.array_equal(g["sel"], o["sel"]), (tag, "selectors differ at", int(np.argmax(g["sel"] != o["sel"])))
    assert np.array_equal(g["lens"], o["lens"]), (tag, "code lengths differ")
    assert gi.bits == oi.bits, (tag, "bits", gi.bits, oi.bits)
    assert np.array_equal(g["bits"], o["bits"]), (tag, "bitstream differs at byte", int(np.argmax(g["bits"] != o["bits"])))


BLOCK_CASES = {
    "empty": lambda: np.zeros(0, np.uint8),
    "one": lambda: np.array([65], np.uint8),
    "two_same": lambda: np.array([7, 7], np.uint8),
    "abc": lambda: np.frombuffer(b"abc", np.uint8),
    "run4": lambda: np.full(4, 9, np.uint8),
    "run259": lambda: np.full(259, 200, np.uint8),
    "run260": lambda: np.full(260, 200, np.uint8),
    "run1000": lambda: np.full(1000, 0, np.uint8),
    "period3": lambda: np.tile(np.frombuffer(b"abc", np.uint8), 5000),
    "period2_long": lambda: np.tile(np.frombuffer(b"xy", np.uint8), 40000),
    "text_3k": lambda: datagen.text(3000, 1),
    "text_60k": lambda: datagen.text(60000, 2),
    "text_300k": lambda: datagen.text(300000, 3),
    "random_20k": lambda: datagen.random_bytes(20000, 4),
    "random_alpha_50k": lambda: datagen.random_bytes(50000, 5, 65, 90),
    "sparse_100k": lambda: datagen.sparse_binary(100000, 6),
    "all256_twice": lambda: np.tile(np.arange(256, dtype=np.uint8), 2),
    "repeat_unit": lambda: np.tile(datagen.random_bytes(5000, 7), 9),
    "runs_mixed": lambda: np.repeat(datagen.random_bytes(3000, 8, 0, 3), datagen.random_bytes(3000, 9, 1, 12)),
}


@pytest.mark.parametrize("name", list(BLOCK_CASES))
def test_block_taps(enc9, name):
    _cmp_block(enc9, BLOCK_CASES[name](), name)


def test_block_full_size_text(enc9):
    _cmp_block(enc9, datagen.text(890000, 11), "text_890k")


def _cmp_stream(enc, data, hint, tag):
    out = enc.encode(data, hint).tobytes()
    ref, otr = orc.encode_stream(data, 9, hint, 0, want_trace=True)
    gtr = enc.trace()
    assert len(gtr) == len(otr), (tag, "chunks", len(gtr), len(otr))
    for k, (a, b) in enumerate(zip(gtr, otr)):
        assert (a.start, a.len, a.dyn_capacity) == (b.start, b.len, b.dyn_capacity), (tag, "chunk", k, (a.start, a.len, a.dyn_capacity), (b.start, b.len, b.dyn_capacity))
        assert (a.n_seg1, a.n_seg2) == (b.n_seg1, b.n_seg2), (tag, "segments", k, (a.n_seg1, a.n_seg2), (b.n_seg1, b.n_seg2))
        assert list(a.bits) == list(b.bits), (tag, "tactic bits", k, list(a.bits), list(b.bits))
        assert a.winner == b.winner, (tag, "winner", k, a.winner, b.winner)
    assert out == ref, (tag, "stream bytes differ", len(out), len(ref))
    if len(data):
        assert bz2.decompress(out) == bytes(data), (tag, "does not decode")


STREAM_CASES = {
    "tiny": lambda: (np.frombuffer(b"hello hello hello", np.uint8), 17),
    "three_bytes": lambda: (np.frombuffer(b"abc", np.uint8), 3),
    "text_200k_hint": lambda: (datagen.text(200000, 21), 200000),
    "text_200k_nohint": lambda: (datagen.text(200000, 21), -1),
    "text_1p1M_balanced": lambda: (datagen.text(1_100_000, 22), 1_100_000),   # last-two-blocks balancing window
    "text_2p5M": lambda: (datagen.text(2_500_000, 23), 2_500_000),
    "mixed_3M": lambda: (datagen.mixed(3_000_000, 300_000, 24), 3_000_000),
    "zeros_2M": lambda: (np.zeros(2_000_000, np.uint8), 2_000_000),
    "random_1M": lambda: (datagen.random_bytes(1_000_000, 25), -1),
    "runs4": lambda: (np.repeat(datagen.random_bytes(300_000, 26), 4), 1_200_000),   # worst-case RLE1 expansion
    "runs_long": lambda: (np.repeat(datagen.random_bytes(4000, 27, 0, 3), np.random.default_rng(28).integers(1, 3000, 4000)), -1),
    "zeros_9p5M_rawcap": lambda: (np.zeros(9_500_000, np.uint8), 9_500_000),          # 10 x capacity raw limit
    "runs259": lambda: (np.repeat(datagen.random_bytes(12_000, 29), 259), 12_000 * 259),
    "runs260_then_text": lambda: (np.concatenate([np.repeat(datagen.random_bytes(9_000, 30), 260), datagen.text(700_000, 31)]), -1),
    "text_then_zeros_then_text": lambda: (np.concatenate([datagen.text(880_000, 32), np.zeros(3_000_000, np.uint8), datagen.text(500_000, 33)]), 4_380_000),
}


@pytest.mark.parametrize("name", list(STREAM_CASES))
def test_stream(enc9, name):
    data, hint = STREAM_CASES[name]()
    _cmp_stream(enc9, data, hint, name)


def test_empty_stream(enc9):
    out = enc9.encode(np.zeros(0, np.uint8), 0).tobytes()
    assert out == orc.encode_stream(b"", 9, 0)


@pytest.mark.parametrize("level", [1, 4])
def test_other_levels(b2mod, level):
    data = datagen.text(700_000, 31)
    with b2mod.Encoder(level, 0) as e:
        out = e.encode(data, data.size).tobytes()
    assert out == orc.encode_stream(data, level, data.size)
    assert bz2.decompress(out) == data.tobytes()


@pytest.mark.parametrize("nsym", [1, 2, 3, 5, 9, 17, 33, 64, 65, 129, 255, 256])
def test_block_alphabet_sizes(enc9, nsym):
    """Every width of the packed round-0 sort keys (1 .. 8 bits per character) and both move-to-front kernels
    (<= 64 distinct bytes: four segments per warp; more: one): no long runs, so RLE1 adds no count bytes."""
    rng = np.random.default_rng(900 + nsym)
    alphabet = rng.permutation(256)[:nsym].astype(np.uint8)
    if nsym == 1:
        data = np.full(3000, alphabet[0], np.uint8)          # a single byte value: runs only (count bytes join the alphabet)
    else:
        idx = rng.integers(0, nsym, 40_000)
        idx[1:][idx[1:] == idx[:-1]] = (idx[1:][idx[1:] == idx[:-1]] + 1) % nsym     # avoid runs of four
        data = alphabet[idx]
    _cmp_block(enc9, data, "alphabet%d" % nsym)


def test_batch_of_mixed_alphabets_equals_single_calls(enc9):
    """The key width of a batch is set by its largest alphabet; results must not depend on the company."""
    rng = np.random.default_rng(77)
    entries = [datagen.text(60_000, 5), rng.integers(0, 2, 50_000).astype(np.uint8), datagen.random_bytes(30_000, 6),
               np.frombuffer(b"abc" * 9000, np.uint8), datagen.sparse_binary(70_000, 8)]
    outs = enc9.encode_batch(entries, "size")
    for e, o in zip(entries, outs):
        assert o == orc.encode_stream(e, 9, e.size)
