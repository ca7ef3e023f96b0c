"""b2_verify_stream: block-parallel BZip2 decode / verify on the device.  `-m gpu`."""
import bz2

import numpy as np
import pytest

import corpus
import datagen
import oracle_lib as orc

pytestmark = pytest.mark.gpu


CASES = {
    "markov_3M": lambda: corpus.markov_text(3_000_000, 71),
    "mixed_6M": lambda: corpus.mixed(6_000_000, 72, 1 << 19),
    "zeros_9p5M": lambda: np.zeros(9_500_000, np.uint8),
    "random_1M": lambda: datagen.random_bytes(1_000_000, 73),
    "runs": lambda: np.repeat(datagen.random_bytes(5000, 74, 0, 5), np.random.default_rng(75).integers(1, 700, 5000)),
    "tiny": lambda: np.frombuffer(b"hello hello hello", np.uint8),
    "one_byte": lambda: np.array([0], np.uint8),
}


@pytest.mark.parametrize("name", list(CASES))
def test_verify_own_streams(enc9, name):
    data = CASES[name]()
    out = enc9.encode(data, data.size)
    tr = enc9.trace()
    r = enc9.verify(out, data)
    assert r.ok == 1, (name, r.first_bad_block, r.first_bad_status, r.mismatch_at, hex(r.stored_stream_crc), hex(r.computed_stream_crc))
    assert r.decoded_bytes == data.size and r.level == 9
    assert r.stored_stream_crc == r.computed_stream_crc
    n_blocks = sum({0: 1, 1: 4, 2: max(1, t.n_seg1), 3: max(1, t.n_seg2)}[t.winner] for t in tr)
    assert r.blocks == n_blocks
    assert r.candidates >= r.blocks + 1
    assert enc9.verify(out).ok == 1                       # CRCs only


def test_verify_empty_stream(enc9):
    out = enc9.encode(np.zeros(0, np.uint8), 0)
    r = enc9.verify(out, np.zeros(0, np.uint8))
    assert r.ok == 1 and r.blocks == 1 and r.decoded_bytes == 0      # one empty block, as the reference writes it


@pytest.mark.parametrize("level", [1, 5, 9])
def test_verify_libbz2_streams(enc9, level):
    """Streams of an independent encoder (libbz2): the decoder is not tied to this encoder's choices."""
    data = corpus.mixed(2_500_000, 76, 1 << 18)
    s = np.frombuffer(bz2.compress(data.tobytes(), level), np.uint8)
    r = enc9.verify(s, data)
    assert r.ok == 1 and r.decoded_bytes == data.size and r.level == level, (r.first_bad_block, r.first_bad_status, r.mismatch_at)
    assert r.blocks >= 1 + (level < 9)


def test_verify_other_levels(b2mod):
    data = corpus.markov_text(1_500_000, 77)
    for level in (1, 4):
        with b2mod.Encoder(level, 0) as e:
            out = e.encode(data, data.size)
            r = e.verify(out, data)
            assert r.ok == 1 and r.level == level and r.decoded_bytes == data.size


def test_verify_detects_damage(enc9):
    data = corpus.markov_text(2_500_000, 78)
    out = enc9.encode(data, data.size).copy()
    assert enc9.verify(out, data).ok == 1
    # a flipped bit inside a block: its CRC (or its syntax) fails
    bad = out.copy(); bad[out.size // 2] ^= 0x10
    r = enc9.verify(bad, data)
    assert r.ok == 0 and r.first_bad_block >= 0
    # the stream CRC in the footer
    bad = out.copy(); bad[-3] ^= 0x01
    r = enc9.verify(bad)
    assert r.ok == 0 and r.first_bad_block == -1 and r.stored_stream_crc != r.computed_stream_crc
    # a truncated stream has no footer
    r = enc9.verify(out[:out.size - 20])
    assert r.ok == 0 and r.first_bad_status in (22, 12, 30, 7, 8, 9, 11)
    # trailing garbage
    r = enc9.verify(np.concatenate([out, np.zeros(3, np.uint8)]))
    assert r.ok == 0 and r.first_bad_status == 21
    # not the expected bytes
    other = data.copy(); other[1_234_567] ^= 0xFF
    r = enc9.verify(out, other)
    assert r.ok == 0 and r.mismatch_at == 1_234_567
    r = enc9.verify(out, data[:-5])
    assert r.ok == 0 and r.mismatch_at == data.size - 5
    # not a stream at all
    assert enc9.verify(np.frombuffer(b"PK\x03\x04 not bzip2 at all......", np.uint8)).ok == 0


def test_verify_large_stream_matches_libbz2(enc9):
    """48 MiB of the config-3 corpus: device verify and libbz2 agree, several waves of blocks."""
    import os
    data = corpus.mixed(48 << 20, 79, 1 << 22)
    out = enc9.encode(data, data.size)
    os.environ["B2GPU_VERIFY_WAVE"] = "40"
    try:
        r = enc9.verify(out, data)
    finally:
        del os.environ["B2GPU_VERIFY_WAVE"]
    assert r.ok == 1 and r.decoded_bytes == data.size
    assert bz2.decompress(out.tobytes()) == data.tobytes()
