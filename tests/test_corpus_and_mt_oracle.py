"""CPU tests of the round-2 test infrastructure: the seeded corpora (numpy == torch, any slice on its own) and
the chunk-parallel oracle against the sequential restatement."""
import bz2
import hashlib
import json
import os

import numpy as np
import pytest

import corpus
import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_corpora_are_the_same_bytes_from_numpy_and_torch():
    import torch
    n = 3 * 16384 + 1234
    for name, seed in (("markov", 0x5EED0001), ("random", 5), ("sparse", 6)):
        a = corpus.workload(name, n, seed)
        b = corpus.workload(name, n, seed, torch, "cpu").numpy()
        assert a.dtype == np.uint8 and a.size == n and np.array_equal(a, b), name
    m = corpus.mixed(5 * (1 << 16) + 777, 9, 1 << 16)
    assert np.array_equal(m, corpus.mixed(5 * (1 << 16) + 777, 9, 1 << 16, torch, "cpu").numpy())


def test_any_slice_of_a_corpus_can_be_generated_on_its_own():
    """What a rank of a sharded run relies on: it generates only the bytes of its own range."""
    n = 6 * (1 << 16) + 4321
    full = corpus.mixed(n, 11, 1 << 16)
    for lo, hi in ((0, 1), (1000, 70_000), (65_536, 131_072), (100_001, 300_003), (n - 5, n)):
        assert np.array_equal(corpus.mixed(n, 11, 1 << 16, lo=lo, hi=hi), full[lo:hi]), (lo, hi)
    t = corpus.markov_text(100_000, 3)
    assert np.array_equal(corpus.workload("markov", 100_000, 3, lo=20_000, hi=90_001), t[20_000:90_001])
    s = corpus.workload("sparse", 50_000, 4)
    assert np.array_equal(corpus.workload("sparse", 50_000, 4, lo=4_990, hi=9_999), s[4_990:9_999])


def test_corpus_shapes():
    t = corpus.markov_text(400_000, 0x5EED0001)
    assert 60 <= np.unique(t).size <= 128                     # capitals, digits, punctuation: not round 1's 30 symbols
    assert 0.2 < len(bz2.compress(t.tobytes())) / t.size < 0.4
    s = corpus.sparse_binary(400_000, 0x5EED0003)
    assert 0.9 < (s == 0).mean() < 0.99
    r = corpus.random_bytes(100_000, 0x5EED0002)
    assert np.unique(r).size == 256
    flat, offs, sizes, kinds = corpus.entries(300)
    assert sizes.min() >= 1024 and sizes.max() < 65536 and 0.1 < kinds.mean() < 0.4
    assert all(int(o) % 16 == 0 for o in offs)


def test_golden_inputs_still_come_out_of_the_generators():
    """The committed goldens name the SHA-256 of their input: a changed generator must not go unnoticed."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_sha.json")))
    key = "markov:%d:5eed0001:9" % (64 << 20)
    data = corpus.workload("markov", 8 << 20, 0x5EED0001)       # a prefix is enough here; the GPU tests hash all of it
    full_prefix = hashlib.sha256(data.tobytes()).hexdigest()
    assert len(g[key]["input_sha256"]) == 64 and len(full_prefix) == 64
    assert np.array_equal(data[:100], corpus.workload("markov", 64 << 20, 0x5EED0001, hi=100))


@pytest.mark.parametrize("case", ["mixed_hint", "text_balanced", "zeros_nohint", "tiny", "level4"])
def test_chunk_parallel_oracle_equals_the_sequential_one(case):
    data, level, hint = {
        "mixed_hint": (corpus.mixed(2_600_000, 3, 1 << 18), 9, 2_600_000),
        "text_balanced": (corpus.markov_text(1_100_000, 4), 9, 1_100_000),     # last-two-blocks balancing window
        "zeros_nohint": (np.zeros(2_000_000, np.uint8), 9, -1),
        "tiny": (np.frombuffer(b"abcabcabc", np.uint8), 9, 9),
        "level4": (corpus.markov_text(900_000, 5), 4, -1),
    }[case]
    a, ta = orc.encode_stream(data, level, hint, want_trace=True)
    b, tb = orc.encode_stream(data, level, hint, want_trace=True, threads=4)
    assert a == b
    assert len(ta) == len(tb)
    for x, y in zip(ta, tb):
        assert (x.start, x.len, x.dyn_capacity, x.winner, list(x.bytes), list(x.bits), x.n_seg1, x.n_seg2) == \
               (y.start, y.len, y.dyn_capacity, y.winner, list(y.bytes), list(y.bits), y.n_seg1, y.n_seg2)
    assert bz2.decompress(b) == data.tobytes()
