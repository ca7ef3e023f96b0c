"""Archive side on the device (b2_zip_create, b2_zip_crc32) against the oracle's Zip.Create restatement
and against Python's zipfile.  `-m gpu`."""
import importlib
import io
import zipfile
import zlib

import numpy as np
import pytest

import datagen
import oracle_lib as orc

pytestmark = pytest.mark.gpu


def _entries():
    return [("a/text.txt", datagen.text(30_000, 5).tobytes()),
            ("rnd.bin", datagen.random_bytes(5_000, 6).tobytes()),
            ("empty", b""),
            ("dir\\sub\\back.txt", b"hello " * 40),
            ("one", b"x"),
            ("sparse.bin", datagen.sparse_binary(40_000, 7).tobytes()),
            ("big.txt", datagen.text(1_200_000, 8).tobytes()),                 # two chunks, balanced
            ("mixed.bin", datagen.mixed(300_000, 50_000, 9).tobytes())]


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 4095, 65_536, 65_537, 200_001, 5_000_000])
def test_zip_crc32_device(enc9, n):
    d = datagen.random_bytes(n, 100 + n % 11)
    assert enc9.zip_crc32(d) == zlib.crc32(d.tobytes())


def test_archive_equals_oracle_and_reads_back(enc9):
    ents = _entries()
    t = [16789 * 65536 + 7 * i for i in range(len(ents))]
    fl = [0, 2, 0, 1, 0, 0, 3, 0]
    arc, info = enc9.zip_create(ents, dos_times=t, flags=fl, want_info=True)
    ref, methods = orc.zip_create(ents, 9, dos_times=t, flags=fl)
    assert arc.tobytes() == ref
    assert [i.zip_type for i in info] == methods
    z = zipfile.ZipFile(io.BytesIO(arc.tobytes()))
    assert z.testzip() is None
    for (name, data), zi, inf in zip(ents, z.infolist(), info):
        assert z.read(zi) == data
        assert inf.crc32 == zlib.crc32(data) and inf.compressed_size == zi.compress_size
        assert inf.local_header_offset == zi.header_offset


@pytest.mark.parametrize("level", [1, 4])
def test_archive_levels(level):
    b2 = importlib.import_module("zip-ada_b200")
    ents = [("t.txt", datagen.text(250_000, 21).tobytes()), ("r", datagen.random_bytes(3_000, 22).tobytes())]
    with b2.Encoder(level, 0) as enc:
        arc = enc.zip_create(ents)
    assert arc.tobytes() == orc.zip_create(ents, level)[0]


def test_empty_archive_and_duplicates(enc9):
    b2 = importlib.import_module("zip-ada_b200")
    assert enc9.zip_create([]).tobytes() == orc.zip_create([], 9)[0]
    ents = [("same", b"abc" * 50), ("other", b"x"), ("same", b"def" * 50)]
    assert enc9.zip_create(ents, duplicates=0).tobytes() == orc.zip_create(ents, 9)[0]     # admit_duplicates
    with pytest.raises(b2.B2Error, match="Duplicate_name"):
        enc9.zip_create(ents, duplicates=1)                                                 # error_on_duplicate


def test_many_small_entries_zip64_end_records(enc9):
    n = 65_535
    ents = [("e%05d" % i, b"") for i in range(n - 3)] + [("t1", b"abc" * 400), ("t2", datagen.text(9000, 3).tobytes()), ("t3", b"z")]
    arc = enc9.zip_create(ents).tobytes()
    assert arc[-42:-38] == b"PK\x06\x07" and arc[-98:-94] == b"PK\x06\x06"
    z = zipfile.ZipFile(io.BytesIO(arc))
    assert len(z.infolist()) == n and z.read("t2") == ents[-2][1] and z.read("t1") == ents[-3][1]
    # the same central directory as the oracle's for a prefix that the CPU finishes quickly
    small = ents[-2000:]
    assert enc9.zip_create(small).tobytes() == orc.zip_create(small, 9)[0]


def test_archive_of_many_entries_reads_back(enc9):
    rng = np.random.default_rng(77)
    pool_t, pool_r = datagen.text(3_000_000, 71), datagen.random_bytes(500_000, 72)
    ents = []
    for i in range(1500):
        n = int(2 ** rng.uniform(6, 16))
        pool = pool_r if i % 5 == 0 else pool_t
        o = int(rng.integers(0, pool.size - n))
        ents.append(("dir%02d/file%04d.dat" % (i % 17, i), pool[o:o + n].tobytes()))
    arc, info = enc9.zip_create(ents, want_info=True)
    z = zipfile.ZipFile(io.BytesIO(arc.tobytes()))
    assert z.testzip() is None
    for (name, data), zi in zip(ents, z.infolist()):
        assert zi.filename == name and z.read(zi) == data
    assert {i.zip_type for i in info} == {0, 12}
    # a sample of entries against the oracle's stream for that entry
    for k in (0, 7, 500, 1499):
        data = ents[k][1]
        s = orc.encode_stream(data, 9, len(data))
        if info[k].zip_type == 12:
            off = info[k].local_header_offset + 30 + len(ents[k][0])
            assert arc[off:off + len(s)].tobytes() == s
        else:
            assert len(s) >= len(data)
