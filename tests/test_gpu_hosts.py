"""The C++ and Python mirrors of the generic Encode, driven through the C ABI.  `-m gpu`."""
import os
import subprocess

import numpy as np
import pytest

import datagen
import oracle_lib as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_b2enc(tmp_path):
    exe = tmp_path / "b2enc"
    pkg = os.path.join(ROOT, "zip-ada_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(pkg, "host", "b2enc.cpp"), "-L" + pkg, "-lb2gpu",
                           "-Wl,-rpath," + pkg, "-o", str(exe)])
    data = datagen.mixed(700_000, 100_000, 41)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bz2"
    fin.write_bytes(data.tobytes())
    subprocess.check_call([str(exe), str(fin), str(fout), "-3"])
    # bzip2_enc passes no size hint (extras/bzip2_enc.adb:51-55)
    assert fout.read_bytes() == orc.encode_stream(data, 9, -1)


def test_python_callbacks_mirror(enc9):
    data = datagen.text(50_000, 42).tobytes()
    pos = [0]
    out = bytearray()

    def read_byte():
        b = data[pos[0]]
        pos[0] += 1
        return b

    enc9.encode_callbacks(read_byte, lambda: pos[0] < len(data), out.append, len(data))
    assert bytes(out) == orc.encode_stream(data, 9, len(data))


def test_two_handles_are_independent(b2mod):
    d1, d2 = datagen.text(300_000, 43), datagen.random_bytes(200_000, 44)
    with b2mod.Encoder(9, 0) as a, b2mod.Encoder(4, 0) as b:
        o1 = a.encode(d1, d1.size).tobytes()
        o2 = b.encode(d2, d2.size).tobytes()
        o1b = a.encode(d1, d1.size).tobytes()
    assert o1 == o1b == orc.encode_stream(d1, 9, d1.size)
    assert o2 == orc.encode_stream(d2, 4, d2.size)


def test_small_batches_and_pipeline_give_identical_bytes(b2mod, monkeypatch):
    data = datagen.mixed(6_000_000, 700_000, 45)
    ref = None
    for pos, pipe in (("4000000", "1"), ("4000000", "2"), ("600000000", "1")):
        monkeypatch.setenv("B2GPU_BATCH_POSITIONS", pos)
        monkeypatch.setenv("B2GPU_PIPELINE", pipe)
        with b2mod.Encoder(9, 0) as e:
            out = e.encode(data, data.size).tobytes()
        ref = ref or out
        assert out == ref
    assert ref == orc.encode_stream(data, 9, data.size)


def test_batch_of_entries_equals_one_stream_each(enc9):
    # config 5 shape (zip_with_many_files): many small entries, sizes log-uniform in 1 B .. 64 KiB,
    # text and incompressible, plus the edge sizes; every entry is its own BZh9 stream
    rng = np.random.default_rng(46)
    entries = [np.zeros(0, np.uint8), np.array([7], np.uint8), np.frombuffer(b"abc", np.uint8)]
    for i in range(120):
        n = int(2 ** rng.uniform(0, 16))
        entries.append(datagen.text(n, 1000 + i) if i % 3 else datagen.random_bytes(n, 1000 + i))
    entries.append(datagen.mixed(1_300_000, 200_000, 47))        # one entry that spans two chunks
    entries.append(datagen.text(1_000_000, 48))                  # balancing window when the hint is the size
    outs = enc9.encode_batch(entries, "size")
    for e, o in zip(entries, outs):
        assert o == orc.encode_stream(e, 9, e.size), e.size
    outs2 = enc9.encode_batch(entries[:40], None)
    for e, o in zip(entries[:40], outs2):
        assert o == orc.encode_stream(e, 9, -1), e.size


def test_cpp_zip_create_mirror(tmp_path):
    """host/zip_create.hpp (Create_Archive / Add_File / Finish) through host/b2zip.cpp, against the oracle archive."""
    import io
    import zipfile
    exe = tmp_path / "b2zip"
    pkg = os.path.join(ROOT, "zip-ada_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(pkg, "host", "b2zip.cpp"), "-L" + pkg, "-lb2gpu",
                           "-Wl,-rpath," + pkg, "-o", str(exe)])
    files = {"t.txt": datagen.text(120_000, 81).tobytes(), "r.bin": datagen.random_bytes(4_000, 82).tobytes(), "e": b""}
    for k, v in files.items():
        (tmp_path / k).write_bytes(v)
    subprocess.check_call([str(exe), "-eb3", "out.zip", "t.txt", "r.bin", "e"], cwd=str(tmp_path))
    arc = (tmp_path / "out.zip").read_bytes()
    assert arc == orc.zip_create(list(files.items()), 9)[0]
    z = zipfile.ZipFile(io.BytesIO(arc))
    assert z.testzip() is None and z.read("t.txt") == files["t.txt"]


def test_cpp_encode_loop_pools_the_handle_and_survives_callback_exceptions(tmp_path):
    """The Ada body's call sequence (one Encode per archive entry, exceptions out of Write_Byte / Read_Byte in
    the middle of an entry) through the C++ mirror: one pooled handle serves every entry."""
    exe = tmp_path / "loop"
    pkg = os.path.join(ROOT, "zip-ada_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(pkg, "host", "test_encode_loop.cpp"), "-L" + pkg, "-lb2gpu",
                           "-Wl,-rpath," + pkg, "-o", str(exe)])
    entries = [datagen.text(200_000, 61), datagen.random_bytes(50_000, 62),      # entry 1: incompressible -> Compression_inefficient
               datagen.text(100_000, 63),                                       # entry 2: User_abort half way through Read_Byte
               datagen.mixed(1_200_000, 150_000, 64), datagen.text(3_000, 65)]
    names = []
    for k, e in enumerate(entries):
        p = tmp_path / ("in%d.bin" % k)
        p.write_bytes(e.tobytes())
        names.append(str(p))
    res = subprocess.run([str(exe), str(tmp_path)] + names, check=True, stdout=subprocess.PIPE, text=True).stdout
    assert res.split() == ["created=1", "ok=3", "inefficient=1", "aborted=1"], res
    for k in (0, 3, 4):
        assert (tmp_path / ("%d.bz2" % k)).read_bytes() == orc.encode_stream(entries[k], 9, entries[k].size)
