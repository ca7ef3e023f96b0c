"""The package-merge of one warp (length-limited code lengths, zip-ada_b200/csrc/b2_pm.cuh; reference
huffman-encoding-length_limited_coding.adb:46-280) on the host: both device versions (lists merged by binary
searches / by a merge path) run under tests/emu/cuda_emu.h and must give the code lengths of a plain sequential
package-merge, for alphabets of 2..260 symbols, ties everywhere, length limits that bind."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_merge_emulated(tmp_path):
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers (vector types) not found")
    exe = str(tmp_path / "emu_pm")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wno-attributes", "-I", cuda_inc, "-o", exe,
                    os.path.join(ROOT, "tests", "emu", "emu_pm.cpp")], check=True)
    r = subprocess.run([exe, "400", "11"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:]
