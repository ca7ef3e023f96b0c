"""k_scatter2 / k_scatter3 (the radix pass of the BWT rotation sort, zip-ada_b200/csrc/b2_scatter2.cuh) on the host: the
kernel source is compiled with g++ against tests/emu/cuda_emu.h (one OS thread per CUDA thread, barriers for
__syncthreads / __syncwarp / the warp collectives) and must produce, for blocks of many shapes, the stable sort of
every block by the digit - with the blocks of the grid run one at a time (the look-back always finds its
predecessor finished) and three at a time in plain block order (it has to wait and walk back)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "emu_scatter")
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers (vector types) not found")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wno-attributes", "-I", cuda_inc, "-o", exe,
                    os.path.join(ROOT, "tests", "emu", "emu_scatter.cpp")], check=True)
    return exe


# (kernel MODE, blocks interleaved in the dispatch order, resident blocks, seed, shift step)
# modes 0..3: k_scatter2 (bit 0 match.any, bit 1 keys loaded early); 40..71: k_scatter3 (40 + bits: 2 = tiles of 2048 rows, 4 = rotation
# indices requested early, 8 = digit from the key registers, 16 = second early look; the prefetch distance alternates between 0 and 3
# inside the run)
@pytest.mark.parametrize("args", [(0, 128, 1, 1, 24), (0, 1, 3, 2, 24), (1, 2, 3, 3, 24), (2, 128, 2, 4, 24),
                                  (40, 64, 1, 6, 24), (52, 1, 3, 7, 24), (52, 64, 1, 9, 16), (70, 2, 4, 12, 24)])
def test_scatter_emulated(emu, args):
    r = subprocess.run([emu] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:]
