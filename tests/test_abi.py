"""The C-ABI library loads and exports every symbol include/b2gpu.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "b2gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z_0-9]+)\s*\(", h)))


def test_header_declares_expected_entry_points(b2mod):
    assert set(b2mod.EXPORTS) == set(_declared())


def test_library_exports_every_declared_symbol(b2mod):
    lib = b2mod.lib()
    for name in _declared():
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu(b2mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b2mod.B2Error) as ei:
        b2mod.Encoder(b2mod.block_900k, 0)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "zip-ada_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp", ".hpp", ".adb", ".ads")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "b2oracle" not in src and "oracle_lib" not in src, (dp, f)
