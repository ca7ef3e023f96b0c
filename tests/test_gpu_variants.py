"""Every switchable variant of the CUDA path gives the oracle's bytes (`-m gpu`): the radix pass (B2GPU_SCATTER: round 1's
k_scatter, k_scatter2 with ballots / match.any, two or three CTAs per SM, keys loaded early or late, k_scatter3 with tiles
of 4096 / 2048 rows, with and without the L2 prefetch, the early rotation indices, the digit from the key registers and the
second early look at the predecessor), the package-merge
lists (B2GPU_PM: binary searches / merge path) and round 0 of the rotation sort in eight or seven passes (B2GPU_R0).
The knobs are read per call, so one handle serves all of them."""
import os

import numpy as np
import pytest

import datagen
import oracle_lib as orc
from test_gpu_parity import _cmp_block

pytestmark = pytest.mark.gpu
KNOBS = ("B2GPU_SCATTER", "B2GPU_PM", "B2GPU_R0")
VARIANTS = [{"B2GPU_SCATTER": s} for s in (1, 2, 22, 24, 3, 32, 34, 40, 41, 42, 43, 44, 45, 46, 47, 53, 61, 69, 71)] + [{"B2GPU_PM": 0}, {"B2GPU_PM": 1}, {"B2GPU_R0": 8}, {"B2GPU_R0": 7},
                                                                        {"B2GPU_SCATTER": 1, "B2GPU_PM": 0, "B2GPU_R0": 8}]


@pytest.fixture
def knobs():
    saved = {k: os.environ.get(k) for k in KNOBS}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


_ORACLE = {}


def _inputs():
    """The inputs and what the oracle makes of them, once per session."""
    if not _ORACLE:
        # a stream of text, random bytes (more than 128 distinct: the seven-pass round 0) and sparse data: several chunks,
        # blocks of many sizes, partial tiles
        data = datagen.mixed(2_600_000, 300_000, 21)
        _ORACLE["data"] = data
        _ORACLE["stream"] = orc.encode_stream(data, 9, data.size)
        # one block with all 256 byte values and long repeats (doubling rounds)
        _ORACLE["blk"] = np.concatenate([np.tile(datagen.random_bytes(7000, 22), 6), np.arange(256, dtype=np.uint8), datagen.text(90_000, 23)])
    return _ORACLE


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join("%s=%s" % (k[6:], x) for k, x in v.items()))
def test_variant_equals_oracle(enc9, knobs, variant):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in variant.items()})
    x = _inputs()
    out = enc9.encode(x["data"], x["data"].size).tobytes()
    assert out == x["stream"], variant
    _cmp_block(enc9, x["blk"], str(variant))          # every tap of the block
