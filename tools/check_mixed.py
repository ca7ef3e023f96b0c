#!/usr/bin/env python
"""Config-3 shape (BASELINE.json configs[2]): mixed corpus in 16 MiB stripes cycling {text, random bytes,
sparse binary}; encodes it on the GPU, decodes with libbz2 and compares, reports throughput and the
per-chunk winners.  Also checks a few chunk-sized pieces against the oracle."""
import argparse
import bz2
import collections
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen
import oracle_lib as orc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mb", type=int, default=512)
    ap.add_argument("--stripe-mb", type=int, default=16)
    a = ap.parse_args()
    b2 = importlib.import_module("zip-ada_b200")
    n = a.size_mb << 20
    data = datagen.mixed(n, a.stripe_mb << 20, 0x5EED0003)
    with b2.Encoder(9, 0) as enc:
        out = enc.encode(data, n)
        t0 = time.perf_counter()
        out = enc.encode(data, n)
        dt = time.perf_counter() - t0
        tr = enc.trace()
        st = enc.stats()
        enc.reset_stats(); enc.set_timing(2)
        enc.encode(data, n)
        st2 = enc.stats()
        stage = dict(zip(["cut_segment", "rle1", "bwt_sort", "mtf_rle2", "entropy_search", "pack", "concat", "copies"], [round(x, 1) for x in st2.stage_ms]))
        enc.set_timing(0)
        # three chunk-sized pieces straddling stripe boundaries, against the oracle
        ok_pieces = True
        for k in (1, 2, 3):
            lo = k * (a.stripe_mb << 20) - 600_000
            piece = data[lo:lo + 1_500_000]
            ok_pieces &= enc.encode(piece, piece.size).tobytes() == orc.encode_stream(piece, 9, piece.size)
    dec = bz2.decompress(out.tobytes())
    winners = collections.Counter(t.winner for t in tr)
    print(json.dumps({"workload": "%d MiB mixed corpus, %d MiB stripes {text, random, sparse}" % (a.size_mb, a.stripe_mb),
                      "MBps_e2e_pageable": round(n / 1e6 / dt, 1), "seconds": round(dt, 3), "ratio": round(out.size / n, 4),
                      "decodes_with_libbz2": bool(dec == data.tobytes()), "pieces_equal_oracle": bool(ok_pieces),
                      "chunks": len(tr), "winners": {str(k): v for k, v in sorted(winners.items())},
                      "sorted_bytes_per_input_byte": round(st.block_bytes / max(1, st.input_bytes), 3),
                      "sort_rounds": int(st.sort_rounds), "stage_ms": stage}))


if __name__ == "__main__":
    main()
