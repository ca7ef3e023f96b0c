#!/usr/bin/env python
"""Config-5 shape (BASELINE.json configs[4]): many small archive entries (sizes log-uniform in
1..64 KiB, text and incompressible), every entry its own BZh9 stream, encoded with ONE
b2_encode_batch call.  Prints entries/s and uncompressed MB/s; checks a sample against the oracle."""
import argparse
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
    ap.add_argument("--entries", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    b2 = importlib.import_module("zip-ada_b200")
    rng = np.random.default_rng(0x5EED0005)
    sizes = (2 ** rng.uniform(10, 16, a.entries)).astype(np.int64)
    total = int(sizes.sum())
    pool_t = datagen.text(8_000_000, 51)
    pool_r = datagen.random_bytes(2_000_000, 52)
    entries = []
    for i, n in enumerate(sizes):
        pool = pool_r if i % 4 == 0 else pool_t
        o = int(rng.integers(0, pool.size - n))
        entries.append(pool[o:o + n])
    with b2.Encoder(b2.block_900k, 0) as enc:
        outs = enc.encode_batch(entries, "size")          # warm-up (allocations)
        t0 = time.perf_counter()
        for _ in range(a.reps):
            outs = enc.encode_batch(entries, "size")
        dt = (time.perf_counter() - t0) / a.reps
    ok = all(outs[i] == orc.encode_stream(entries[i], 9, entries[i].size) for i in range(0, a.entries, max(1, a.entries // 40)))
    stored = sum(1 for e, o in zip(entries, outs) if len(o) >= e.size)
    print(json.dumps({"workload": "%d entries, log-uniform 1-64 KiB, 3/4 text 1/4 random, one b2_encode_batch call" % a.entries,
                      "entries_per_s": round(a.entries / dt, 1), "MBps": round(total / 1e6 / dt, 1), "seconds": round(dt, 3),
                      "input_bytes": total, "output_bytes": int(sum(len(o) for o in outs)),
                      "entries_not_smaller_than_input": stored, "sample_equals_oracle": bool(ok)}))


if __name__ == "__main__":
    main()
