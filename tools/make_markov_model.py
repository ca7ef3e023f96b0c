#!/usr/bin/env python
"""Builds tests/golden/markov3_model.npz: the order-3 byte Markov model of SURVEY.md §8(d)'s text
generator, trained on the concatenation of /root/reference/doc/*.txt and zip_lib/*.ad? in sorted path
order.  Runs only where /root/reference exists (the build container); the model file is committed so
that tests and bench.py never read the reference at run time.

Stored: `pairs` = (context << 8 | next byte) as uint32, sorted; `counts` = occurrences of each pair.
"""
import glob
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def main():
    files = sorted(glob.glob(os.path.join(REF, "doc", "*.txt")) + glob.glob(os.path.join(REF, "zip_lib", "*.ad?")))
    assert files, "reference not found"
    data = b"".join(open(f, "rb").read() for f in files)
    a = np.frombuffer(data, np.uint8).astype(np.uint32)
    pair = (a[:-3] << 24) | (a[1:-2] << 16) | (a[2:-1] << 8) | a[3:]
    pairs, counts = np.unique(pair, return_counts=True)
    out = os.path.join(ROOT, "tests", "golden", "markov3_model.npz")
    np.savez_compressed(out, pairs=pairs.astype(np.uint32), counts=counts.astype(np.uint32))
    print("%d files, %d bytes, %d (context, next) pairs -> %s (%d bytes)" % (len(files), len(data), pairs.size, out, os.path.getsize(out)))


if __name__ == "__main__":
    main()
