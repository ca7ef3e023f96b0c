#!/usr/bin/env python
"""Per-stage device times of one encode for a given synthetic data kind (text / random / sparse / mixed)."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen

STAGES = ["cut_segment", "rle1", "bwt_sort", "mtf_rle2", "entropy_search", "pack", "concat", "copies"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="text", choices=["text", "random", "sparse", "mixed"])
    ap.add_argument("--size-mb", type=int, default=128)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--no-stage", action="store_true")
    a = ap.parse_args()
    b2 = importlib.import_module("zip-ada_b200")
    n = a.size_mb << 20
    data = {"text": datagen.text, "random": datagen.random_bytes, "sparse": datagen.sparse_binary,
            "mixed": lambda k: datagen.mixed(k, 16 << 20)}[a.kind](n)
    with b2.Encoder(a.level, 0) as enc:
        enc.encode(data, n)
        enc.reset_stats()
        t0 = time.perf_counter()
        out = enc.encode(data, n)
        dt = time.perf_counter() - t0
        st = enc.stats()
        res = {"kind": a.kind, "MiB": a.size_mb, "MBps_e2e_pageable": round(n / 1e6 / dt, 1), "ratio": round(out.size / n, 4),
               "chunks": int(st.chunks), "blocks": int(st.blocks), "block_bytes_per_input_byte": round(st.block_bytes / n, 3),
               "sort_rounds": int(st.sort_rounds), "kernel_launches": int(st.kernel_launches)}
        if not a.no_stage:
            enc.reset_stats(); enc.set_timing(2)
            enc.encode(data, n)
            res["stage_ms"] = dict(zip(STAGES, [round(x, 1) for x in enc.stats().stage_ms]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
