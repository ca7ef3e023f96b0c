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
import corpus

STAGES = ["cut_segment", "rle1", "bwt_sort", "mtf_rle2", "entropy_search", "pack", "concat", "copies"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="text", choices=["text", "random", "sparse", "mixed", "markov", "zipf"])
    ap.add_argument("--size-mb", type=int, default=128)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--no-stage", action="store_true")
    ap.add_argument("--cpu-gen", action="store_true", help="generate the input with numpy (no torch kernels: for runs under ncu)")
    a = ap.parse_args()
    b2 = importlib.import_module("zip-ada_b200")
    n = a.size_mb << 20
    if a.kind in ("text", "zipf"):
        data = datagen.text(n)
    else:
        seed = {"markov": 0x5EED0001, "mixed": 0x5EED0004, "random": 0x5EED0002, "sparse": 0x5EED0003}[a.kind]
        if a.cpu_gen:
            data = corpus.workload(a.kind, n, seed)
        else:
            import torch
            data = corpus.workload(a.kind, n, seed, torch, "cuda").cpu().numpy()
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
