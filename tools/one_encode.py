#!/usr/bin/env python
"""One Encode of a text stream and nothing else (for `ncu -k regex:^k_` launch lists: no warm-up; the corpus comes from
torch on the GPU, whose kernels the filter leaves out).
usage: one_encode.py <MiB>"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
b2 = importlib.import_module("zip-ada_b200")
import torch
data = corpus.workload("markov", n, 0x5EED0001, torch, "cuda").cpu().numpy()
torch.cuda.synchronize()
torch.cuda.empty_cache()
with b2.Encoder(b2.block_900k, 0) as enc:
    out = enc.encode(data, n)
print(n, out.size)
