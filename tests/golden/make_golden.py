"""Regenerates tests/golden/oracle_vectors.json from the oracle.  The reference (Ada) cannot run in
this image, so these are regression vectors of the oracle, not reference outputs; if a GNAT
toolchain ever becomes available, run the reference's `zipada -eb3` / `bzip2_enc` over the same
inputs (tests/make_inputs.py) and compare the SHA-256 values."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import make_inputs
import oracle_lib as orc

out = {}
for name, fn in make_inputs.CASES.items():
    data, level, hint = fn()
    s = orc.encode_stream(data, level, hint)
    out[name] = {"len": len(s), "sha256": hashlib.sha256(s).hexdigest(), "level": level, "size_hint": hint, "input_len": int(data.size),
                 "input_sha256": hashlib.sha256(data.tobytes()).hexdigest()}
json.dump(out, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
