#!/usr/bin/env python
"""Pins the oracle against the real Zip-Ada encoder — to be run on a machine that HAS the reference
built (GNAT + `gprbuild -P zipada -XZip_Build_Mode=Fast`); nothing in this repository's image can build it,
so tests/golden/oracle_vectors.json are regression vectors of the oracle until this script has been run.

  python tools/pin_against_reference.py --zipada /path/to/zipada --bzip2-enc /path/to/bzip2_enc

For every named seeded input of tests/make_inputs.py:
  * size_hint = -1  -> `bzip2_enc <in> <out> -<1|4|9>`  (extras/bzip2_enc.adb passes no hint);
  * size_hint = n   -> `zipada -eb<1|2|3> <zip> <in>`    (Zip.Compress.BZip2_E passes the size), and the entry's
    raw payload is read from the archive; an entry the reference decided to store cannot be compared.
The SHA-256 of the reference's bytes is compared with the golden vector.  The whole archive written by
zipada is not compared: it carries the file's time stamp (see tests/test_zip_oracle.py for the archive side).
"""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_inputs

LEVEL_DIGIT = {1: "1", 4: "4", 9: "9"}
LEVEL_METHOD = {1: "-eb1", 4: "-eb2", 9: "-eb3"}


def payload_of_single_entry(zip_bytes):
    sig, ver, flag, method, dostime, crc, csize, usize, nlen, xlen = struct.unpack("<IHHHIIIIHH", zip_bytes[:30])
    assert sig == 0x04034B50, "not a local header"
    off = 30 + nlen + xlen
    return method, zip_bytes[off:off + csize]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zipada")
    ap.add_argument("--bzip2-enc")
    a = ap.parse_args()
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))
    bad = 0
    compared = 0
    with tempfile.TemporaryDirectory() as tmp:
        for name, fn in make_inputs.CASES.items():
            data, level, hint = fn()
            fin = os.path.join(tmp, name + ".bin")
            open(fin, "wb").write(data.tobytes())
            got = None
            if hint < 0:
                if not a.bzip2_enc:
                    print("%-28s skipped (no --bzip2-enc)" % name)
                    continue
                fout = os.path.join(tmp, name + ".bz2")
                subprocess.check_call([a.bzip2_enc, fin, fout, "-" + LEVEL_DIGIT[level]])
                got = open(fout, "rb").read()
            else:
                if not a.zipada:
                    print("%-28s skipped (no --zipada)" % name)
                    continue
                fzip = os.path.join(tmp, name + ".zip")
                subprocess.check_call([a.zipada, LEVEL_METHOD[level], fzip, fin], stdout=subprocess.DEVNULL)
                method, got = payload_of_single_entry(open(fzip, "rb").read())
                if method != 12:
                    print("%-28s stored by the reference (method %d): nothing to compare" % (name, method))
                    continue
            ok = hashlib.sha256(got).hexdigest() == golden[name]["sha256"]
            bad += not ok
            compared += 1
            print("%-28s %s  (%d bytes, oracle %d)" % (name, "IDENTICAL" if ok else "DIFFERENT", len(got), golden[name]["len"]))
    if compared == 0:
        print("nothing compared: give --zipada and / or --bzip2-enc")
        return 2
    print("oracle pinned on %d case(s)" % compared if bad == 0 else "%d case(s) differ: see SURVEY.md 8c for the two places that depend on the GNAT runtime" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
