#!/usr/bin/env python3
"""Generate small golden archives with the CPU oracle (oracle/), one per method of the hot-path
scope table, over deterministic inputs from tools/synth.  The reference itself
ships no archive and cannot run, so these vectors pin the ORACLE (and through it
the CUDA path) against regressions; they are not outputs of ZPAQSharp.  Re-run after an
intentional oracle change:  python tests/golden/make_archives.py
"""
import base64, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po
from tools import synth

CASES = [
    # (name, kind, first_block, nbytes, how, arg)
    ("c1_cm_order2", "text", 3, 6000, "method", "x0,0c256,0,255,255"),
    ("c2a_mid", "mixed", 5, 5000, "level", 2),
    ("min", "text", 6, 5000, "level", 1),
    ("c4_max", "mixed", 7, 4000, "level", 3),
    ("c2b_lz77_sa", "text", 8, 6000, "method", "20"),
    ("c2c_lz77_cm", "mixed", 9, 6000, "method", "x0,2,12,0,7,21,1c0,0,511i2m"),
    ("c3_bwt", "text", 10, 6000, "method", "32,128,1"),
    ("m1_lz77_hash", "mixed", 11, 6000, "method", "1"),
    ("m4_cm", "text", 12, 5000, "method", "4"),
    ("m5_cm_many", "mixed", 13, 4000, "method", "5"),
    ("e8e9_cm", "mixed", 14, 6000, "method", "x0,4ci1"),
    ("e8e9_bwt", "mixed", 15, 6000, "method", "x0,7ci1"),
    ("stored", "mixed", 16, 3000, "method", "0"),
    ("empty", "text", 17, 0, "level", 2),
    ("one_byte", "text", 18, 1, "level", 2),
]


def make_input(kind, first, n):
    if n == 0:
        return b""
    return synth.blocks(kind, first, 1, max(n, 1)).tobytes()[:n]


def main():
    out = []
    for name, kind, first, n, how, arg in CASES:
        data = make_input(kind, first, n)
        if how == "level":
            arc = po.compress_block_level(data, arg)
        else:
            arc = po.compress_block(data, arg)
        back, st = po.decompress(arc)
        assert back == data, name
        out.append({"name": name, "kind": kind, "first_block": first, "nbytes": n, "how": how, "arg": arg,
                    "input_sha1": hashlib.sha1(data).hexdigest(), "archive_sha1": hashlib.sha1(arc).hexdigest(),
                    "archive_b64": base64.b64encode(arc).decode()})
        print(name, n, "->", len(arc))
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "oracle_archives.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
