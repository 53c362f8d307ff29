"""Pins the block FRAMING -- tag, block header, segment header, PCOMP preamble, end-of-segment marker, SHA-1 trailer, end of
block, and the stored (component-free) mode -- against the REFERENCE's own text: Compressor.writeTag / startBlock /
startSegment / postProcess / compress / endSegment / endBlock (Compressor.cs:27-99, 133-249, 294-299), ZPAQL.read / write
(ZPAQL.cs:112-179) and Encoder.init / compress / encode (Encoder.cs:26-103), compiled where they lie by
oracle/build_ref.py on top of the reference predictor fragment and driven in the order of LibZPAQ.compressBlock
(LibZPAQ.cs:296-325).  The whole archive block must equal the oracle's, byte for byte.  The ZPAQL *compiler*
(Compiler.cs, config text -> header bytes) is not part of this fragment; startBlock(level) covers the built-in headers.
Skipped when the fragment is not available."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, frontend, pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build_compressor()
    if not path or not os.path.exists(path):
        pytest.skip("reference Compressor fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
    tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
            np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
    L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
    L.ref_predictor_tables(*[t.ctypes.data for t in tabs])
    L.ref_compress_block.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_ulonglong,
                                     C.c_char_p, C.c_int, C.c_void_p, C.c_ulonglong]
    L.ref_compress_block.restype = C.c_longlong
    return L


def _ref_block(L, level, hdr, pcomp, filename, comment, payload, sha1, with_tag=True):
    cap = len(payload) * 2 + 70000
    out = C.create_string_buffer(cap)
    n = L.ref_compress_block(level, hdr, pcomp if pcomp else None, len(pcomp), filename, comment, payload, len(payload), sha1,
                             1 if with_tag else 0, out, cap)
    assert 0 <= n <= cap
    return out.raw[:n]


def _data(seed, n):
    return synth.blocks("mixed", seed, 1, n).tobytes()


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("n", [0, 1, 30000])
def test_builtin_levels_whole_block_matches_reference_compressor(ref, level, n):
    # startBlock(int level): the reference's own model table picks the header (Compressor.cs:45-83)
    data = _data(500 + level, n)
    want = po.compress_block_level(data, level, filename="dir/name.bin", comment=None)
    got = _ref_block(ref, level, None, b"", b"dir/name.bin", str(len(data)).encode(), data, po.sha1(data))
    assert got == want


METHODS = ["0", "x0,0c256,0,255,255", "1", "2", "x4,1,4,0,3,24c0,0,511", "x4,3ci1", "x4,4c0,0,255", "x4,5,12,0,3,20,1c0,0,511i2",
           "x4,7ci1,1m", "s4,0,0,255i1,2ms20"]


@pytest.mark.parametrize("method", METHODS)
def test_methods_whole_block_matches_reference_compressor(ref, method):
    # startBlock(hcomp) + postProcess(pcomp, len): stored mode, plain CM, LZ77 (hash / suffix array, bit-packed / byte),
    # BWT, E8E9 -- the payload is the oracle's pre-processed stream (pinned against LZBuffer.cs by test_reference_lzbuffer)
    data = _data(77, 20000) + b"\xe8\x10\x00\x00\x00" * 50 + _data(78, 5000)
    plan = frontend.plan_block(method, data)
    want = po.compress_block(data, method, filename=None, comment="a comment")
    payload = po.preprocess(data, plan["args"]) if plan["pcomp"] else data
    got = _ref_block(ref, 0, bytes(plan["hdr"]), bytes(plan["pcomp"]), None, (plan["comment"] + " a comment").encode(), payload, po.sha1(data))
    assert got == want


def test_no_checksum_and_no_tag(ref):
    data = _data(9, 3000)
    hdr, _ = frontend.builtin_model(1)
    want = po.compress_block_level(data, 1, filename=None, comment="", dosha1=False, with_tag=False)
    got = _ref_block(ref, 0, bytes(hdr), b"", None, b"", data, None, with_tag=False)
    assert got == want and got[-2:] == b"\xfe\xff" and got[:3] == b"zPQ"


def test_stored_mode_chunks_of_64k(ref):
    # component-free model: Encoder.compress buffers 1 << 16 bytes per length-prefixed chunk (Encoder.cs:59-72)
    data = _data(11, 200000)
    plan = frontend.plan_block("0", data)
    assert bytes(plan["hdr"])[6] == 0
    want = po.compress_block(data, "0")
    got = _ref_block(ref, 0, bytes(plan["hdr"]), b"", None, plan["comment"].encode(), data, po.sha1(data))
    assert got == want


GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_archives.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_committed_golden_archives_are_what_the_reference_text_writes(ref, case):
    # tests/golden/oracle_archives.json (made by tests/golden/make_archives.py with the oracle) is what the GPU tests decode
    # and reproduce on the box, where /root/reference does not exist: every fixture must be exactly what the reference's
    # Compressor text writes for that input (the model header from the front end, the payload pre-processed as
    # LibZPAQ.compressBlock does before it hands the data over).
    import base64
    n = case["nbytes"]
    data = synth.blocks(case["kind"], case["first_block"], 1, max(n, 1)).tobytes()[:n] if n else b""
    golden = base64.b64decode(case["archive_b64"])
    if case["how"] == "level":
        got = _ref_block(ref, case["arg"], None, b"", None, str(len(data)).encode(), data, po.sha1(data))
    else:
        plan = frontend.plan_block(case["arg"], data)
        payload = po.preprocess(data, plan["args"]) if plan["pcomp"] else data
        got = _ref_block(ref, 0, bytes(plan["hdr"]), bytes(plan["pcomp"]), None, plan["comment"].encode(), payload, po.sha1(data))
    assert got == golden
