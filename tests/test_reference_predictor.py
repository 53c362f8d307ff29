"""Pins the bit predictor -- the component formulas of CONS/CM/ICM/MATCH/AVG/MIX2/MIX/ISSE/SSE -- against the REFERENCE's
own text: the bodies of Predictor.init, predict0, update0 and find (Predictor.cs:39-172, 245-350, 353-475, 550-567) are
compiled where they lie by oracle/build_ref.py on top of the reference's ZPAQL interpreter (ZPAQL.cs:1028-1251), with the
static tables taken from the reference's literals (tests/golden/reference_kat.json).  Five one-line helpers whose C# text
is wrong (train, squash, stretch, clamp2k, clamp512k) are supplied by the harness as the reference's JIT
comments state them.  The probability handed to the arithmetic coder must agree with the oracle's for EVERY bit.
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
    path = build_ref.build_predictor()
    if not path or not os.path.exists(path):
        pytest.skip("reference predictor fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
    tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
            np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
    L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
    L.ref_predictor_tables(*[t.ctypes.data for t in tabs])
    L.ref_predict_trace.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_void_p]
    L.ref_predict_trace.restype = C.c_longlong
    return L


MODELS = [("level", 1), ("level", 2), ("level", 3),
          ("method", "x0,0c256,0,255,255"),                                   # CM
          ("method", "s0,0c0,0,255i2"),                                       # ICM + ISSE chain
          ("method", "x0,0ci1,1,1,1,2am"),                                    # ICM, ISSE x5, MATCH, MIX
          ("method", "x0,0c1,0,255,255a24mm16ts19t0w2"),                      # CM, MATCH, MIX x2, MIX2, SSE, word model
          ("method", "x0,0c0,1003,255c0,7c0,0,1300,255c200,0,511,300a24,1,1m12,20s9,20,100t3"),   # sparse / periodic contexts, SSE
          ("method", "x0,0w2,48,10,255,16,1a30,1,2"),                         # word model, AVG-free
          ("method", "s4,4c0,0,255i1,2,3,4ms20,10,100t5,20")]                 # E8E9 config: CM-free chain + MIX + SSE + MIX2


@pytest.mark.parametrize("what", MODELS, ids=[str(w[1]) for w in MODELS])
def test_every_bit_probability_matches_reference_predictor(ref, what):
    if what[0] == "level":
        hdr, _ = frontend.builtin_model(what[1])
    else:
        text, args = frontend.make_config(what[1])
        hdr, _, _ = frontend.compile_config(text, args)
    hdr = bytes(hdr)
    data = (synth.blocks("mixed", 321, 1, 40000).tobytes() + synth.blocks("text", 322, 1, 20000).tobytes()
            + b"\x00" * 300 + bytes(range(256)) * 3 + b"abcabcabcabc" * 200)
    want = np.zeros(len(data) * 8, dtype=np.uint16)
    n = ref.ref_predict_trace(hdr, len(hdr), data, len(data), want.ctypes.data)
    assert n == len(data) * 8
    got = po.predict_trace(hdr, data)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, "first differing bit %d of %d: oracle %d, reference %d" % (bad[0], n, got[bad[0]], want[bad[0]])


def _coded_payload(archive: bytes, hdr: bytes, filename: bytes = b"", comment: bytes = b"") -> bytes:
    """The coded bytes of a one-segment archive block: between the segment header and `00 00 00 00 FD sha1 FF`."""
    prefix = 13 + 5 + len(hdr) + 1 + len(filename) + 1 + len(comment) + 1 + 1
    assert archive[-1] == 255 and archive[-22] == 253 and archive[-26:-22] == b"\x00\x00\x00\x00"
    return archive[prefix:-26]


@pytest.mark.parametrize("level", [1, 2, 3])
def test_coded_bytes_match_reference_predictor_and_coder(ref, level):
    # reference predictor text x reference Encoder.encode text == the coded payload of the oracle's archive block
    ref.ref_code_block.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong]
    ref.ref_code_block.restype = C.c_longlong
    hdr, _ = frontend.builtin_model(level)
    hdr = bytes(hdr)
    data = synth.blocks("mixed", 400 + level, 1, 50000).tobytes()
    arc = po.compress_block_level(data, level)
    want = _coded_payload(arc, hdr, b"", str(len(data)).encode())
    cap = len(data) * 2 + 4096
    out = C.create_string_buffer(cap)
    n = ref.ref_code_block(hdr, len(hdr), b"\x00", 1, data, len(data), out, cap)      # preamble 0: no PCOMP (Compressor.cs:188)
    assert n == len(want) and out.raw[:n] == want
