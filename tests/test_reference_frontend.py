"""Pins the method-string front end against the REFERENCE's own text: makeConfig (LibZPAQ.cs:388-1044: method string ->
ZPAQL config text + the nine arguments) and the part of compressBlock that expands a numeric method "LB,R,t" into its
x-method from the block size, the redundancy / type hints and, for levels 5+, an analysis of the data
(LibZPAQ.cs:125-141, 158-290), compiled where they lie by oracle/build_ref.py.  The oracle's front end
(oracle/frontend.py) must produce the same strings; the product's C++ front end is checked against the oracle's byte
for byte by tests/test_frontend_host.py.  What stays restated is the ZPAQL assembler (Compiler.cs: config text -> header
bytes), pinned by the three built-in bytecodes.  Skipped when the fragment is not available."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, frontend  # noqa: E402
from tools import synth  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build_frontend()
    if not path or not os.path.exists(path):
        pytest.skip("reference front-end fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    L.ref_make_config.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_int]
    L.ref_expand_method.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_char_p, C.c_int]
    return L


def _ref_config(L, method):
    args = (C.c_int * 9)()
    out = C.create_string_buffer(1 << 16)
    n = L.ref_make_config(method.encode(), args, out, 1 << 16)
    assert n >= 0, method
    return out.raw[:n].decode("latin-1"), list(args)


def _tokens(text):
    """The token stream the ZPAQL compiler sees (Compiler.cs:191-227): comments in (nested) parentheses and white space dropped,
    case folded."""
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch == "(":
            depth += 1
        if depth == 0 and not ch.isspace():
            cur += ch.lower()
        elif cur:
            out.append(cur)
            cur = ""
        if ch == ")" and depth > 0:
            depth -= 1
    if cur:
        out.append(cur)
    return out


def _ref_expand(L, method, data):
    out = C.create_string_buffer(4096)
    n = L.ref_expand_method(method.encode(), data, len(data), out, 4096)
    assert n >= 0, method
    return out.raw[:n].decode()


X_METHODS = ["0", "00,0", "x0,0c256,0,255,255", "x0,6,4,0,3,19", "x0,2,12,0,7,21,1c0,0,511i2m", "x0,5,4,3,3,19,1", "x0,7ci1", "x4,3ci1",
             "s0,0c0,0,255i2", "x0,0c1,0,255,255a24mm16ts19t0w2", "x5,3ci1", "x6,7ci1", "x6,1,4,0,3,24", "x8,5,4,0,3,24", "x3,4c0,0,255",
             "x0,0c0,1003,255c0,7c0,0,1300,255c0,0,1005,0,255c1000,3c200,0,511,300", "s4,4c0,0,255i1,2,3,4ms20,10,100t5,20",
             "x1,0w2,48,10,255,16,1a30,1,2", "x0,0ci1,1,1,1,2am", "x2,0w1i1c256ci1,1,1,1,1,1,2ac0,2,0,255i1c0,3,0,0,255i1c0,4,0,0,0,255i1mm16ts19t0",
             "x0,0c0,0,1015,255i1c0,16i1", "x4,1,5,0,3,25", "x4,2,8,0,4,22,2", "x0,0f", "x0,3ci1,2m8s", "x0,0ci2,3a16m12,30t4,10,50s8,40,200"]


@pytest.mark.parametrize("method", X_METHODS)
def test_make_config_text_and_args_match_reference(ref, method):
    want_text, want_args = _ref_config(ref, method)
    text, args = frontend.make_config(method)
    assert list(args) == want_args
    assert _tokens(text) == _tokens(want_text)         # the oracle's text carries no comments


NUMERIC = ["0", "1", "2", "3", "4", "5", "6", "9", "10,128,0", "11,50,0", "12,200,1", "14,255,3", "20,128,2", "21,20,0", "24,100,1", "30,128,1",
           "30,128,0", "30,10,2", "30,30,0", "33,200,2", "40,128,3", "41,200,3", "40,250,1", "40,2,0", "40,5,0", "40,10,1", "44,230,2", "50,128,0",
           "55,128,1", "60,128,3", "x0,0c256,0,255,255"]


@pytest.mark.parametrize("method", NUMERIC)
@pytest.mark.parametrize("size", [0, 1, 5000, 300000, 1044480, 5000000])
def test_numeric_method_expansion_matches_reference(ref, method, size):
    if size <= 300000:
        data = synth.blocks("mixed", 31, 1, size).tobytes() if size else b""
    else:                          # a record structure: the level 5+ analysis finds the periods
        rec = bytes(range(37)) + b"\x00" * 11
        data = (rec * (size // len(rec) + 1))[:size]
    want = _ref_expand(ref, method, data)
    assert frontend.expand_method(method, data) == want
    # and the expanded method goes through makeConfig identically
    if want[0] in "xs0":
        want_text, want_args = _ref_config(ref, want)
        text, args = frontend.make_config(want)
        assert (_tokens(text), list(args)) == (_tokens(want_text), want_args)
