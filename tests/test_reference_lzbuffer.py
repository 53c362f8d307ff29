"""Pins the pre-processing path against the REFERENCE's own code: the body of /root/reference/ZPAQSharp/LZBuffer.cs
(LZ77 with the hash and the suffix-array matcher, both code formats, and the BWT emit, still C++ text) and `e8e9` from
LibZPAQ.cs:371-384 are compiled where they lie by oracle/build_ref.py, around a harness for the three classes the
reference lacks, into oracle/_ref/ (git-ignored, travels to the GPU box).  The oracle's pre-processor -- which the GPU
tests pin the device kernels to -- must produce the same bytes.  Skipped when the fragment is not available."""
import ctypes as C
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
    path = build_ref.build_lzbuffer()
    if not path or not os.path.exists(path):
        pytest.skip("reference LZBuffer fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    L.ref_lzbuffer.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_int), C.c_void_p, C.c_ulonglong]
    L.ref_lzbuffer.restype = C.c_longlong
    L.ref_e8e9.argtypes = [C.c_void_p, C.c_int]
    return L


def ref_preprocess(L, data: bytes, args) -> bytes:
    a = (C.c_int * 9)(*[int(x) for x in args[:9]])
    cap = len(data) * 2 + 4096
    out = C.create_string_buffer(cap)
    n = L.ref_lzbuffer(data, len(data), a, out, cap)
    assert 0 <= n <= cap
    return out.raw[:n]


def inputs():
    rng = np.random.default_rng(20261019)
    x86 = bytearray(rng.integers(0, 256, 30000, dtype=np.uint8).tobytes())
    for i in range(0, len(x86) - 8, 23):
        x86[i] = 0xE8 if i % 2 else 0xE9
        x86[i + 4] = 0x00 if i % 3 else 0xFF
    return {
        "empty": b"", "one": b"a", "five": b"abcde", "run": b"\x00" * 5000 + b"\x01" + b"\x00" * 3000,
        "period": (b"0123456789abcdef" * 3000)[:40001],
        "text": synth.blocks("text", 11, 1, 90000).tobytes(),
        "mixed": synth.blocks("mixed", 12, 1, 150000).tobytes(),
        "x86": bytes(x86),
        "random": bytes(rng.integers(0, 256, 20000, dtype=np.uint8)),
    }


# method strings whose pre-processing arguments cover: bit-packed (1) and byte-aligned (2) LZ77, hash matcher with and
# without a second context order and look-ahead, suffix-array matcher (N6 - N1 >= 21), BWT (3), each with and without E8E9
METHODS = ["x0,1,4,0,3,20", "x0,1,4,0,7,21,1", "x0,1,5,0,1,16", "x0,1,6,8,2,18,2", "x0,2,12,0,7,21,1", "x0,2,4,0,3,19", "x0,2,3,5,2,17,2",
           "x0,2,1,0,0,12", "x0,3", "x0,5,4,0,3,19", "x0,5,4,0,7,21,1", "x0,6,8,0,5,18", "x0,7", "x1,1,4,0,3,22", "x1,2,5,0,4,22,1"]


@pytest.mark.parametrize("method", METHODS)
def test_preprocessor_matches_reference_lzbuffer(ref, method):
    _, args = frontend.make_config(method)      # LibZPAQ.cs:394-415: the numeric arguments of the method string
    for name, data in inputs().items():
        if name in ("mixed",) and args[1] & 3 != 3 and args[5] - args[0] >= 21 and len(data) > 100000:
            data = data[:100000]        # the reference's windowed-ISA rebuild is quadratic-ish: keep the SA cases short
        got = po.preprocess(data, args)
        want = ref_preprocess(ref, data, args)
        assert got == want, (method, name, len(got), len(want))


def test_e8e9_matches_reference(ref):
    for name, data in inputs().items():
        if len(data) < 5:
            continue
        a = C.create_string_buffer(data, len(data))
        ref.ref_e8e9(a, len(data))
        b = C.create_string_buffer(data, len(data))
        po.lib().orc_e8e9(C.cast(b, C.c_void_p), len(data))
        assert a.raw == b.raw, name
