"""Pins the 32-bit arithmetic coder against the REFERENCE's own text: the bodies of Encoder.encode (Encoder.cs:86-103)
and Decoder.decode (Decoder.cs:136-158) are compiled where they lie by oracle/build_ref.py into oracle/_ref/ and driven
with the same (bit, probability) sequences as the oracle's coder.  Skipped when the fragment is not available."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, pyoracle as po  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build_coder()
    if not path or not os.path.exists(path):
        pytest.skip("reference coder fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    L.ref_arith_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_ulonglong]
    L.ref_arith_encode.restype = C.c_longlong
    L.ref_arith_decode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_uint, C.c_void_p]
    L.ref_arith_decode.restype = C.c_int
    return L


def sequences():
    rng = np.random.default_rng(424242)
    out = []
    for n, skew in ((1, 0.5), (9, 0.5), (5000, 0.5), (200000, 0.9), (200000, 0.02), (100000, 0.999)):
        p = rng.integers(0, 32768, n).astype(np.uint32) * 2 + 1          # predict() * 2 + 1, Encoder.cs:51
        if skew != 0.5:
            p = np.clip((p.astype(np.float64) * 0 + 65536 * skew + rng.normal(0, 3000, n)), 1, 65535).astype(np.uint32) | 1
        bits = (rng.random(n) < p / 65536.0).astype(np.uint8)            # bit 1 with probability p / 64K
        flags = rng.random(n) < 0.11                                      # the per-byte encode(0, 0) flag, Encoder.cs:49
        p = np.where(flags, 0, p).astype(np.uint16)
        bits = np.where(flags, 0, bits).astype(np.uint8)
        out.append((bits, p))
    # adversarial: probabilities at the extremes, bits against the prediction
    n = 50000
    p = np.where(rng.random(n) < 0.5, 1, 65535).astype(np.uint16)
    out.append(((rng.random(n) < 0.5).astype(np.uint8), p))
    return out


@pytest.mark.parametrize("k", range(7))
def test_arithmetic_coder_matches_reference_text(ref, k):
    bits, probs = sequences()[k]
    n = len(bits)
    cap = n * 3 + 64
    a = C.create_string_buffer(cap)
    b = C.create_string_buffer(cap)
    L = po.lib()
    L.orc_arith_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
    L.orc_arith_encode.restype = C.c_int64
    L.orc_arith_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]
    L.orc_arith_decode.restype = C.c_int
    na = L.orc_arith_encode(bits.ctypes.data, probs.ctypes.data, n, a, cap)
    nb = ref.ref_arith_encode(bits.ctypes.data, probs.ctypes.data, n, b, cap)
    assert na == nb and na > 0 and a.raw[:na] == b.raw[:nb]
    stream = a.raw[:na] + b"\x00\x00\x00\x00"                           # Compressor.cs:235-238
    d1 = np.zeros(n, dtype=np.uint8)
    d2 = np.zeros(n, dtype=np.uint8)
    assert L.orc_arith_decode(stream, len(stream), probs.ctypes.data, n, d1.ctypes.data) == 0
    assert ref.ref_arith_decode(stream, len(stream), probs.ctypes.data, n, d2.ctypes.data) == 0
    assert np.array_equal(d1, bits) and np.array_equal(d2, bits)
