"""Pins the suffix-array / BWT path against the REFERENCE's own code: the body of
/root/reference/ZPAQSharp/divsufsort.cs (libdivsufsort-lite as C text) is compiled where it lies by
oracle/build_ref.py into oracle/_ref/ (git-ignored, travels to the GPU box).  The oracle's suffix sorter and its
BWT stream (LZBuffer.cs:229-241 on divsufsort.cs:1940) must agree with it on every input below.
Skipped when the fragment is not available (no /root/reference and no prebuilt oracle/_ref)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if not path or not os.path.exists(path):
        pytest.skip("reference divsufsort fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    L.divsufsort.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    L.divsufsort.restype = C.c_int
    L.divbwt.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
    L.divbwt.restype = C.c_int
    return L


def cases():
    rng = np.random.default_rng(20261018)
    out = [b"", b"a", b"ab", b"ba", b"aa", b"banana", b"mississippi", b"abracadabra" * 50, b"\x00" * 1000, b"\xff" * 777 + b"\x00",
           bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), bytes(rng.integers(0, 2, 4096, dtype=np.uint8)),
           bytes(rng.integers(97, 101, 20000, dtype=np.uint8))]
    out.append(synth.blocks("text", 3, 1, 60000).tobytes())
    out.append(synth.blocks("mixed", 5, 1, 70001).tobytes())
    out.append((b"the quick brown fox jumps over the lazy dog. " * 400)[:17001])
    return out


def ref_sa(L, data: bytes) -> np.ndarray:
    n = len(data)
    sa = np.zeros(max(n, 1), dtype=np.int32)
    if n:
        assert L.divsufsort(data, sa.ctypes.data, n) == 0
    return sa[:n]


def oracle_sa(data: bytes) -> np.ndarray:
    n = len(data)
    sa = np.zeros(max(n, 1), dtype=np.int32)
    po._check(po.lib().orc_suffix_array(data, n, sa.ctypes.data))
    return sa[:n]


@pytest.mark.parametrize("k", range(16))
def test_suffix_array_matches_reference(ref, k):
    data = cases()[k]
    assert np.array_equal(oracle_sa(data), ref_sa(ref, data))


def bwt_stream_from_sa(data: bytes, sa: np.ndarray) -> bytes:
    # LZBuffer.cs:229-241, written out with the REFERENCE's suffix array
    n = len(data)
    out = bytearray()
    idx = 0
    for i in range(n + 5):
        if i == 0:
            out.append(data[n - 1] if n else 255)
        elif i > n:
            out.append(idx & 255); idx >>= 8
        elif sa[i - 1] == 0:
            idx = i; out.append(255)
        else:
            out.append(data[sa[i - 1] - 1])
    return bytes(out)


@pytest.mark.parametrize("k", [3, 5, 6, 7, 8, 10, 12, 13])
def test_bwt_stream_matches_reference(ref, k):
    data = cases()[k]
    args = [0, 3, 0, 0, 0, 0, 0, 0, 0]
    got = po.preprocess(data, args)
    assert got == bwt_stream_from_sa(data, ref_sa(ref, data))


def test_reference_divbwt_agrees_with_its_own_suffix_array(ref):
    # divbwt (divsufsort.cs:1973) is the reference's direct BWT: same permutation as reading T[SA[i]-1]
    data = cases()[7]
    n = len(data)
    u = C.create_string_buffer(n)
    pidx = ref.divbwt(data, u, None, n)
    sa = ref_sa(ref, data)
    expect = bytearray()
    expect.append(data[n - 1])
    primary = 0
    for i in range(n):
        if sa[i] == 0:
            primary = i + 1
        else:
            expect.append(data[sa[i] - 1])
    assert pidx == primary and bytes(u.raw) == bytes(expect)
