"""Pins the DECODE side -- block search, header parsing, segment header, arithmetic decoding, the PostProcessor state
machine with PCOMP programs run by the reference interpreter, end-of-segment markers -- against the REFERENCE's own text:
Decompresser.findBlock / findFilename / readComment / decompress / readSegmentEnd (Decompresser.cs:29-194),
Decoder.decompress / skip / init / decode (Decoder.cs:32-111, 136-158), PostProcessor.init / write (PostProcessor.cs:27-86),
compiled where they lie by oracle/build_ref.py and driven like LibZPAQ.decompress (LibZPAQ.cs:65-79).  The reference text
must restore the original bytes from the oracle's archives (which the device's archives equal byte for byte, tests -m gpu),
and agree with the oracle's own decoder on the segment checksums.  Skipped when the fragment is not available."""
import ctypes as C
import json
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
    path = build_ref.build_decompresser()
    if not path or not os.path.exists(path):
        pytest.skip("reference Decompresser fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
    tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
            np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
    L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
    L.ref_predictor_tables(*[t.ctypes.data for t in tabs])
    L.ref_decompress.argtypes = [C.c_char_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.ref_decompress.restype = C.c_longlong
    return L


def _ref_decompress(L, arc, cap):
    out = C.create_string_buffer(max(cap, 1))
    marks = C.create_string_buffer(21 * 64)
    nseg = C.c_int(0)
    n = L.ref_decompress(arc, len(arc), out, cap, marks, 64, C.byref(nseg))
    return n, out.raw[:max(n, 0)], [marks.raw[21 * i:21 * i + 21] for i in range(min(nseg.value, 64))]


def _data(seed, n):
    return synth.blocks("mixed", seed, 1, n).tobytes()


@pytest.mark.parametrize("level", [1, 2, 3])
def test_reference_decoder_restores_builtin_level_blocks(ref, level):
    data = _data(600 + level, 30000)
    arc = po.compress_block_level(data, level, filename="f", comment=None)
    n, out, marks = _ref_decompress(ref, arc, len(data) + 16)
    assert n == len(data) and out == data
    assert marks == [b"\x01" + po.sha1(data)]


METHODS = ["0", "x0,0c256,0,255,255", "1", "2", "3", "x4,1,4,0,3,24c0,0,511", "x4,3ci1", "x4,4c0,0,255", "x4,5,12,0,3,20,1c0,0,511i2",
           "x4,7ci1,1m", "s4,0,0,255i1,2ms20"]


@pytest.mark.parametrize("method", METHODS)
def test_reference_decoder_restores_method_blocks(ref, method):
    # LZ77 / BWT / E8E9 archives carry a PCOMP program: the reference PostProcessor loads it and the reference interpreter runs it
    data = _data(87, 20000) + b"\xe8\x10\x00\x00\x00" * 50 + _data(88, 5000)
    arc = po.compress_block(data, method, filename="x", comment="c")
    n, out, marks = _ref_decompress(ref, arc, len(data) + 16)
    assert n == len(data) and out == data
    assert marks == [b"\x01" + po.sha1(data)]


def test_several_blocks_garbage_in_front_and_no_checksum(ref):
    a, b, c = _data(1, 5000), b"", _data(2, 70000)
    arc = (b"junk before the first tag" + po.compress_block(a, "1") + po.compress_block_level(b, 1) +
           po.compress_block(c, "x0,0c256,0,255,255", dosha1=False))
    n, out, marks = _ref_decompress(ref, arc, len(a) + len(c) + 16)
    assert out == a + b + c
    assert marks == [b"\x01" + po.sha1(a), b"\x01" + po.sha1(b), b"\x00" * 21]
    want, st = po.decompress(arc)
    assert want == out and len(st) == 3


def test_reference_decoder_rejects_a_damaged_stream(ref):
    data = _data(5, 20000)
    arc = bytearray(po.compress_block_level(data, 2))
    arc[len(arc) // 2] ^= 0x40
    n, out, marks = _ref_decompress(ref, bytes(arc), len(data) * 4)
    # the coder either runs off its range ("archive corrupted") or decodes garbage whose end marker is missing
    assert n == -1 or out != data
    try:                                     # and the oracle's decoder takes the same view of it
        got, st = po.decompress(bytes(arc))
        assert got != data or st[0] != 1
    except po.OracleError:
        pass


@pytest.mark.parametrize("what", [("level", 1), ("level", 2), ("level", 3), ("method", "x0,0c256,0,255,255"), ("method", "x4,3ci1"),
                                  ("method", "x0,0c1,0,255,255a24mm16ts19t0w2"), ("method", "x6,1,4,0,3,24"), ("method", "0")],
                         ids=lambda w: str(w[1]))
def test_block_memory_matches_reference_zpaql_memory(ref, what):
    # ZPAQL.memory() (ZPAQL.cs:58-81) as it lies, on the header as ZPAQL.read parses it; mid.cfg: 111,424,926
    from oracle import frontend
    from zpaqsharp_b200 import libzpaq as z
    hdr = bytes(frontend.builtin_model(what[1])[0]) if what[0] == "level" else bytes(frontend.plan_block(what[1], b"x" * 1000)["hdr"])
    ref.ref_block_memory.argtypes = [C.c_char_p]
    ref.ref_block_memory.restype = C.c_double
    want = ref.ref_block_memory(hdr)
    assert want > 0
    assert po.block_memory(hdr) == want
    assert z.block_memory(hdr) == want                      # the product's host code (no GPU needed)
    if what == ("level", 2):
        assert int(want) == 111424926


@pytest.mark.parametrize("level", [1, 2, 3])
def test_reference_decoder_restores_multi_segment_blocks(ref, level):
    """Several segments in one block (Compressor.cs:133-146: startSegment after endSegment; Decompresser.cs:128-134: the decoder
    and the post-processor are initialised for the first segment only): the reference's Decompresser text decodes what the
    oracle's multi-segment writer produces, segment by segment, with every stored SHA-1."""
    from oracle import frontend as fe
    data = _data(90 + level, 9000)
    hdr, _ = fe.builtin_model(level)
    cuts = [0, 2500, 2500, 2501, 7000, 9000]                      # an empty and a one-byte segment among them
    arc = po.compress_segments(bytes(hdr), b"", data, cuts)
    n, out, marks = _ref_decompress(ref, arc, len(data) + 16)
    assert n == len(data) and out == data
    assert marks == [b"\x01" + po.sha1(data[cuts[k]:cuts[k + 1]]) for k in range(len(cuts) - 1)]
    got, st = po.decompress(arc)
    assert got == data and st == [1] * (len(cuts) - 1)


@pytest.mark.parametrize("method", ["x0,4c0,0,255", "x0,2,12,0,7,21,1c0,0,255", "x0,1,4,0,7,21,1", "x0,5,4,0,3,19"])
def test_reference_decoder_restarts_the_pcomp_programs_at_every_segment_end(ref, method):
    """The native post-processors restore every segment on its own: makeConfig's programs reset their registers in their
    end-of-segment branch (LibZPAQ.cs:465, :599, :805-812).  The reference's Decompresser / PostProcessor / ZPAQL text says the same:
    a block whose segments carry independently transformed data decodes to the concatenation of the originals."""
    from oracle import frontend as fe
    text, args = fe.make_config(method)
    hdr, pcomp = fe.compile_config(text, args)[:2]
    data = _data(120, 24000)
    cuts = [0, 5000, 5000, 5001, 12000, 24000]
    parts = [data[cuts[k]:cuts[k + 1]] for k in range(len(cuts) - 1)]
    stream = [po.preprocess(p, list(args)) for p in parts]
    scuts = [0]
    for x in stream:
        scuts.append(scuts[-1] + len(x))
    arc = po.compress_segments(bytes(hdr), bytes(pcomp), b"".join(stream), scuts, dosha1=False)
    n, out, marks = _ref_decompress(ref, arc, len(data) + 64)
    assert n == len(data) and out == data
    got, _ = po.decompress(arc, cap=1 << 20)
    assert got == data
