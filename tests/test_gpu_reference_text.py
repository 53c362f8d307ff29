"""GPU: the CUDA path against the REFERENCE's own text directly, on the GPU box (oracle/_ref travels with the snapshot):
archives produced by the device are decoded by the reference's Decompresser / Decoder / PostProcessor / predictor text, and
archive blocks produced by the reference's Compressor text are decoded by the device.  Skipped where oracle/_ref is absent."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _tables(L):
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
    tabs = [np.asarray(kat["sdt2k"], dtype=np.int32), np.asarray(kat["sdt"], dtype=np.int32), np.asarray(kat["ssquasht"], dtype=np.uint16),
            np.asarray(kat["stdt"], dtype=np.int32), np.asarray(kat["sns"], dtype=np.uint8)]
    L.ref_predictor_tables.argtypes = [C.c_void_p] * 5
    L.ref_predictor_tables(*[t.ctypes.data for t in tabs])


@pytest.fixture(scope="module")
def ref_dec():
    from oracle import build_ref
    path = build_ref.build_decompresser()
    if not path or not os.path.exists(path):
        pytest.skip("oracle/_ref/libdecompresser_ref.so not available")
    L = C.CDLL(path)
    _tables(L)
    L.ref_decompress.argtypes = [C.c_char_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.ref_decompress.restype = C.c_longlong
    return L


@pytest.fixture(scope="module")
def ref_comp():
    from oracle import build_ref
    path = build_ref.build_compressor()
    if not path or not os.path.exists(path):
        pytest.skip("oracle/_ref/libcompressor_ref.so not available")
    L = C.CDLL(path)
    _tables(L)
    L.ref_compress_block.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_ulonglong,
                                     C.c_char_p, C.c_int, C.c_void_p, C.c_ulonglong]
    L.ref_compress_block.restype = C.c_longlong
    return L


@pytest.mark.parametrize("how,arg", [("level", 1), ("level", 2), ("level", 3), ("method", "x0,0c256,0,255,255"), ("method", "2"),
                                     ("method", "30,128,1"), ("method", "x0,2,12,0,7,21,1c0,0,511i2m"), ("method", "x0,7ci1"), ("method", "0")])
def test_reference_decompresser_text_decodes_device_archives(gpu_ctx, ref_dec, how, arg):
    from tools import synth
    data = synth.blocks("mixed", 910, 1, 60000).tobytes()
    cuts = [0, 25000, 25000, 25001, 60000]                   # ragged: an empty and a one-byte block
    offs = np.asarray(cuts, dtype=np.uint64)
    arc, ooff = (gpu_ctx.compress_blocks_level(data, offs, arg) if how == "level" else gpu_ctx.compress_blocks(data, offs, arg))
    a = arc.tobytes()
    out = C.create_string_buffer(len(data) + 16)
    marks = C.create_string_buffer(21 * 8)
    nseg = C.c_int(0)
    n = ref_dec.ref_decompress(a, len(a), out, len(data) + 16, marks, 8, C.byref(nseg))
    assert n == len(data) and out.raw[:n] == data
    assert nseg.value == 4 and all(marks.raw[21 * i] == 1 for i in range(4))


@pytest.mark.parametrize("level", [1, 2, 3])
def test_device_decodes_reference_compressor_text_blocks(gpu_ctx, ref_comp, level):
    import hashlib
    from tools import synth
    blocks = [synth.blocks("mixed", 920 + i, 1, n).tobytes() if n else b"" for i, n in enumerate([30000, 0, 1, 12345])]
    arcs = []
    for b in blocks:
        cap = len(b) * 2 + 70000
        out = C.create_string_buffer(cap)
        n = ref_comp.ref_compress_block(level, None, None, 0, None, str(len(b)).encode(), b, len(b), hashlib.sha1(b).digest(), 1, out, cap)
        assert 0 < n <= cap
        arcs.append(out.raw[:n])
    arc = b"".join(arcs)
    offs = np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(arc, offs)
    assert out.tobytes() == b"".join(blocks)
    assert sha.tolist() == [1] * 4 and bst.tolist() == [0] * 4
