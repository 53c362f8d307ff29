"""ctypes wrapper over oracle/libzpqoracle.so plus the Python front end.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import frontend as fe

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libzpqoracle.so")
    src = os.path.join(_HERE, "zpq_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libzpqoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_compress_block.restype = C.c_int64
        L.orc_compress_block.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_int),
                                         C.c_char_p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_uint64]
        L.orc_decompress.restype = C.c_int64
        L.orc_decompress.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                     C.POINTER(C.c_uint32)]
        L.orc_preprocess.restype = C.c_int64
        L.orc_preprocess.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_int), C.c_void_p, C.c_uint64]
        L.orc_block_memory.restype = C.c_double
        L.orc_block_memory.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_zpaql_run.restype = C.c_int64
        L.orc_zpaql_run.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, C.c_int,
                                    C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        L.orc_predict_trace.restype = C.c_int64
        L.orc_predict_trace.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p]
        L.orc_sha1.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
        L.orc_e8e9.argtypes = [C.c_void_p, C.c_int]
        L.orc_suffix_array.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.orc_tables.argtypes = [C.c_void_p] * 5
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    pass


def _check(r):
    if r < 0:
        raise OracleError(lib().orc_last_error().decode())
    return r


def _args9(args):
    return (C.c_int * 9)(*args)


def compress_with_model(hdr: bytes, pcomp: bytes, args, data: bytes, filename: str | None = None,
                        comment: str | None = None, dosha1: bool = True, with_tag: bool = True) -> bytes:
    """One archive block from explicit model bytes (Compressor.startBlock(hcomp) path)."""
    cap = len(data) * 2 + len(hdr) + len(pcomp) * 4 + 4096
    out = C.create_string_buffer(cap)
    n = _check(lib().orc_compress_block(hdr, len(hdr), pcomp, len(pcomp), _args9(args), data, len(data),
                                        filename.encode() if filename else None,
                                        comment.encode() if comment is not None else None,
                                        1 if dosha1 else 0, 1 if with_tag else 0, out, cap))
    return out.raw[:n]


def compress_segments(hdr: bytes, pcomp: bytes, data: bytes, cuts, dosha1: bool = True, with_tag: bool = True) -> bytes:
    """One block whose segments are data[cuts[k]:cuts[k+1]] (names seg0, seg1, ..., comment = size), Compressor.cs:133-248."""
    import numpy as np
    off = np.asarray(cuts, dtype=np.uint64)
    cap = len(data) * 2 + len(hdr) + len(pcomp) * 4 + 4096 + 64 * len(cuts)
    out = C.create_string_buffer(cap)
    L = lib()
    L.orc_compress_segments.restype = C.c_int64
    L.orc_compress_segments.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                        C.c_void_p, C.c_uint64]
    n = _check(L.orc_compress_segments(hdr, len(hdr), pcomp, len(pcomp), data, off.ctypes.data, len(cuts) - 1, 1 if dosha1 else 0,
                                       1 if with_tag else 0, out, cap))
    return out.raw[:n]


def compress_block(data: bytes, method: str, filename: str | None = None, comment: str | None = None,
                   dosha1: bool = True) -> bytes:
    """LibZPAQ.compressBlock, LibZPAQ.cs:117-325."""
    plan = fe.plan_block(method, data)
    cs = plan["comment"] + ((" " + comment) if comment else "")
    return compress_with_model(plan["hdr"], plan["pcomp"], plan["args"], data, filename, cs, dosha1, True)


def compress_block_level(data: bytes, level: int, filename: str | None = None, comment: str | None = None,
                         dosha1: bool = True, with_tag: bool = True) -> bytes:
    """Compressor.startBlock(int level) path (Compressor.cs:45-83): one block, one segment,
    comment = decimal size by the convention of SURVEY.md C2a."""
    hdr, pcomp = fe.builtin_model(level)
    cs = comment if comment is not None else str(len(data))
    return compress_with_model(hdr, pcomp, [0] * 9, data, filename, cs, dosha1, with_tag)


def block_size_of(method: str) -> int:
    """LibZPAQ.Compress block size, LibZPAQ.cs:87-94."""
    bs = 4
    if len(method) > 1 and method[1].isdigit():
        bs = int(method[1])
        if len(method) > 2 and method[2].isdigit():
            bs = bs * 10 + int(method[2])
        bs = min(bs, 11)
    return (0x100000 << bs) - 4096


def compress(data: bytes, method: str, filename: str | None = None, comment: str | None = None,
             dosha1: bool = True) -> bytes:
    """LibZPAQ.Compress, LibZPAQ.cs:84-108 (block split by the method's block-size digits)."""
    bs = block_size_of(method)
    out = bytearray()
    for off in range(0, len(data), bs):
        out += compress_block(data[off:off + bs], method, filename, comment, dosha1)
        filename = comment = None
    return bytes(out)


def decompress(archive: bytes, cap: int | None = None):
    """LibZPAQ.decompress, LibZPAQ.cs:65-79.  Returns (bytes, [sha status per segment])."""
    if cap is None:
        cap = max(1 << 20, len(archive) * 64)
    while True:
        out = C.create_string_buffer(cap)
        st = C.create_string_buffer(65536)
        nseg = C.c_uint32(0)
        n = lib().orc_decompress(archive, len(archive), out, cap, st, 65536, C.byref(nseg))
        if n < 0 and b"too small" in lib().orc_last_error():
            cap *= 4
            continue
        _check(n)
        return out.raw[:n], list(st.raw[:nseg.value])


def preprocess(data: bytes, args) -> bytes:
    cap = len(data) * 2 + 4096
    out = C.create_string_buffer(cap)
    n = _check(lib().orc_preprocess(data, len(data), _args9(args), out, cap))
    return out.raw[:n]


def sha1(data: bytes) -> bytes:
    out = C.create_string_buffer(20)
    lib().orc_sha1(data, len(data), out)
    return out.raw


def block_memory(hdr: bytes) -> float:
    return lib().orc_block_memory(hdr, len(hdr))


def predict_trace(hdr: bytes, data: bytes):
    import numpy as np
    probs = np.zeros(len(data) * 8, dtype=np.uint16)
    _check(lib().orc_predict_trace(hdr, len(hdr), data, len(data), probs.ctypes.data))
    return probs


def zpaql_run(hdr: bytes, data: bytes, pp: bool = False, eof_call: bool = False, nh: int = 8):
    import numpy as np
    cap = len(data) * 4 + 65536
    out = C.create_string_buffer(cap)
    h = np.zeros(nh, dtype=np.uint32)
    n = _check(lib().orc_zpaql_run(hdr, len(hdr), 1 if pp else 0, data, len(data), 1 if eof_call else 0,
                                   out, cap, h.ctypes.data, nh))
    return out.raw[:n], h
