"""Pins the ZPAQL virtual machine (HCOMP contexts, PCOMP post-processing) against the REFERENCE's own text: the body
of ZPAQL.execute (ZPAQL.cs:1028-1251, the 256-way interpreter) is compiled where it lies by oracle/build_ref.py into
oracle/_ref/ and run on the same programs and inputs as the oracle's VM, which decodes the instruction set by field.
Skipped when the fragment is not available."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, frontend, pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

COMPSIZE = [0, 2, 3, 2, 3, 4, 6, 6, 3, 5]      # Component.cs:27-43


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build_zpaql()
    if not path or not os.path.exists(path):
        pytest.skip("reference ZPAQL fragment not built (no /root/reference here and no oracle/_ref)")
    L = C.CDLL(path)
    L.ref_zpaql_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_ulonglong, C.c_ulonglong]
    L.ref_zpaql_run.restype = C.c_longlong
    return L


def split_header(hdr: bytes):
    """-> (hh, hm, ph, pm, n, HCOMP program incl. END)."""
    n = hdr[6]
    pos = 7
    for _ in range(n):
        pos += COMPSIZE[hdr[pos]]
    assert hdr[pos] == 0
    return hdr[2], hdr[3], hdr[4], hdr[5], n, hdr[pos + 1:]


def synthetic_header(hh, hm, ph, pm, prog: bytes) -> bytes:
    body = bytes([hh, hm, ph, pm, 0, 0]) + prog       # n = 0, COMP END, program (incl. its END)
    return bytes([len(body) & 255, len(body) >> 8]) + body


def ref_run(L, prog, hbits, mbits, inputs, hn, cap=1 << 22):
    arr = np.asarray(inputs, dtype=np.uint32)
    h = np.zeros(max(hn, 1), dtype=np.uint32)
    out = C.create_string_buffer(cap)
    n = L.ref_zpaql_run(prog, len(prog), hbits, mbits, arr.ctypes.data, len(arr), h.ctypes.data, hn, out, cap, 1 << 34)
    return n, out.raw[:max(n, 0)], h[:hn]


HCOMP_MODELS = [("level", 1), ("level", 2), ("level", 3), ("method", "x0,0c256,0,255,255"), ("method", "x0,0w2,48,10,255,16,1a30,1,2"),
                ("method", "s4,4c0,0,255i1,2,3,4ms20,10,100t5,20"), ("method", "x0,2,12,0,7,21,1c0,0,511i2m"),
                ("method", "x0,0c0,1003,255c0,7c0,0,1300,255c200,0,511,300a24,1,1m12,20s9,20,100t3"), ("method", "x0,3ci1")]


@pytest.mark.parametrize("what", HCOMP_MODELS, ids=[str(w[1]) for w in HCOMP_MODELS])
def test_hcomp_contexts_match_reference_interpreter(ref, what):
    if what[0] == "level":
        hdr, _ = frontend.builtin_model(what[1])
    else:
        text, args = frontend.make_config(what[1])
        hdr, _, _ = frontend.compile_config(text, args)
    hh, hm, ph, pm, n, prog = split_header(bytes(hdr))
    data = synth.blocks("mixed", 77, 1, 6000).tobytes() + synth.blocks("text", 78, 1, 3000).tobytes()
    hn = min(max(n, 8), 1 << hh)
    for upto in (1, 2, 3, 17, 1000, len(data)):
        _, h_orc = po.zpaql_run(bytes(hdr), data[:upto], pp=False, eof_call=False, nh=hn)
        rc, _, h_ref = ref_run(ref, prog, hh, hm, list(data[:upto]), hn)
        assert rc == 0 and np.array_equal(h_orc, h_ref), (what, upto)


PCOMP_METHODS = ["x0,1,4,0,3,20", "x0,1,4,0,7,21,1", "x0,2,12,0,7,21,1", "x0,2,4,0,3,19", "x0,3", "x0,5,4,0,3,19", "x0,6,8,0,5,18", "x0,7", "x0,4"]


@pytest.mark.parametrize("method", PCOMP_METHODS)
def test_pcomp_postprocessing_matches_reference_interpreter(ref, method):
    # the PCOMP programs makeConfig emits (LibZPAQ.cs:427-830: lazy2, lzpre, bwtrle, e8e9), run by both interpreters on the
    # pre-processed stream: same OUT bytes (the original data), same final H
    text, args = frontend.make_config(method)
    hdr, pcomp, _ = frontend.compile_config(text, args)
    assert len(pcomp) > 0
    _, _, ph, pm, _, _ = split_header(bytes(hdr))
    data = synth.blocks("mixed", 91, 1, 30000).tobytes()
    pre = po.preprocess(data, args) if args[1] else data      # args[1] == 4 is E8E9 only: done by compressBlock itself
    if args[1] == 4:
        buf = C.create_string_buffer(data, len(data))
        po.lib().orc_e8e9(C.cast(buf, C.c_void_p), len(data))
        pre = buf.raw
    shdr = synthetic_header(0, 0, ph, pm, bytes(pcomp))
    out_orc, h_orc = po.zpaql_run(shdr, pre, pp=True, eof_call=True, nh=8 if ph >= 3 else 1 << ph)
    rc, out_ref, h_ref = ref_run(ref, bytes(pcomp), ph, pm, list(pre) + [0xFFFFFFFF], len(h_orc))
    assert rc == len(out_ref) and out_ref == out_orc and np.array_equal(h_orc, h_ref)
    assert out_ref == data                                    # and the reference's own VM restores the input


def test_random_programs_match_reference_interpreter(ref):
    # every defined opcode with random operands; jumps only forward (the oracle has no instruction budget)
    rng = np.random.default_rng(777)
    undefined = {0, 58} | set(range(120, 128)) | set(range(240, 255)) | {op for op in range(64) if op % 8 in (5, 6)}
    ops = [op for op in range(255) if op not in undefined and op not in (56, 255)]      # no HALT / LJ inside
    agree = errors = 0
    for trial in range(300):
        prog = bytearray()
        for _ in range(int(rng.integers(5, 60))):
            op = int(rng.choice(ops))
            prog.append(op)
            if op & 7 == 7:
                prog.append(int(rng.integers(0, 6)) if op in (39, 47, 63) else int(rng.integers(0, 256)))   # short forward jumps
        prog += bytes([56, 0])                               # HALT, END
        hh, hm = int(rng.integers(0, 6)), int(rng.integers(0, 9))
        inputs = [int(x) for x in rng.integers(0, 256, 40)]
        shdr = synthetic_header(hh, hm, 0, 0, bytes(prog))
        rc, out_ref, h_ref = ref_run(ref, bytes(prog), hh, hm, inputs, 1 << hh)
        try:
            out_orc, h_orc = po.zpaql_run(shdr, bytes(inputs), pp=False, eof_call=False, nh=1 << hh)
        except Exception:
            assert rc == -1, trial                           # both must reject the program
            errors += 1
            continue
        assert rc >= 0 and np.array_equal(h_orc, h_ref), trial
        # (HCOMP has no OUT destination in the oracle's inith mode; OUT is compared in the PCOMP test)
        agree += 1
    assert agree > 200
