"""CPU: the ZPAQL -> C translation the device compiles with NVRTC (zpq_codegen.cpp: translate_zpaql, used for HCOMP in the coding
kernels and for stored PCOMP programs in the post-processing pass -- the device analogue of the reference's x86 JIT,
ZPAQL.cs:353-1008).  The generated function body is plain C with gotos: here it is compiled with g++ and run on the host against
the oracle's ZPAQL machine (itself pinned to the reference's ZPAQL.execute text, tests/test_reference_zpaql.py) -- for the four
programs makeConfig emits on real transformed streams, and for random programs over every defined opcode."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

HARNESS = r"""
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
namespace zpq {
%s
}
extern "C" long long run_all(const uint32_t* inputs, long long n, int ph, int pm, uint8_t* out, long long cap, uint32_t* hout, int nh) {
  uint32_t* H = (uint32_t*)calloc((size_t)1 << ph, 4);
  uint8_t* M = (uint8_t*)calloc(((size_t)1 << pm) + 8, 1);
  uint32_t R[256]; memset(R, 0, sizeof R);
  zpq::PVm vm; vm.b = vm.c = vm.d = vm.f = 0;
  uint64_t opos = 0;
  long long rc = 0;
  for (long long i = 0; i < n && !rc; ++i) if (zpq::zpq_prog_run(inputs[i], vm, H, M, R, out, opos, (uint64_t)cap, 1LL << 40)) rc = -1;
  for (int i = 0; i < nh; ++i) hout[i] = H[i & ((1u << ph) - 1)];
  free(H); free(M);
  return rc ? rc : (long long)opos;
}
"""


def _host_function(zlib_, ph, pm, prog):
    """The translated program as a host shared library, or None when it cannot be translated."""
    n, src, log = zlib_.specialize_pcomp(ph, pm, prog)
    if n < 0:
        assert "middle of an instruction" in log, log
        return None
    body = src[src.index("struct PVm"):src.index("}  // namespace zpq")]
    body = body.replace("static __device__ __noinline__ int", "static int")
    d = tempfile.mkdtemp(prefix="zpq_tr_")
    cpp, so = os.path.join(d, "p.cpp"), os.path.join(d, "p.so")
    open(cpp, "w").write(HARNESS % body)
    subprocess.check_call(["g++", "-O1", "-w", "-shared", "-fPIC", "-o", so, cpp])
    L = C.CDLL(so)
    L.run_all.restype = C.c_longlong
    L.run_all.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int]
    return L


def _run(L, inputs, ph, pm, nh, cap=1 << 22):
    arr = np.asarray(inputs, dtype=np.uint32)
    out = C.create_string_buffer(cap)
    h = np.zeros(max(nh, 1), dtype=np.uint32)
    n = L.run_all(arr.ctypes.data, len(arr), ph, pm, out, cap, h.ctypes.data, nh)
    return n, out.raw[:max(n, 0)], h


def _shdr(ph, pm, prog):
    body = bytes([0, 0, ph, pm, 0, 0]) + bytes(prog)
    return bytes([len(body) & 255, len(body) >> 8]) + body


@pytest.mark.parametrize("method", ["x0,1,4,0,7,21,1", "x0,5,4,3,3,19,1", "x0,2,12,0,7,21,1c0,0,255", "x0,6,8,0,5,18c0,0,255", "x0,3ci1", "x0,7ci1",
                                    "x5,3ci1", "x5,7ci1", "x0,4c0,0,255"])
def test_translated_makeconfig_programs_restore_the_data(zlib_, oracle, method):
    from oracle import frontend as fe
    from tools import synth
    text, args = fe.make_config(method)
    hdr, pcomp = fe.compile_config(text, args)[:2]
    ph, pm = hdr[4], hdr[5]
    L = _host_function(zlib_, ph, pm, bytes(pcomp))
    assert L is not None
    data = synth.blocks("mixed", 4000, 1, 30000).tobytes()
    pre = oracle.preprocess(data, list(args))
    for stream in (pre, pre[:len(pre) // 2] if method.split(",")[1][0] not in "37" else pre):
        n, out, h = _run(L, list(stream) + [0xFFFFFFFF], ph, pm, 1)
        want, _ = oracle.zpaql_run(_shdr(ph, pm, pcomp), stream, pp=True, eof_call=True, nh=1)
        assert n == len(want) and out == want
    n, out, h = _run(L, list(pre) + [0xFFFFFFFF], ph, pm, 1)
    assert out == data


def test_translated_random_programs_match_the_oracle(zlib_, oracle):
    # every defined opcode with random operands, OUT included; short forward jumps (some land inside an instruction: those
    # programs are refused by the translator and belong to the interpreter)
    rng = np.random.default_rng(4242)
    undefined = {0, 58} | set(range(120, 128)) | set(range(240, 255)) | {op for op in range(64) if op % 8 in (5, 6)}
    ops = [op for op in range(255) if op not in undefined and op not in (56, 255)]
    agree = refused = errors = 0
    for trial in range(60):
        prog = bytearray()
        for _ in range(int(rng.integers(5, 50))):
            op = int(rng.choice(ops + [57] * 6))                 # more OUTs
            prog.append(op)
            if op & 7 == 7:
                prog.append(int(rng.integers(0, 6)) if op in (39, 47, 63) else int(rng.integers(0, 256)))
        prog += bytes([56, 0])
        ph, pm = int(rng.integers(0, 6)), int(rng.integers(0, 9))
        inputs = [int(x) for x in rng.integers(0, 256, 40)]
        L = _host_function(zlib_, ph, pm, bytes(prog))
        if L is None:
            refused += 1
            continue
        n, out, h = _run(L, inputs, ph, pm, 1 << ph)
        try:
            want, h_orc = oracle.zpaql_run(_shdr(ph, pm, prog), bytes(inputs), pp=True, eof_call=False, nh=1 << ph)
        except Exception:
            assert n == -1, trial                                # both must reject the program
            errors += 1
            continue
        assert n == len(want) and out == want and np.array_equal(h, h_orc), trial
        agree += 1
    assert agree >= 40 and agree + refused + errors == 60
