"""CPU: the decoder's choice of post-processor (zpq_post_kind).  The four PCOMP programs makeConfig emits
(LibZPAQ.cs:427-826) are recognised byte for byte -- also when they come from the oracle's independently written front end;
anything else is left to the interpreter pass."""
import pytest

PK_LZ_BITS, PK_LZ_BYTES, PK_BWT, PK_E8E9 = 2, 3, 4, 5

CASES = [
    ("x0,1,4,0,7,21,1", PK_LZ_BITS, 0, 0), ("x0,5,4,3,3,19,1", PK_LZ_BITS, 1, 0), ("x6,1,4,0,3,24", PK_LZ_BITS, 0, 2),
    ("x8,5,4,0,3,24", PK_LZ_BITS, 1, 4),
    ("x0,2,12,0,7,21,1c0,0,511i2m", PK_LZ_BYTES, 0, 12), ("x0,6,4,0,3,19", PK_LZ_BYTES, 1, 4), ("x4,2,5,0,3,22c0,0,511", PK_LZ_BYTES, 0, 5),
    ("x0,3ci1", PK_BWT, 0, 0), ("x2,3ci1", PK_BWT, 0, 0), ("x0,7ci1", PK_BWT, 1, 0), ("x5,3ci1", PK_BWT, 0, 0), ("x6,7ci1", PK_BWT, 1, 0),
    ("x0,4c0,0,255", PK_E8E9, 1, 0),
]


@pytest.mark.parametrize("method,kind,e8,param", CASES)
def test_makeconfig_programs_are_recognised(zlib_, method, kind, e8, param):
    from oracle import frontend as fe
    text, args = zlib_.make_config(method)
    hdr, pcomp = zlib_.compile_config(text, args)
    assert len(pcomp) > 0
    ph, pm = hdr[4], hdr[5]
    got = zlib_.post_kind(ph, pm, pcomp)
    assert (got & 15, (got >> 4) & 1, got >> 8) == (kind, e8, param)
    # the oracle's front end assembles the same program bytes
    otext, oargs = fe.make_config(method)
    ohdr, opcomp = fe.compile_config(otext, oargs)[:2]
    assert zlib_.post_kind(ohdr[4], ohdr[5], bytes(opcomp)) == got


def test_other_programs_are_interpreted(zlib_):
    text, args = zlib_.make_config("x0,1,4,0,7,21,1")
    hdr, pcomp = zlib_.compile_config(text, args)
    ph, pm = hdr[4], hdr[5]
    assert zlib_.post_kind(ph, pm, pcomp) != 0
    bad = bytearray(pcomp); bad[10] ^= 1
    assert zlib_.post_kind(ph, pm, bytes(bad)) == 0              # one opcode differs
    assert zlib_.post_kind(ph, pm, pcomp[:-1]) == 0               # truncated
    assert zlib_.post_kind(ph, pm + 1, pcomp) in (0, zlib_.post_kind(ph, pm, pcomp))   # rb depends on pm only above 16 MB blocks
    assert zlib_.post_kind(0, 0, pcomp) == 0                      # an LZ77 program needs its buffer
    # a hand-written program (a foreign archive's): copies its input
    h2, p2 = zlib_.compile_config("comp 0 0 0 0 0 hcomp halt pcomp copy ; a> 255 ifnot out endif halt end", [0] * 9)
    assert zlib_.post_kind(0, 0, p2) == 0


@pytest.mark.parametrize("method", ["x0,1,4,0,7,21,1", "x0,6,8,0,5,18c0,0,255", "x0,7ci1", "x5,7ci1", "x0,4c0,0,255"])
def test_pcomp_programs_translate_and_compile_with_nvrtc(zlib_, method):
    """The device analogue of the reference's x86 JIT for PCOMP (ZPAQL.cs:353-1008): ZPAQL -> CUDA -> sm_100a cubin, no GPU needed."""
    text, args = zlib_.make_config(method)
    hdr, pcomp = zlib_.compile_config(text, args)
    n, src, log = zlib_.specialize_pcomp(hdr[4], hdr[5], pcomp)
    assert n > 1000, log
    assert "zpq_post_rt" in src and "zpq_prog_run" in src and "goto Lhalt;" in src


def test_foreign_pcomp_with_loops_and_long_jumps_compiles(zlib_):
    cfg = ("comp 2 4 0 0 1 0 cm 16 255 hcomp c++ *c=a b=c a=0 hash *d=a halt "
           "pcomp foreign ; a> 255 ifnotl a^= 32 b=a c= 3 do a=b out c-- a=c a> 0 while elsel a= 33 out endif halt end")
    hdr, pcomp = zlib_.compile_config(cfg, [0] * 9)
    assert 255 in pcomp                                           # an LJ is in there
    n, src, log = zlib_.specialize_pcomp(0, 0, pcomp)
    assert n > 1000, log
    assert "if (--budget == 0) goto Lerr;" in src                 # the backward jump of the loop is bounded
    # a jump into the middle of an instruction cannot be translated: reported, left to the interpreter
    bad = bytes([63, 0, 71, 5, 63, 0xFD, 56, 0])                   # jmp +0 ; a= 5 ; jmp -3 (into the operand of a=) ; halt
    n, src, log = zlib_.specialize_pcomp(0, 0, bad)
    assert n < 0 and "middle of an instruction" in log
