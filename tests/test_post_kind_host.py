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
