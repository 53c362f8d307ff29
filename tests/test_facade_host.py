"""The streaming facade (zpaqsharp_b200/facade.py: Compressor / Decompresser / Reader / Writer with the reference's names,
Compressor.cs:12-304, Decompresser.cs:13-204) on the host: call order, error messages, framing bytes and the splicing of
the device's block into the Writer.  The GPU is replaced by a stand-in context that answers the two batch calls with
the oracle (tests only; the product never does) -- the archive must equal the oracle's / the reference text's."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend, pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402
from zpaqsharp_b200 import facade as F, libzpaq as z  # noqa: E402


class StandInContext:
    """Answers compress_blocks_model / decompress_blocks like zpaqsharp_b200.libzpaq.Context, computed by the oracle."""

    def compress_blocks_model(self, data, offsets, hdr, pcomp=b"", args=None, filename=None, comment=None, dosha1=True, with_tag=True, out=None):
        assert len(offsets) == 2
        a = po.compress_with_model(bytes(hdr), bytes(pcomp), [0] * 9, bytes(data), filename, comment if comment else str(len(data)), dosha1, with_tag)
        return np.frombuffer(a, dtype=np.uint8), np.asarray([0, len(a)], dtype=np.uint64)

    def decompress_blocks(self, archive, offsets, out=None):
        got, st = po.decompress(TAGGED(bytes(archive)))
        return np.frombuffer(got, dtype=np.uint8), np.asarray([0, len(got)], dtype=np.uint64), np.asarray([st[0] if st else 0], dtype=np.uint8), np.zeros(1, np.uint8)


def TAGGED(blk: bytes) -> bytes:
    return blk if blk.startswith(F.TAG) else F.TAG + blk


def _compress_like_reference(data, start, filename, comment, sha, tag=True, pcomp=None, chunk=-1):
    w = F.BytesWriter()
    co = F.Compressor(StandInContext())
    co.setOutput(w)
    if tag:
        co.writeTag()
    co.startBlock(start)
    co.startSegment(filename, comment)
    if pcomp is not None:
        co.postProcess(pcomp, len(pcomp))
    co.setInput(F.BytesReader(data))
    if chunk < 0:
        assert co.compress() is False
    else:
        while co.compress(chunk):
            pass
    co.endSegment(sha)
    co.endBlock()
    return w.getvalue()


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("n", [0, 1, 20000])
def test_compressor_facade_writes_the_reference_bytes(level, n):
    data = synth.blocks("mixed", 40 + level, 1, max(n, 1)).tobytes()[:n]
    sha = hashlib.sha1(data).digest()
    want = po.compress_block_level(data, level, filename="a/b.txt", comment="hello")
    assert _compress_like_reference(data, level, "a/b.txt", "hello", sha) == want
    assert _compress_like_reference(data, level, "a/b.txt", "hello", sha, chunk=777) == want
    # no comment, no checksum, no tag -- an explicit empty comment, which the batch ABI cannot express itself
    want2 = po.compress_block_level(data, level, filename=None, comment="", dosha1=False, with_tag=False)
    assert _compress_like_reference(data, level, None, None, None, tag=False) == want2


def test_compressor_facade_with_hcomp_bytes_and_pcomp():
    data = synth.blocks("mixed", 77, 1, 30000).tobytes()
    plan = frontend.plan_block("x0,5,4,0,3,19", data)                  # byte-aligned LZ77: the caller pre-processes, as with the reference
    payload = po.preprocess(data, plan["args"])
    sha = hashlib.sha1(data).digest()
    got = _compress_like_reference(payload, bytes(plan["hdr"]), None, plan["comment"], sha, pcomp=bytes(plan["pcomp"]))
    assert got == po.compress_block(data, "x0,5,4,0,3,19")


def test_compressor_facade_rules_and_errors():
    co = F.Compressor(StandInContext())
    co.setOutput(F.BytesWriter())
    with pytest.raises(z.ZpaqError, match="compression level must be at least 1"):
        co.startBlock(0)
    with pytest.raises(z.ZpaqError, match="compression level too high"):
        co.startBlock(4)
    co.startBlock(1)
    co.startSegment("f")
    co.setInput(F.BytesReader(b"abc"))
    co.compress()
    co.endSegment()
    with pytest.raises(z.ZpaqError, match="one segment per block"):
        co.startSegment("g")
    with pytest.raises(AssertionError):
        F.Compressor(StandInContext()).endBlock()                       # the reference asserts the call order (Compressor.cs:296)


def _archive():
    a, b = synth.blocks("mixed", 5, 1, 9000).tobytes(), b"second block"
    arc = b"garbage" + po.compress_block_level(a, 2, filename="one", comment="c1") + po.compress_block(b, "1", filename=None, comment=None, dosha1=False)
    return arc, a, b


def test_decompresser_facade_walks_an_archive_like_the_reference():
    arc, a, b = _archive()
    d = F.Decompresser(StandInContext())
    d.setInput(F.BytesReader(arc))
    out = F.BytesWriter()
    d.setOutput(out)
    names, comments, marks = [], [], []
    mem, mems = [], []
    while d.findBlock(mem):
        mems.append(mem[0])
        while True:
            fn = F.BytesWriter()
            if not d.findFilename(fn):
                break
            cm = F.BytesWriter()
            d.readComment(cm)
            assert d.decompress() is False
            s = bytearray(21)
            d.readSegmentEnd(s)
            names.append(fn.getvalue()); comments.append(cm.getvalue()); marks.append(bytes(s))
    assert out.getvalue() == a + b
    assert names == [b"one", b""] and comments == [b"c1", b"12"]
    assert marks[0] == b"\x01" + hashlib.sha1(a).digest() and marks[1][0] == 0
    assert len(mems) == 2 and mems[0] > 1e8 > mems[1]                    # ZPAQL.memory(): mid.cfg, then an LZ77-only block


def test_decompresser_facade_partial_reads_and_skip():
    arc, a, b = _archive()
    d = F.Decompresser(StandInContext())
    d.setInput(F.BytesReader(arc))
    out = F.BytesWriter()
    d.setOutput(out)
    assert d.findBlock() and d.findFilename()
    d.readComment()
    assert d.decompress(1000) is True and d.decompress(8000) is True and len(out.getvalue()) == 9000
    assert d.decompress(10) is False                                     # end of segment reached
    d.readSegmentEnd()
    assert d.findFilename() is False and d.findBlock() and d.findFilename()
    d.readComment()
    d.readSegmentEnd()                                                   # skipped without decoding (Decompresser.cs:168-172)
    assert d.findFilename() is False and d.findBlock() is False
    assert out.getvalue() == a


def test_decompresser_facade_errors():
    arc, a, b = _archive()
    bad = bytearray(arc)
    i = arc.index(b"zPQ")
    bad[i + 3] = 3
    d = F.Decompresser(StandInContext())
    d.setInput(F.BytesReader(bytes(bad)))
    with pytest.raises(z.ZpaqError, match="unsupported ZPAQ level"):
        d.findBlock()
