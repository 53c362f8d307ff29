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
        self.calls = getattr(self, "calls", 0) + 1
        data = bytes(data)
        arcs = []
        for i in range(len(offsets) - 1):
            b = data[int(offsets[i]):int(offsets[i + 1])]
            arcs.append(po.compress_with_model(bytes(hdr), bytes(pcomp), [0] * 9, b, filename if i == 0 else None,
                                               comment if (comment and i == 0) else str(len(b)), dosha1, with_tag))
        return np.frombuffer(b"".join(arcs), dtype=np.uint8), np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)

    def decompress_blocks(self, archive, offsets, out=None):
        self.dcalls = getattr(self, "dcalls", 0) + 1
        archive = bytes(archive)
        outs, shas = [], []
        for i in range(len(offsets) - 1):
            got, st = po.decompress(TAGGED(archive[int(offsets[i]):int(offsets[i + 1])]))
            outs.append(got); shas.append(st[0] if st else 0)
        return (np.frombuffer(b"".join(outs), dtype=np.uint8), np.concatenate([[0], np.cumsum([len(o) for o in outs])]).astype(np.uint64),
                np.asarray(shas, dtype=np.uint8), np.zeros(len(outs), np.uint8))


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


def test_compressor_queues_blocks_and_codes_them_in_one_call():
    """Finished blocks wait for a wave (Compressor.cs:193-248 on a device that only pays off in batches): nothing a later block
    writes may overtake them, one batch call codes them all, and the archive is what block-at-a-time coding gives."""
    ctx = StandInContext()
    w = F.BytesWriter()
    parts = [synth.blocks("mixed", 80 + i, 1, 3000 + 500 * i).tobytes() for i in range(5)]
    with F.Compressor(ctx, batch_blocks=4) as co:
        co.setOutput(w)
        for i, part in enumerate(parts):
            co.writeTag()
            co.startBlock(1 if i != 3 else 2)                      # block 3 uses another model: its own batch call
            co.startSegment("f%d" % i, None)
            co.setInput(F.BytesReader(part))
            co.compress()
            if i % 2:
                co.endSegment(hashlib.sha1(part).digest())
            else:
                co.setVerify(True)
                sha, size = co.endSegmentChecksum()
                assert sha == hashlib.sha1(part).digest() and size == len(part)
            co.endBlock()
            if i == 0:
                held = len(w.buf)                                    # tag, block header and segment header of block 0 went straight through
            if i < 3:
                assert len(w.buf) == held and getattr(ctx, "calls", 0) == 0   # everything behind them is still queued
            if i == 3:
                # the fourth pending block starts the wave: two models -> two batch calls (hdr/pcomp/checksum mode group the blocks)
                assert ctx.calls >= 2 and len(w.buf) > 0
    want = b"".join(po.compress_block_level(p, 1 if i != 3 else 2, filename="f%d" % i, comment="") for i, p in enumerate(parts))
    assert w.getvalue() == want


def test_reading_the_writer_flushes_and_batch_one_is_immediate():
    part = synth.blocks("text", 90, 1, 5000).tobytes()
    for batch in (None, 1):
        ctx = StandInContext()
        w = F.BytesWriter()
        co = F.Compressor(ctx, batch_blocks=batch)
        co.setOutput(w)
        co.startBlock(1)
        co.startSegment(None, None)
        co.setInput(F.BytesReader(part))
        co.compress()
        co.endSegment()
        co.endBlock()
        if batch == 1:
            assert len(w.buf) > 100                                  # coded when the block ended
        got = w.getvalue()                                           # reading the Writer codes what is queued
        assert got == po.compress_block_level(part, 1, comment="", dosha1=False, with_tag=False)


def test_decompresser_decodes_a_wave_ahead():
    parts = [synth.blocks("mixed", 95 + i, 1, 2000 + 100 * i).tobytes() for i in range(6)]
    arc = b"".join(po.compress_block_level(p, 1) for p in parts)
    ctx = StandInContext()
    d = F.Decompresser(ctx, batch_blocks=4)
    d.setInput(F.BytesReader(arc))
    got = []
    while d.findBlock():
        while d.findFilename():
            d.readComment()
            out = F.BytesWriter()
            d.setOutput(out)
            d.decompress()
            d.readSegmentEnd()
            assert d.sha1_verified() == 1
            got.append(out.getvalue())
    assert got == parts
    assert ctx.dcalls == 2                                           # 6 blocks in waves of 4
