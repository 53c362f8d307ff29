"""GPU: the streaming facade (Compressor / Decompresser with the reference's names) through the real C ABI: the archive
equals the oracle's byte for byte and the Decompresser walks it back."""
import hashlib

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [1, 2, 3])
def test_facade_round_trip_matches_oracle(gpu_ctx, oracle, level):
    from tools import synth
    from zpaqsharp_b200 import facade as F
    data = synth.blocks("mixed", 70 + level, 1, 50000).tobytes()
    w = F.BytesWriter()
    co = F.Compressor(gpu_ctx)
    co.setOutput(w)
    for part, name in ((data[:30000], "first"), (data[30000:], None)):          # two blocks, one segment each
        co.writeTag()
        co.startBlock(level)
        co.startSegment(name, "note" if name else None)
        co.setInput(F.BytesReader(part))
        while co.compress(4096):
            pass
        co.endSegment(hashlib.sha1(part).digest())
        co.endBlock()
    arc = w.getvalue()
    assert arc == (oracle.compress_block_level(data[:30000], level, filename="first", comment="note")
                   + oracle.compress_block_level(data[30000:], level, filename=None, comment=""))
    d = F.Decompresser(gpu_ctx)
    d.setInput(F.BytesReader(arc))
    out = F.BytesWriter()
    d.setOutput(out)
    marks = []
    while d.findBlock():
        while d.findFilename():
            d.readComment()
            d.decompress()
            s = bytearray(21)
            d.readSegmentEnd(s)
            marks.append(bytes(s))
            assert d.sha1_verified() == 1
    assert out.getvalue() == data
    assert marks == [b"\x01" + hashlib.sha1(data[:30000]).digest(), b"\x01" + hashlib.sha1(data[30000:]).digest()]
