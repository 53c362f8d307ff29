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


def test_facade_batches_64_blocks_like_the_batch_abi(gpu_ctx, oracle):
    """64 blocks through the streaming classes cost about what one batch call costs (they are queued and coded as a wave),
    and the archive is byte for byte the batch ABI's / the oracle's."""
    import time
    import numpy as np
    from tools import synth
    from zpaqsharp_b200 import facade as F
    nb, size = 64, 100000
    data = synth.blocks("mixed", 300, nb, size)
    offs = np.arange(0, (nb + 1) * size, size, dtype=np.uint64)
    gpu_ctx.compress_blocks_level(data, offs, 2, comment=None)                 # warm-up (allocations)
    t0 = time.perf_counter()
    arc, ooff = gpu_ctx.compress_blocks_level(data, offs, 2)
    t_batch = time.perf_counter() - t0
    w = F.BytesWriter()
    t0 = time.perf_counter()
    with F.Compressor(gpu_ctx) as co:
        co.setOutput(w)
        for i in range(nb):
            part = data[i * size:(i + 1) * size].tobytes()
            co.writeTag()
            co.startBlock(2)
            co.startSegment(None, str(size))
            co.setInput(F.BytesReader(part))
            co.compress()
            co.setVerify(True)
            co.endSegmentChecksum()
            co.endBlock()
    t_facade = time.perf_counter() - t0
    got = w.getvalue()
    assert got == arc.tobytes()
    assert got[:int(ooff[1])] == oracle.compress_block_level(data[:size].tobytes(), 2)
    assert t_facade < 1.5 * t_batch + 0.5, (t_facade, t_batch)
    # and back: one GPU call decodes the wave
    d = F.Decompresser(gpu_ctx)
    d.setInput(F.BytesReader(got))
    out = F.BytesWriter()
    d.setOutput(out)
    t0 = time.perf_counter()
    n = 0
    while d.findBlock():
        while d.findFilename():
            d.readComment()
            d.decompress()
            d.readSegmentEnd()
            assert d.sha1_verified() == 1
            n += 1
    t_dec = time.perf_counter() - t0
    assert n == nb and out.getvalue() == data.tobytes()
    t0 = time.perf_counter()
    gpu_ctx.decompress_blocks(arc, ooff)
    t_dbatch = time.perf_counter() - t0
    assert t_dec < 1.5 * t_dbatch + 0.5, (t_dec, t_dbatch)
