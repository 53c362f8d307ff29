"""GPU: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Bit-exact for everything: archives byte-identical, decompression identical to the input."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVES = json.load(open(os.path.join(HERE, "golden", "oracle_archives.json")))
DEVICE_PREPROC = set()   # every pre-processor (E8E9, LZ77 hash / suffix array, BWT) runs on the device


def _input(case):
    from tools import synth
    n = case["nbytes"]
    return synth.blocks(case["kind"], case["first_block"], 1, max(n, 1)).tobytes()[:n] if n else b""


@pytest.mark.parametrize("case", ARCHIVES, ids=[c["name"] for c in ARCHIVES])
def test_decode_golden_archives(gpu_ctx, case):
    arc = base64.b64decode(case["archive_b64"])
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(arc, np.asarray([0, len(arc)], dtype=np.uint64))
    data = _input(case)
    assert out.tobytes() == data
    assert sha.tolist() == [1] and bst.tolist() == [0]


@pytest.mark.parametrize("case", [c for c in ARCHIVES if c["name"] not in DEVICE_PREPROC],
                         ids=[c["name"] for c in ARCHIVES if c["name"] not in DEVICE_PREPROC])
def test_encode_matches_golden_archives(gpu_ctx, case):
    data = _input(case)
    offs = np.asarray([0, len(data)], dtype=np.uint64)
    if case["how"] == "level":
        arc, _ = gpu_ctx.compress_blocks_level(data, offs, case["arg"])
    else:
        arc, _ = gpu_ctx.compress_blocks(data, offs, case["arg"])
    assert hashlib.sha1(arc.tobytes()).hexdigest() == case["archive_sha1"]


@pytest.mark.parametrize("method", ["1", "2", "3", "30,128,1", "11,60,0", "11,100,2", "21,40,0", "x0,5,4,0,3,19", "x0,6,8,0,5,18c0,0,511",
                                    "x0,7ci1", "x0,3ci1", "x0,1,4,2,3,16,1", "x0,2,3,5,2,17,2c0,0,511i1", "x0,1,4,0,7,21,2",
                                    "x0,2,12,0,7,21,1c0,0,511i2", "x0,2,1,0,2,15", "x0,1,6,0,1,12"])
def test_lz77_bwt_preprocessing_matches_oracle(gpu_ctx, oracle, method):
    # LZBuffer.cs:151-486 on the device: hash-table and suffix-array matchers, both code formats, BWT, +E8E9
    from tools import synth
    data = synth.blocks("mixed", 600, 1, 200000).tobytes()
    cuts = [0, 70000, 70001, 70001, 200000]
    offs = np.asarray(cuts, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks(data, offs, method)
    ref = b"".join(oracle.compress_block(data[cuts[i]:cuts[i + 1]], method) for i in range(len(cuts) - 1))
    assert arc.tobytes() == ref
    out, _, sha, _ = gpu_ctx.decompress_blocks(arc, ooff)
    assert out.tobytes() == data and set(sha.tolist()) == {1}


@pytest.mark.parametrize("kind,method", [("text", "2"), ("text", "3"), ("mixed", "1"), ("text", "30,128,1")])
def test_preprocessing_full_size_blocks(gpu_ctx, oracle, kind, method):
    from tools import synth
    nb = 3
    data = synth.blocks(kind, 700, nb, synth.BLOCK_1MB)
    offs = np.arange(0, (nb + 1) * synth.BLOCK_1MB, synth.BLOCK_1MB, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks(data, offs, method)
    ref0 = oracle.compress_block(data[:synth.BLOCK_1MB].tobytes(), method)
    assert arc[:int(ooff[1])].tobytes() == ref0
    out, _, sha, _ = gpu_ctx.decompress_blocks(arc, ooff)
    assert np.array_equal(out, data) and set(sha.tolist()) == {1}


@pytest.mark.parametrize("level", [1, 2, 3])
def test_builtin_levels_ragged_batch_matches_oracle(gpu_ctx, oracle, level):
    from tools import synth
    data = synth.blocks("mixed", 100 + level, 1, 150000).tobytes()
    cuts = [0, 1, 1, 40000, 40007, 90000, 150000]      # includes an empty block and tiny ones
    offs = np.asarray(cuts, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks_level(data, offs, level, filename="a/b.bin", comment="note")
    ref = b"".join(oracle.compress_block_level(data[cuts[i]:cuts[i + 1]], level,
                                               filename="a/b.bin" if i == 0 else None,
                                               comment="note" if i == 0 else None)
                   for i in range(len(cuts) - 1))
    assert arc.tobytes() == ref
    out, o2, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert out.tobytes() == data
    assert o2.tolist() == cuts and set(sha.tolist()) == {1} and not bst.any()


@pytest.mark.parametrize("method", ["x0,0c256,0,255,255", "s0,0c0,0,255i2", "x0,0ci1,1,1,1,2am", "4", "5", "60,200,3",
                                    "x0,0c1,0,255,255a24mm16ts19t0w2", "x0,4ci1,1,1,1,2awm", "0",
                                    "x0,0c0,1003,255c0,7c0,0,1300,255c200,0,511,300a24,1,1m12,20s9,20,100t3"])
def test_methods_match_oracle(gpu_ctx, oracle, method):
    from tools import synth
    data = synth.blocks("mixed", 200, 1, 120000).tobytes()
    offs = np.asarray([0, 50000, 120000], dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks(data, offs, method, filename="x", comment="c")
    ref = oracle.compress_block(data[:50000], method, "x", "c") + oracle.compress_block(data[50000:], method)
    assert arc.tobytes() == ref
    out, _, sha, _ = gpu_ctx.decompress_blocks(arc, ooff)
    assert out.tobytes() == data and set(sha.tolist()) == {1}


@pytest.mark.parametrize("method", ["1", "2", "3", "30,128,1", "x0,5,4,0,3,19", "x0,6,8,0,5,18c0,0,511", "x0,7ci1",
                                    "x0,1,4,2,3,16,1", "x0,2,3,5,2,17,2c0,0,511i1", "x4,3ci1", "x5,3ci1"])
def test_decode_oracle_archives_with_postprocessing(gpu_ctx, oracle, method):
    # PostProcessor.cs:37-86 with the PCOMP programs of LibZPAQ.cs:427-830 interpreted on the device
    from tools import synth
    data = synth.blocks("mixed", 300, 1, 100000).tobytes()
    a = oracle.compress_block(data[:60000], method)
    b = oracle.compress_block(data[60000:], method)
    arc = a + b
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(arc, np.asarray([0, len(a), len(arc)], dtype=np.uint64))
    assert out.tobytes() == data and sha.tolist() == [1, 1]


def test_generic_kernel_path_agrees(zlib_, oracle):
    # the step-scheduled kernels (used for models with more than 32 components) on a small model
    from tools import synth
    os.environ["ZPQ_FORCE_GENERIC"] = "1"
    try:
        data = synth.blocks("text", 400, 1, 30000).tobytes()
        with zlib_.Context() as ctx:
            offs = np.asarray([0, len(data)], dtype=np.uint64)
            arc, ooff = ctx.compress_blocks_level(data, offs, 3)
            assert arc.tobytes() == oracle.compress_block_level(data, 3)
            out, _, sha, _ = ctx.decompress_blocks(arc, ooff)
            assert out.tobytes() == data
    finally:
        os.environ.pop("ZPQ_FORCE_GENERIC", None)


def test_more_than_32_components(gpu_ctx, oracle):
    from tools import synth
    method = "x0,0" + "c0,0,255" * 20 + "i1,1,1,1,1,1,1,1,1,1,1,1,1,1,1m"       # 36 components
    data = synth.blocks("text", 401, 1, 20000).tobytes()
    arc, ooff = gpu_ctx.compress_blocks(data, np.asarray([0, len(data)], dtype=np.uint64), method)
    assert arc.tobytes() == oracle.compress_block(data, method)
    out, _, sha, _ = gpu_ctx.decompress_blocks(arc, ooff)
    assert out.tobytes() == data and sha.tolist() == [1]


def test_corrupt_block_is_isolated(gpu_ctx, zlib_, oracle):
    # Decoder.cs:141 "archive corrupted": one bad block must not kill the batch
    from tools import synth
    data = synth.blocks("text", 402, 1, 40000).tobytes()
    a = bytearray(oracle.compress_block_level(data[:20000], 1))
    b = oracle.compress_block_level(data[20000:], 1)
    a[len(a) // 2] ^= 0x55
    arc = bytes(a) + b
    with pytest.raises(zlib_.ZpaqError) as e:
        gpu_ctx.decompress_blocks(arc, np.asarray([0, len(a), len(arc)], dtype=np.uint64))
    assert e.value.code == zlib_.E_CORRUPT
    assert e.value.block_status[0] != 0 and e.value.block_status[1] == 0


def _cpu_archives(blocks, fn):
    """Archive blocks of the CPU oracle, one block per host thread (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex:
        return list(ex.map(fn, blocks))


def test_full_size_blocks_every_block_against_the_oracle(gpu_ctx, oracle):
    # BASELINE size: 24 blocks of 1,044,480 bytes, mid.cfg: EVERY archive block byte for byte against the oracle
    # (and through it the reference's Compressor text, tests/test_reference_compressor.py), then the round trip with the
    # stored SHA-1 verified on the device
    from tools import synth
    nb = 24
    bs = synth.BLOCK_1MB
    data = synth.blocks("mixed", 0, nb, bs)
    offs = np.arange(0, (nb + 1) * bs, bs, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks_level(data, offs, 2)
    ref = _cpu_archives([data[i * bs:(i + 1) * bs].tobytes() for i in range(nb)], lambda b: oracle.compress_block_level(b, 2))
    for i in range(nb):
        assert arc[int(ooff[i]):int(ooff[i + 1])].tobytes() == ref[i], i
    out, o2, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert np.array_equal(out, data) and set(sha.tolist()) == {1} and not bst.any()
    # the device also decodes what the CPU wrote, given as one archive
    cat = b"".join(ref)
    roff = np.concatenate([[0], np.cumsum([len(r) for r in ref])]).astype(np.uint64)
    out2, _, sha2, _ = gpu_ctx.decompress_blocks(cat, roff)
    assert np.array_equal(out2, data) and set(sha2.tolist()) == {1}


def test_max_cfg_full_size_blocks_against_the_oracle(gpu_ctx, oracle):
    # BASELINE configs[3]: max.cfg, 22 components, 1,044,480-byte blocks
    from tools import synth
    nb = 4
    bs = synth.BLOCK_1MB
    data = synth.blocks("mixed", 40, nb, bs)
    offs = np.arange(0, (nb + 1) * bs, bs, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks_level(data, offs, 3)
    ref = _cpu_archives([data[i * bs:(i + 1) * bs].tobytes() for i in range(nb)], lambda b: oracle.compress_block_level(b, 3))
    for i in range(nb):
        assert arc[int(ooff[i]):int(ooff[i + 1])].tobytes() == ref[i], i
    out, _, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert np.array_equal(out, data) and set(sha.tolist()) == {1} and not bst.any()


@pytest.mark.parametrize("size,method,kind,nb", [
    (4190208, "32,128,1", "text", 3),                         # BASELINE configs[2]: BWT, arg0 = 2
    (4190208, "x2,2,12,0,7,23,1c0,0,511i2m", "mixed", 2),    # C5: byte LZ77 + CM chain at 4 MB
    (16773120, "x4,3ci1", "text", 2),                         # C5: BWT at 16 MB, arg0 = 4 (other checkbits / table sizes)
    (16773120, "x4,1,4,0,7,25,1", "mixed", 2),                # C5: bit-packed LZ77 at 16 MB, stored
])
def test_large_blocks_against_the_oracle(gpu_ctx, oracle, size, method, kind, nb):
    from tools import synth
    data = synth.blocks(kind, 60, nb, size)
    offs = np.arange(0, (nb + 1) * size, size, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks(data, offs, method)
    ref = _cpu_archives([data[i * size:(i + 1) * size].tobytes() for i in range(nb)], lambda b: oracle.compress_block(b, method))
    for i in range(nb):
        assert arc[int(ooff[i]):int(ooff[i + 1])].tobytes() == ref[i], i
    out, _, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert np.array_equal(out, data) and set(sha.tolist()) == {1} and not bst.any()
    assert gpu_ctx.stats().post_native_blocks == nb


def test_mid_cfg_4mb_blocks_against_the_oracle(gpu_ctx, oracle):
    # C5: the mid.cfg quarter of the mixed archives at 4,190,208 bytes (a block is coded at the speed of its own bit chain: a
    # 16,773,120-byte mid.cfg block takes ~40 s to code and ~90 s to decode on any number of SMs, so that size is round-tripped by
    # `bench.py --config C5` and not here)
    from tools import synth
    size = 4190208
    data = synth.blocks("mixed", 70, 2, size)
    offs = np.arange(0, 3 * size, size, dtype=np.uint64)
    arc, ooff = gpu_ctx.compress_blocks_level(data, offs, 2)
    ref = _cpu_archives([data[i * size:(i + 1) * size].tobytes() for i in range(2)], lambda b: oracle.compress_block_level(b, 2))
    for i in range(2):
        assert arc[int(ooff[i]):int(ooff[i + 1])].tobytes() == ref[i], i
    out, _, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert np.array_equal(out, data) and set(sha.tolist()) == {1} and not bst.any()


@pytest.mark.parametrize("method", ["x0,0c256,0,255,255", "x0,3ci1", "x0,2,12,0,7,21,1c0,0,511i2m"])
def test_explicit_methods_run_nvrtc_kernels(zlib_, method):
    # a header that is not one of the three built-in models is specialised at run time (NVRTC for sm_100a); a silent
    # fall-back to the run-time model walker would show here (zpq_stats.kernel names what ran, and why not)
    from tools import synth
    data = synth.blocks("mixed", 800, 1, 50000).tobytes()
    offs = np.asarray([0, 20000, 50000], dtype=np.uint64)
    with zlib_.Context() as ctx:
        arc, ooff = ctx.compress_blocks(data, offs, method)
        k_enc = ctx.stats().kernel.decode()
        out, _, sha, _ = ctx.decompress_blocks(arc, ooff)
        k_dec = ctx.stats().kernel.decode()
    assert "nvrtc" in k_enc, k_enc
    assert "nvrtc" in k_dec, k_dec
    assert out.tobytes() == data and set(sha.tolist()) == {1}


def test_one_call_api(zlib_, oracle):
    from tools import synth
    data = synth.blocks("text", 500, 1, 70000).tobytes()
    arc = zlib_.compressBlock(data, "x0,0c0,0,255i1", "n.txt", "cmt")
    assert arc == oracle.compress_block(data, "x0,0c0,0,255i1", "n.txt", "cmt")
    assert zlib_.decompress(arc + arc) == data + data


# ---- the encoder generations must agree byte for byte, and the fast one must be the one that runs ----
def _encode_with(zlib_, env, data, offs, level=None, method=None):
    old = {k: os.environ.get(k) for k in ("ZPQ_DUO", "ZPQ_PIPE")}
    os.environ.update(env)
    try:
        with zlib_.Context() as ctx:
            arc, ooff = (ctx.compress_blocks_level(data, offs, level) if level else ctx.compress_blocks(data, offs, method))
            return arc.tobytes(), ooff.tolist(), ctx.stats().kernel.decode()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("what", [("level", 1), ("level", 2), ("level", 3), ("method", "x0,0c256,0,255,255"),
                                  ("method", "x0,4ci1,1,1,1,2awm"), ("method", "x0,0c1,0,255,255a24mm16ts19t0w2")])
def test_encoder_generations_agree(zlib_, oracle, what):
    # role-split encoder (zpq_duo.cuh) == time-skewed encoder (zpq_pipe.cuh) == bit-by-bit lane encoder == oracle,
    # on a ragged batch that makes the blocks of one warp start and finish at different times
    from tools import synth
    data = synth.blocks("mixed", 500, 1, 260000).tobytes()
    cuts = [0, 0, 3, 20000, 20001, 75000, 140000, 141000, 260000]
    offs = np.asarray(cuts, dtype=np.uint64)
    kw = {"level": what[1]} if what[0] == "level" else {"method": what[1]}
    duo, off_d, k_d = _encode_with(zlib_, {"ZPQ_DUO": "1", "ZPQ_PIPE": "1"}, data, offs, **kw)
    pipe, off_p, k_p = _encode_with(zlib_, {"ZPQ_DUO": "0", "ZPQ_PIPE": "1"}, data, offs, **kw)
    lanes, off_l, k_l = _encode_with(zlib_, {"ZPQ_DUO": "0", "ZPQ_PIPE": "0"}, data, offs, **kw)
    assert k_d.startswith("duo/"), k_d          # the role-split kernel really ran
    assert "time-skewed" in k_p and not k_l.startswith("duo/") and "time-skewed" not in k_l
    assert duo == pipe == lanes and off_d == off_p == off_l
    ref = b"".join((oracle.compress_block_level(data[cuts[i]:cuts[i + 1]], what[1]) if what[0] == "level"
                    else oracle.compress_block(data[cuts[i]:cuts[i + 1]], what[1])) for i in range(len(cuts) - 1))
    assert duo == ref


def test_role_split_encoder_many_blocks_per_sm(gpu_ctx, oracle):
    # more blocks than one wave of warps holds, so every group takes several jobs from the queue
    from tools import synth
    nblk, size = 700, 6000
    data = synth.blocks("text", 900, nblk, size).tobytes()
    offs = np.arange(0, nblk * size + 1, size, dtype=np.uint64)
    gpu_ctx.set_max_resident(64)
    try:
        arc, ooff = gpu_ctx.compress_blocks_level(data, offs, 2)
        assert gpu_ctx.stats().resident_blocks <= 64
    finally:
        gpu_ctx.set_max_resident(0)
    a = arc.tobytes()
    for i in (0, 1, 63, 64, 65, 333, 698, 699):
        assert a[int(ooff[i]):int(ooff[i + 1])] == oracle.compress_block_level(data[i * size:(i + 1) * size], 2), i
    out, _, sha, bst = gpu_ctx.decompress_blocks(arc, ooff)
    assert out.tobytes() == data and set(sha.tolist()) == {1}


def test_level5_analysis_on_the_device_matches_oracle(gpu_ctx, oracle, zlib_):
    # LibZPAQ.cs:242-280: levels 5..9 pick periodic models from the byte-gap histogram of each block; the histogram is
    # counted on the device (k_gap_hist), the resulting method -- and so the archive -- must be the oracle's
    from oracle import frontend as fe
    from tools import synth
    rng = np.random.default_rng(3)
    rec37 = (bytes(range(37)) * 3000)[:90000]                                       # period 37
    rec300 = bytes(rng.integers(0, 256, 300, dtype=np.uint8)) * 250                # period 300 (> 255: no second model)
    mixed = synth.blocks("mixed", 650, 1, 70000).tobytes()
    tiny = b"abcabcabc"
    parts = [rec37, mixed, rec300, tiny, b"", rec37[:5000] + mixed[:5000]]
    data = b"".join(parts)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint64)
    for method in ("5", "58,200,1"):
        want_methods = [fe.expand_method(method, p) for p in parts]
        assert "c0,0,1036,255i1" in want_methods[0] and "c0,0,1299,255i1" in want_methods[2]
        assert [zlib_.expand_method(method, p) for p in parts] == want_methods
        arc, ooff = gpu_ctx.compress_blocks(data, offs, method)
        ref = b"".join(oracle.compress_block(p, method) for p in parts)
        assert arc.tobytes() == ref
        out, _, sha, _ = gpu_ctx.decompress_blocks(arc, ooff)
        assert out.tobytes() == data and set(sha.tolist()) == {1}


# ---- blocks with several segments (Compressor.cs:133-146; Decompresser.cs:128-134: decoder and post-processor start once per block) ----
def _model(method):
    from oracle import frontend as fe
    if method in (1, 2, 3):
        return bytes(fe.builtin_model(method)[0]), b"", [0] * 9
    text, args = fe.make_config(method)
    hdr, pcomp = fe.compile_config(text, args)[:2]
    return bytes(hdr), bytes(pcomp), list(args)


@pytest.mark.parametrize("method", [1, 2, 3, "x0,0c256,0,255,255", "x0,4c0,0,255", "x0,2,12,0,7,21,1c0,0,255", "x0,1,4,0,7,21,1", "x0,5,4,0,3,19"])
def test_multi_segment_blocks_decode_and_verify_every_checksum(gpu_ctx, oracle, method):
    from tools import synth
    hdr, pcomp, args = _model(method)
    data = synth.blocks("mixed", 1300, 1, 60000).tobytes()
    cuts = [0, 9000, 9000, 9001, 30000, 60000]                 # an empty and a one-byte segment among them
    parts = [data[cuts[k]:cuts[k + 1]] for k in range(len(cuts) - 1)]
    # a segment carries its own transformed data (the programs restart at every end of segment, LibZPAQ.cs:465, :599, :805-812)
    stream = [oracle.preprocess(p, args) if pcomp else p for p in parts]
    cat = b"".join(stream)
    scuts = np.concatenate([[0], np.cumsum([len(x) for x in stream])]).tolist()
    arc = oracle.compress_segments(hdr, pcomp, cat, scuts, dosha1=False)
    want, _ = oracle.decompress(arc, cap=1 << 20)
    assert want == data
    two = arc + oracle.compress_segments(hdr, pcomp, cat[:scuts[2]], scuts[:3], dosha1=False)
    # (the comments of these segments give the size of the TRANSFORMED data: the caller sizes the output itself)
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(two, np.asarray([0, len(arc), len(two)], dtype=np.uint64), out=np.empty(1 << 20, dtype=np.uint8))
    assert out.tobytes() == data + data[:cuts[2]] and ooff.tolist() == [0, len(data), len(data) + cuts[2]] and not bst.any()
    assert sha.tolist() == [0, 0]                               # nothing stored
    if not pcomp:
        # with checksums (they cover the segment's own bytes): all verified on the device; one damaged checksum is found
        arc = oracle.compress_segments(hdr, pcomp, cat, scuts, dosha1=True)
        out, ooff, sha, bst = gpu_ctx.decompress_blocks(arc, np.asarray([0, len(arc)], dtype=np.uint64))
        assert out.tobytes() == data and sha.tolist() == [1]
        bad = bytearray(arc)
        bad[-3] ^= 1                                           # inside the last segment's stored SHA-1
        out, ooff, sha, bst = gpu_ctx.decompress_blocks(bytes(bad), np.asarray([0, len(bad)], dtype=np.uint64))
        assert out.tobytes() == data and sha.tolist() == [2]
        k = arc.index(b"\x00\x00\x00\x00\xfd") + 7              # inside the FIRST segment's stored SHA-1
        bad = bytearray(arc); bad[k] ^= 1
        out, ooff, sha, bst = gpu_ctx.decompress_blocks(bytes(bad), np.asarray([0, len(bad)], dtype=np.uint64))
        assert out.tobytes() == data and sha.tolist() == [2]


def test_many_two_segment_blocks_cover_every_coder_state_at_a_segment_end(gpu_ctx, oracle):
    """The arithmetic coder is initialised once per block (Encoder.init in startBlock, Decoder.init for the first segment only): after
    the end-of-segment flag `low` is 0x100, 0x10000 or 0x1000000 instead of 1 about once in 256 segment ends, and a decoder that
    started every segment afresh would lose the block there.  1500 small two-segment blocks meet that case with probability 0.997."""
    from tools import synth
    hdr, pcomp, args = _model(1)
    nb = 1500
    data = synth.blocks("mixed", 1400, 1, nb * 160).tobytes()
    arcs = []
    for i in range(nb):
        part = data[i * 160:(i + 1) * 160]
        arcs.append(oracle.compress_segments(hdr, b"", part, [0, 40 + i % 80, 160]))
    offs = np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(b"".join(arcs), offs)
    assert out.tobytes() == data and set(sha.tolist()) == {1} and not bst.any()


def test_mixed_method_archive_in_one_call(gpu_ctx, oracle, zlib_):
    """blocks of different models interleaved in one batch.  The groups (one per model header) are decoded side by
    side on their own streams; results come back in block order; a damaged block takes down neither its group nor the others."""
    from tools import synth
    methods = [2, "x0,1,4,0,7,21,1", "x0,2,12,0,7,21,1c0,0,511i2m", "x0,3ci1", 1, "x0,0c256,0,255,255"]
    parts, arcs = [], []
    for i in range(18):
        m = methods[i % len(methods)]
        d = synth.blocks("mixed" if i % 2 else "text", 1500 + i, 1, 20000 + 1000 * i).tobytes()
        parts.append(d)
        arcs.append(oracle.compress_block_level(d, m) if isinstance(m, int) else oracle.compress_block(d, m))
    offs = np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(b"".join(arcs), offs)
    assert out.tobytes() == b"".join(parts) and set(sha.tolist()) == {1} and not bst.any()
    assert ooff.tolist() == np.concatenate([[0], np.cumsum([len(p) for p in parts])]).tolist()
    bad = [bytearray(a) for a in arcs]
    bad[6][len(bad[6]) // 2] ^= 0x77                            # a mid.cfg block
    with pytest.raises(zlib_.ZpaqError) as e:
        gpu_ctx.decompress_blocks(b"".join(bytes(a) for a in bad), offs)
    assert e.value.code == zlib_.E_CORRUPT and [i for i, s in enumerate(e.value.block_status) if s] == [6]
    # the library's own compressor on the same ragged mix of sizes: one call per method, archives equal the oracle's
    for m in methods[1:4]:
        idx = [i for i in range(18) if methods[i % len(methods)] == m]
        data = b"".join(parts[i] for i in idx)
        o2 = np.concatenate([[0], np.cumsum([len(parts[i]) for i in idx])]).astype(np.uint64)
        arc, _ = gpu_ctx.compress_blocks(data, o2, m)
        assert arc.tobytes() == b"".join(arcs[i] for i in idx)
