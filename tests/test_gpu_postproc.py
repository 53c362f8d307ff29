"""GPU: the native post-processors (zpq_post.cu) against the reference's way of restoring a block -- the stored PCOMP
program run by PostProcessor.write / ZPAQL.run (PostProcessor.cs:37-86, LibZPAQ.cs:427-826): on the device's own
interpreter pass (ZPQ_NATIVE_POST=0) and on the CPU oracle.  Bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# method -> what restores it
METHODS = [("x0,1,4,0,7,21,1", 2), ("x0,5,4,3,3,19,1", 2), ("x0,1,4,0,3,24", 2), ("x0,2,12,0,7,21,1c0,0,511i2m", 3),
           ("x0,6,8,0,5,18c0,0,511", 3), ("x0,2,3,5,2,17,2c0,0,511i1", 3), ("x0,3ci1", 4), ("x0,7ci1", 4), ("x2,3ci1", 4),
           ("x0,4c0,0,255", 5), ("2", 1), ("x6,1,4,0,3,24", 2), ("x6,5,4,0,3,24", 2), ("x5,3ci1", 4), ("x5,7ci1", 4)]


def _x86ish(n, seed):
    """Dense E8/E9 material: opcodes a few bytes apart, targets ending in 00 / FF, runs of E8 E8 E8 so that a rewritten
    field is itself tested as an opcode (LibZPAQ.cs:444-463)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, n, dtype=np.uint8)
    pos = 0
    while pos + 8 < n:
        pos += int(rng.integers(1, 12))
        if pos + 8 >= n:
            break
        a[pos] = 0xE8 + int(rng.integers(0, 2))
        if rng.integers(0, 3) == 0:
            a[pos + 1] = 0xE8
            a[pos + 2] = 0xE9
        a[pos + 4] = 0x00 if rng.integers(0, 2) else 0xFF
        if rng.integers(0, 4) == 0:
            a[pos + 5] = 0xFF
    return a.tobytes()


def _decode_both(ctx, arc, offs):
    """Native kernels, then the stored program compiled with NVRTC (ZPQ_NATIVE_POST=0), then the stored program interpreted
    (ZPQ_POST_NVRTC=0 as well): all three must restore the same bytes."""
    out, ooff, sha, bst = ctx.decompress_blocks(arc, offs)
    st = ctx.stats()
    os.environ["ZPQ_NATIVE_POST"] = "0"
    try:
        out2, ooff2, sha2, bst2 = ctx.decompress_blocks(arc, offs)
        st2 = ctx.stats()
        os.environ["ZPQ_POST_NVRTC"] = "0"
        out3, ooff3, sha3, bst3 = ctx.decompress_blocks(arc, offs)
        st3 = ctx.stats()
    finally:
        os.environ.pop("ZPQ_NATIVE_POST", None)
        os.environ.pop("ZPQ_POST_NVRTC", None)
    assert out.tobytes() == out2.tobytes() == out3.tobytes() and ooff.tolist() == ooff2.tolist() == ooff3.tolist()
    assert sha.tolist() == sha2.tolist() == sha3.tolist() and bst.tolist() == bst2.tolist() == bst3.tolist()
    assert st2.post_native_blocks == st3.post_native_blocks and st3.post_compiled_blocks == 0
    # without the native kernels every block that carries a program runs it as compiled code (PASS blocks are copied either way)
    assert st2.post_compiled_blocks + st2.post_native_blocks + st2.post_interpreted_blocks == len(offs) - 1
    assert st2.post_compiled_blocks >= st3.post_interpreted_blocks - st2.post_interpreted_blocks
    return out, ooff, sha, st


@pytest.mark.parametrize("method,kind", METHODS, ids=[m for m, _ in METHODS])
def test_native_postprocessors_restore_what_the_program_restores(gpu_ctx, method, kind):
    from tools import synth
    mixed = synth.blocks("mixed", 900, 1, 300000).tobytes()
    data = mixed[:90000] + _x86ish(60000, 5) + mixed[90000:150000] + b"\x00" * 5000 + b"ab" * 3000 + mixed[150000:]
    cuts = [0, 0, 1, 5, 70000, 70000 + 131072, len(data)]       # an empty block, tiny ones, ragged ones
    offs = np.asarray(cuts, dtype=np.uint64)
    arc, aoff = gpu_ctx.compress_blocks(data, offs, method)
    out, ooff, sha, st = _decode_both(gpu_ctx, arc, aoff)
    assert out.tobytes() == data and ooff.tolist() == cuts and set(sha.tolist()) == {1}
    # every block went through the native kernel of its program
    assert st.post_native_blocks == len(cuts) - 1 and st.post_interpreted_blocks == 0
    hdr, pcomp = gpu_ctx_model(method)
    if kind > 1:
        from zpaqsharp_b200 import libzpaq as z
        assert z.post_kind(hdr[4], hdr[5], pcomp) & 15 == kind


def gpu_ctx_model(method):
    from zpaqsharp_b200 import libzpaq as z
    if method in ("1", "2", "3"):
        return z.builtin_model(int(method)), b""
    text, args = z.make_config(method)
    return z.compile_config(text, args)


@pytest.mark.parametrize("method", ["x0,1,4,0,7,21,1", "x0,5,4,3,3,19,1", "x0,2,12,0,7,21,1c0,0,255", "x0,6,8,0,5,18c0,0,255",
                                    "x0,3ci1", "x0,7ci1", "x0,4c0,0,255"])
def test_damaged_and_truncated_streams_follow_the_program(gpu_ctx, oracle, method):
    """Streams no compressor writes: the transformed data cut short, with bytes flipped (matches that reach in front of the
    buffer, a BWT start index out of range ...).  Whatever the stored program makes of them is the answer: the oracle runs it
    on the CPU, the device either reproduces it natively or hands the block to its interpreter."""
    from oracle import frontend as fe
    from tools import synth
    text, args = fe.make_config(method)
    hdr, pcomp = fe.compile_config(text, args)[:2]
    plain = [0] * 9                                        # no pre-processing: `data` is stored as the transformed stream
    src = synth.blocks("mixed", 950, 1, 40000).tobytes()[:30000] + _x86ish(6000, 9)
    good = oracle.preprocess(src, args)
    rng = np.random.default_rng(11)
    bwt = method.split(",")[1][0] in "37"
    # (a BWT stream that loses its last four bytes -- the start index -- sends the program through 2^32 positions: only its body is damaged)
    streams = [good] if bwt else [good, good[:len(good) // 2], good[:len(good) - 1], good[:7], good[:1], b""]
    for k in range(8):
        bad = bytearray(good)
        for _ in range(1 + k):
            bad[int(rng.integers(0, min(len(bad) - 4, 2000 + 4000 * k)))] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(bad))
    arcs, want = [], []
    for sdata in streams:
        a = oracle.compress_with_model(bytes(hdr), bytes(pcomp), plain, sdata, dosha1=False)
        try:
            w, _ = oracle.decompress(a, cap=1 << 24)
        except Exception:
            continue                                       # the program does not terminate sanely on this one: skip it
        arcs.append(a); want.append(w)
    assert len(arcs) >= 6
    offs = np.concatenate([[0], np.cumsum([len(a) for a in arcs])]).astype(np.uint64)
    out, ooff, sha, st = _decode_both(gpu_ctx, b"".join(arcs), offs)
    for i, w in enumerate(want):
        assert out[int(ooff[i]):int(ooff[i + 1])].tobytes() == w, (method, i)
    assert st.post_native_blocks >= 1


def test_foreign_program_is_compiled_or_interpreted(gpu_ctx, oracle):
    """A PCOMP program no makeConfig emits (a foreign archive's; loops and a long jump): SURVEY 8f-1."""
    from oracle import frontend as fe
    cfg = ("comp 2 4 0 0 1 0 cm 16 255 hcomp c++ *c=a b=c a=0 hash *d=a halt "
           "pcomp foreign ; a> 255 ifnotl a^= 32 b=a c= 3 do a=b out c-- a=c a> 0 while elsel a= 33 out endif halt end")
    hdr, pcomp = fe.compile_config(cfg, [0] * 9)[:2]
    data = (b"Hello, foreign ZPAQL world! " * 300)[:5000]
    a = oracle.compress_with_model(bytes(hdr), bytes(pcomp), [0] * 9, data, dosha1=False)
    want, _ = oracle.decompress(a, cap=1 << 20)
    assert len(want) == 3 * len(data) + 1 and want[:3] == bytes([data[0] ^ 32]) * 3 and want[-1] == 33
    out, ooff, sha, bst = gpu_ctx.decompress_blocks(a, np.asarray([0, len(a)], dtype=np.uint64))
    st = gpu_ctx.stats()
    assert out.tobytes() == want
    # no native kernel knows this program: it is translated and compiled (NVRTC for sm_100a) ...
    assert st.post_native_blocks == 0 and st.post_compiled_blocks == 1 and st.post_interpreted_blocks == 0
    # ... or, without NVRTC, interpreted
    os.environ["ZPQ_POST_NVRTC"] = "0"
    try:
        out, ooff, sha, bst = gpu_ctx.decompress_blocks(a, np.asarray([0, len(a)], dtype=np.uint64))
        st = gpu_ctx.stats()
    finally:
        os.environ.pop("ZPQ_POST_NVRTC", None)
    assert out.tobytes() == want and st.post_compiled_blocks == 0 and st.post_interpreted_blocks == 1
