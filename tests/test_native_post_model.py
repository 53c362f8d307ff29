"""CPU: the algorithms of the native post-processors (tests/native_post_model.py, a line-by-line model of zpq_post.cu) against the
stored PCOMP programs run by the oracle's ZPAQL machine -- on many more valid, truncated and damaged streams than the GPU tests
can afford.  Whenever the native form does not hand the block back (ODD), it must write exactly what the program writes."""
import multiprocessing as mp
import zlib

import numpy as np
import pytest

from native_post_model import ODD, OK, restore

METHODS = ["x0,1,4,0,7,21,1", "x0,5,4,3,3,19,1", "x6,1,4,0,3,24", "x0,2,12,0,7,21,1c0,0,255", "x0,6,8,0,5,18c0,0,255", "x0,2,3,5,2,17,2c0,0,255",
           "x0,3ci1", "x0,7ci1", "x0,4c0,0,255"]


def _x86ish(n, rng):
    a = rng.integers(0, 256, n, dtype=np.uint8)
    pos = 0
    while pos + 8 < n:
        pos += int(rng.integers(1, 12))
        if pos + 8 >= n:
            break
        a[pos] = 0xE8 + int(rng.integers(0, 2))
        if rng.integers(0, 3) == 0:
            a[pos + 1] = 0xE8
        a[pos + 4] = 0x00 if rng.integers(0, 2) else 0xFF
    return a.tobytes()


def _program_output(args):
    """(in a worker process: a stream on which the program does not terminate sanely must not hang the test run)"""
    shdr, stream = args
    from oracle import pyoracle as po
    try:
        out, _ = po.zpaql_run(shdr, stream, pp=True, eof_call=True, nh=1)
        return out
    except Exception:
        return None


@pytest.mark.parametrize("method", METHODS)
def test_native_algorithms_write_what_the_program_writes(oracle, zlib_, method):
    from oracle import frontend as fe
    from tools import synth
    text, args = fe.make_config(method)
    hdr, pcomp = fe.compile_config(text, args)[:2]
    ph, pm = hdr[4], hdr[5]
    k = zlib_.post_kind(ph, pm, bytes(pcomp))
    kind, e8, param = k & 15, (k >> 4) & 1, k >> 8
    assert kind in (2, 3, 4, 5)
    shdr = bytes([0, 0, ph, pm, 0, 0]) + bytes(pcomp)
    shdr = bytes([len(shdr) & 255, len(shdr) >> 8]) + shdr
    rng = np.random.default_rng(zlib.crc32(method.encode()))
    bwt = kind == 4
    streams = []
    for t in range(100 if not bwt else 40):
        n = int(rng.integers(1, 2500))
        src = (synth.blocks("mixed", 3000 + t, 1, 4000).tobytes()[:n] if t % 3 else _x86ish(n, rng)) if t % 7 else bytes(n)
        good = oracle.preprocess(src, list(args))
        streams.append(good)
        if bwt:                                        # (a BWT stream without its start index walks 2^32 positions: only the body is damaged)
            for _ in range(3):
                bad = bytearray(good)
                for _ in range(int(rng.integers(1, 4))):
                    if len(bad) > 5:
                        bad[int(rng.integers(0, len(bad) - 4))] ^= 1 << int(rng.integers(0, 8))
                streams.append(bytes(bad))
            continue
        streams.append(good[:int(rng.integers(0, len(good) + 1))])
        for _ in range(4):
            bad = bytearray(good)
            for _ in range(int(rng.integers(1, 5))):
                if bad:
                    bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            if rng.integers(0, 3) == 0 and bad:
                cut = int(rng.integers(0, len(bad)))
                bad = bad[:cut] + bytes(rng.integers(0, 256, int(rng.integers(0, 6)), dtype=np.uint8)) + bad[cut:]
            streams.append(bytes(bad))
    # the native form first: where it hands the block back nothing is claimed; where it answers, the program is asked
    todo = []
    for s in streams:
        rc, out = restore(kind, e8, param, ph, pm, s)
        if rc == OK:
            todo.append((s, out))
    assert len(todo) >= len(streams) // 4, "the native form handed back nearly everything"
    with mp.get_context("spawn").Pool(2) as pool:
        jobs = [pool.apply_async(_program_output, ((shdr, s),)) for s, _ in todo]
        for (s, out), j in zip(todo, jobs):
            try:
                want = j.get(timeout=120)
            except mp.TimeoutError:
                pytest.fail("the program did not finish on a stream the native form accepted (%d bytes)" % len(s))
            assert want is not None and want == out, (method, len(s), len(out), None if want is None else len(want))


def test_bwt_sub_list_walk_equals_the_sequential_walk(oracle):
    """The parallel list ranking of k_post_bwt (sub-lists between multiples of a stride) emits what the sequential traversal of the
    program emits -- for real BWT streams of many sizes (smaller than one stride, start index a multiple of the stride, ...) and
    for damaged ones whose list is shorter than the block."""
    from native_post_model import bwt_list, bwt_segment, bwt_walk_split
    from tools import synth
    rng = np.random.default_rng(99)
    args = [0, 3, 0, 0, 0, 0, 0, 0, 0]
    sizes = [1, 2, 5, 63, 64, 65, 127, 128, 129, 1000, 4096, 9000, 70000]
    for t, n in enumerate(sizes):
        src = synth.blocks("text" if t % 2 else "mixed", 5000 + t, 1, max(n, 16)).tobytes()[:n]
        good = oracle.preprocess(src, args)
        streams = [good]
        for _ in range(3):
            bad = bytearray(good)
            if len(bad) > 6:
                bad[int(rng.integers(0, len(bad) - 4))] ^= 1 << int(rng.integers(0, 8))
            streams.append(bytes(bad))
        for s in streams:
            rc, want = bwt_segment(s, 24)
            if rc != OK:
                continue
            size = len(s) - 4
            idx = int.from_bytes(s[size:size + 4], "little")
            for max_split in (2048, 4, 1):
                assert bwt_walk_split(bwt_list(s, size, idx), idx, size, s, max_split) == bytes(want), (n, max_split)
        assert bwt_segment(good, 24)[1] == src
    # a start index that is a multiple of the stride: no sub-list ends there (idx is not in the image of the list)
    s = oracle.preprocess(bytes(range(256)) * 2, args)
    size = len(s) - 4
    for idx in (0, 64, 128, 448):
        t = s[:size] + idx.to_bytes(4, "little")
        rc, want = bwt_segment(t, 24)
        assert rc == OK and bwt_walk_split(bwt_list(t, size, idx), idx, size, t) == bytes(want)


def test_chunked_gap_histogram_equals_the_reference_loop(zlib_):
    """k_gap_hist counts per 4096-position chunk what LibZPAQ.cs:242-258 counts in one pass over the block (checked through the
    method string the host front end derives from either: expand_method runs the reference's loop)."""
    from native_post_model import gap_hist_chunked
    from tools import synth
    rng = np.random.default_rng(5)
    cases = [b"", b"a", b"abcabcabc" * 50, bytes(range(37)) * 400, bytes(rng.integers(0, 256, 9000, dtype=np.uint8)),
             synth.blocks("mixed", 77, 1, 20000).tobytes(), b"\\x00" * 5000 + b"\\x01" + b"\\x00" * 5000, bytes(rng.integers(0, 3, 13000, dtype=np.uint8))]
    for d in cases:
        ref = [0] * 4096
        last = [0] * 256
        for i, v in enumerate(d):
            k = i - last[v]
            if 0 < k < 4096:
                ref[k] += 1
            last[v] = i
        assert gap_hist_chunked(d) == ref, len(d)
