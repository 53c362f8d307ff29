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
