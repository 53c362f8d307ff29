"""CPU: the oracle against every known-answer item the reference holds for the path."""
import base64
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
ARCHIVES = json.load(open(os.path.join(HERE, "golden", "oracle_archives.json")))


def _tables(oracle):
    sq = np.zeros(4096, np.uint16); st = np.zeros(32768, np.int16)
    dt = np.zeros(1024, np.int32); dt2k = np.zeros(256, np.int32); ns = np.zeros(1024, np.uint8)
    oracle.lib().orc_tables(sq.ctypes.data, st.ctypes.data, dt.ctypes.data, dt2k.ctypes.data, ns.ctypes.data)
    return sq, st, dt, dt2k, ns


def test_state_table_matches_reference_literal(oracle):
    # StateTable.cs:21-149
    ns = _tables(oracle)[4]
    assert ns.tolist() == KAT["sns"]


def test_division_tables_match_reference_literals(oracle):
    # Predictor.cs:1359-1524
    _, _, dt, dt2k, _ = _tables(oracle)
    assert dt.tolist() == KAT["sdt"]
    assert dt2k.tolist() == KAT["sdt2k"]


def test_squash_stretch_match_reference_literals_and_checksums(oracle):
    # Predictor.cs:54-78, 1527-1790
    sq, st, _, _, _ = _tables(oracle)
    assert sq[1376:1376 + 1344].tolist() == KAT["ssquasht"]
    assert not sq[:1376].any() and (sq[2720:] == 32767).all()
    run = []
    k = 16384
    for i, cnt in enumerate(KAT["stdt"]):
        run += [i] * cnt
    assert st[16384:].tolist() == run
    assert st[:16384].tolist() == [-v for v in reversed(run)]
    stsum = sqsum = 0
    for i in range(32767, -1, -1):
        stsum = (stsum * 3 + int(st[i])) & 0xFFFFFFFF
    for i in range(4095, -1, -1):
        sqsum = (sqsum * 3 + int(sq[i])) & 0xFFFFFFFF
    assert stsum == KAT["stsum"] and sqsum == KAT["sqsum"]


def test_builtin_models_compile_to_reference_bytecode(oracle):
    # Compressor.cs:48-74 (bytes) and LICENSE:391-400 (min.cfg source): the Compiler KAT
    from oracle import frontend as fe
    for level in (1, 2, 3):
        hdr, pcomp = fe.builtin_model(level)
        assert hdr == bytes(KAT["models"][level - 1])
        assert pcomp == b""


def test_compsize(oracle):
    from oracle import frontend as fe
    assert fe.COMPSIZE == KAT["compsize"][:10]


def test_locator_tag_and_findblock_hash_constants():
    # Compressor.cs:27-43 <-> Decompresser.cs:34,43
    mult = [12, 20, 28, 44]
    h = [0, 0, 0, 0]
    for c in KAT["tag"]:
        h = [(h[i] * mult[i] + c) & 0xFFFFFFFF for i in range(4)]
    # the multipliers are multiples of 4, so only the low 26 bits of a seed survive the 3 bytes "zPQ"
    assert [x & 0x3FFFFFF for x in h] == [x & 0x3FFFFFF for x in KAT["findblock_seed"]]
    h = list(KAT["findblock_seed"])
    for c in b"zPQ":
        h = [(h[i] * mult[i] + c) & 0xFFFFFFFF for i in range(4)]
    assert h == KAT["findblock_hit"]


def test_stored_mode_framing_is_hand_checkable(oracle):
    # Encoder.cs:62-71, Compressor.cs:235-238: "abc" with n = 0 -> 00 00 00 04 00 61 62 63 (the leading 00 is
    # the PASS byte of Compressor.postProcess) then 00 00 00 00
    arc = oracle.compress_block(b"abc", "0", dosha1=False)
    assert arc[:13] == bytes(KAT["tag"])
    assert arc[13:18] == b"zPQ\x02\x01"
    body = arc[arc.index(b"\x003\x00\x00") + 4:]     # after comment "3", reserved 0
    assert body == bytes([0, 0, 0, 4, 0, 0x61, 0x62, 0x63, 0, 0, 0, 0, 254, 255])


def test_sha1_is_fips180(oracle):
    for n in (0, 1, 55, 56, 63, 64, 65, 1000):
        data = bytes(range(256)) * 4
        assert oracle.sha1(data[:n]) == hashlib.sha1(data[:n]).digest()


@pytest.mark.parametrize("case", ARCHIVES, ids=[c["name"] for c in ARCHIVES])
def test_oracle_reproduces_golden_archives(oracle, case):
    from tools import synth
    n = case["nbytes"]
    data = synth.blocks(case["kind"], case["first_block"], 1, max(n, 1)).tobytes()[:n] if n else b""
    assert hashlib.sha1(data).hexdigest() == case["input_sha1"]
    if case["how"] == "level":
        arc = oracle.compress_block_level(data, case["arg"])
    else:
        arc = oracle.compress_block(data, case["arg"])
    assert arc == base64.b64decode(case["archive_b64"])
    back, status = oracle.decompress(arc)
    assert back == data and status == [1]


@pytest.mark.parametrize("method", ["1", "2", "3", "4", "x0,5,4,0,3,19", "x0,6,8,0,5,18c0,0,511", "x0,7ci1", "x0,4",
                                    "x0,1,4,2,3,16,1", "x0,2,3,5,2,17,2c0,0,511i1"])
def test_oracle_preprocessors_invert(oracle, method):
    # the DEBUG invariant of LibZPAQ.cs:314-320: PCOMP(pre-processed stream) == input
    from tools import synth
    data = synth.blocks("mixed", 40, 1, 90000).tobytes()
    arc = oracle.compress_block(data, method)
    back, status = oracle.decompress(arc)
    assert back == data and status == [1]


def test_oracle_multi_block_compress_matches_block_api(oracle):
    from tools import synth
    data = synth.blocks("text", 50, 1, 30000).tobytes()
    a = oracle.compress_block(data[:10000], "x0,0c0,0,255i1", "f.txt", "hello")
    assert b"f.txt\x0010000 hello\x00" in a[:200]
    back, status = oracle.decompress(a + oracle.compress_block(data[10000:], "1"))
    assert back == data and status == [1, 1]


def test_suffix_array_small(oracle):
    rng = np.random.default_rng(7)
    for n in (1, 2, 17, 500):
        s = bytes(rng.integers(97, 100, size=n, dtype=np.uint8))
        sa = np.zeros(n, np.int32)
        oracle.lib().orc_suffix_array(s, n, sa.ctypes.data)
        assert sa.tolist() == sorted(range(n), key=lambda i: s[i:])


def test_e8e9_reference_semantics(oracle):
    # LibZPAQ.cs:372-384, including the overlapping-pattern chain the descending scan resolves
    buf = bytearray(b"\x00" * 40)
    buf[10:15] = b"\xe8\x01\x02\x03\x00"
    buf[20:25] = b"\xe9\xff\xff\xff\xff"
    buf[30:35] = b"\xe8\xe8\x00\x00\x00"
    expect = bytearray(buf)
    for i in range(len(expect) - 5, -1, -1):
        if (expect[i] & 254) == 0xe8 and ((expect[i + 4] + 1) & 254) == 0:
            a = (expect[i + 1] | expect[i + 2] << 8 | expect[i + 3] << 16) + i
            expect[i + 1], expect[i + 2], expect[i + 3] = a & 255, (a >> 8) & 255, (a >> 16) & 255
    arr = (C.c_ubyte * len(buf)).from_buffer(buf)
    oracle.lib().orc_e8e9(arr, len(buf))
    assert bytes(buf) == bytes(expect)
