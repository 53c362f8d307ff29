"""CPU: host side of the product (no GPU): C ABI exports, front end vs oracle front end vs the
reference's known answers."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))

METHODS = ["0", "1", "2", "3", "4", "5", "6", "x0,0c256,0,255,255", "10,128,0", "11,50,0", "20,128,2", "30,128,1",
           "30,128,0", "30,10,2", "40,128,3", "41,200,3", "40,250,1", "x0,6,4,0,3,19", "x0,2,12,0,7,21,1c0,0,511i2m",
           "x0,5,4,3,3,19,1", "x0,7ci1", "s0,0c0,0,255i2", "x0,0c1,0,255,255a24mm16ts19t0w2", "x5,3ci1", "x6,7ci1",
           "x6,1,4,0,3,24", "x8,5,4,0,3,24", "x0,0c0,1003,255c0,7c0,0,1300,255c0,0,1005,0,255c1000,3c200,0,511,300",
           "s4,4c0,0,255i1,2,3,4ms20,10,100t5,20", "x1,0w2,48,10,255,16,1a30,1,2"]


def test_library_exports_every_declared_symbol(zlib_):
    L = zlib_.load()
    header = open(os.path.join(ROOT, "include", "zpaqb200.h")).read()
    declared = sorted(set(re.findall(r"\b(zpq_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(zlib_.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.zpq_version()


def test_builtin_models_are_the_reference_bytecode(zlib_):
    # Compressor.cs:48-74
    for level in (1, 2, 3):
        assert zlib_.builtin_model(level) == bytes(KAT["models"][level - 1])
    with pytest.raises(zlib_.ZpaqError):
        zlib_.builtin_model(0)
    with pytest.raises(zlib_.ZpaqError):
        zlib_.builtin_model(4)


def test_block_memory_formula(zlib_, oracle):
    # ZPAQL.memory(), ZPAQL.cs:58-81
    for level in (1, 2, 3):
        h = zlib_.builtin_model(level)
        assert zlib_.block_memory(h) == oracle.block_memory(h)
    assert abs(zlib_.block_memory(zlib_.builtin_model(2)) - 111424512) < 1000
    assert zlib_.device_state_bytes(zlib_.builtin_model(2)) < 111424512


@pytest.mark.parametrize("method", METHODS)
def test_method_expansion_config_and_bytecode_match_oracle(zlib_, oracle, method):
    # LibZPAQ.cs:128-283 (expansion), :388-1044 (makeConfig), Compiler.cs:13-478
    from oracle import frontend as fe
    from tools import synth
    data = synth.blocks("mixed", 21, 1, 70000).tobytes()
    x = zlib_.expand_method(method, data)
    assert x == fe.expand_method(method, data)
    text, args = zlib_.make_config(x)
    otext, oargs = fe.make_config(x)
    assert args == oargs
    hdr, pcomp = zlib_.compile_config(text, args)
    ohdr, opcomp, _ = fe.compile_config(otext, oargs)
    assert hdr == ohdr and pcomp == opcomp
    # each compiler accepts the other generator's text
    assert zlib_.compile_config(otext, oargs) == (ohdr, opcomp)
    assert fe.compile_config(text, args)[:2] == (hdr, pcomp)


def test_level5_period_detection(zlib_):
    from oracle import frontend as fe
    rec = bytes(range(37)) * 3000                     # period 37 records
    x = zlib_.expand_method("5", rec)
    assert "c0,0,1036,255i1" in x and x == fe.expand_method("5", rec)


def test_compiler_structured_words_and_errors(zlib_):
    from oracle import frontend as fe
    src = ("comp 2 3 0 0 1 0 cm 9 $1+3 hcomp (a (nested) comment) a=b a> 3 if a++ else a-- endif "
           "do b++ a=b a< 9 while a== 0 ifnot c=0 endif do d++ a=d a> 5 until ifl a=0 elsel a= 1 endif "
           "jmp 0 lj 0 halt post 0 end")
    args = [5, 0, 0, 0, 0, 0, 0, 0, 0]
    assert zlib_.compile_config(src, args) == fe.compile_config(src, args)[:2]
    for bad in ["comp 0 0 0 0 1 0 cm 9 hcomp halt end", "comp 0 0 0 0 0 hcomp endif halt end",
                "comp 0 0 0 0 0 hcomp a= 256 halt end", "comp 0 0 0 0 0 hcomp bogus end", "comp 0 0 0 0 0 hcomp halt"]:
        with pytest.raises(zlib_.ZpaqError):
            zlib_.compile_config(bad, None)
        with pytest.raises(fe.ConfigError):
            fe.compile_config(bad, None)


def test_find_blocks_and_size_bound_on_host(zlib_, oracle):
    # Decompresser.findBlock, Decompresser.cs:29-58
    from tools import synth
    data = synth.blocks("text", 60, 1, 9000).tobytes()
    a = oracle.compress_block(data[:4000], "x0,0c0,0,255")
    b = oracle.compress_block_level(data[4000:], 1, with_tag=False)
    arc = b"garbage" + a + b        # a tagless block is only found directly after the previous block
    blocks = zlib_.find_blocks(arc)
    assert len(blocks) == 2
    assert arc[blocks[0][0]:blocks[0][0] + 3] == b"zPQ" and arc[blocks[1][0]:blocks[1][0] + 3] == b"zPQ"
    assert blocks[0][1] == 7 + len(a) == blocks[1][0] and blocks[1][1] == len(arc)
    offs = np.asarray([blocks[0][0], blocks[0][1]], dtype=np.uint64)
    buf = np.frombuffer(arc, dtype=np.uint8)
    assert zlib_.load().zpq_decompressed_bound(buf.ctypes.data, offs.ctypes.data, 1) >= 4000


def test_context_creation_fails_loudly_without_gpu(zlib_):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zlib_.ZpaqError) as e:
        zlib_.Context()
    assert "CUDA" in str(e.value)


def test_role_split_encoder_plan_for_builtin_models(zlib_):
    # host logic of the role-split encoder (zpq_duo.cuh) without a device: which models it takes, lanes per block,
    # role warps per group, and that mid.cfg fits its 11 resident blocks per SM into the 227 KB an SM has
    B200_SMEM = 232448
    mid = zlib_.encoder_plan(zlib_.builtin_model(2), B200_SMEM, 11)
    assert mid["applies"] == 1 and mid["lanes_per_block"] == 8 and mid["mixer_role"] == 1 and mid["roles"] == 4
    assert mid["blocks_per_sm"] == 11 and mid["warps_per_cta"] == 1 + 4 * 3 and mid["coder_delay"] == 7
    assert 11 * mid["smem_per_block"] + 81 * 1024 <= B200_SMEM
    mn = zlib_.encoder_plan(zlib_.builtin_model(1), B200_SMEM, 32)
    assert mn["applies"] == 1 and mn["lanes_per_block"] == 8 and mn["mixer_role"] == 0 and mn["roles"] == 3
    assert mn["blocks_per_sm"] == 20          # 15 role warps / 3 roles = 5 groups of 4 blocks
    mx = zlib_.encoder_plan(zlib_.builtin_model(3), B200_SMEM, 5)
    assert mx["applies"] == 1 and mx["lanes_per_block"] == 32 and mx["mixer_role"] == 0   # MIX2/SSE read the MIX outputs
    assert mx["blocks_per_sm"] == 5 and mx["warps_per_cta"] == 16
    # a model with more than 32 components is left to the step-scheduled kernels
    hdr, _ = zlib_.compile_config(zlib_.make_config("x0,0" + "c0,0,255" * 40)[0], [0] * 9) if hasattr(zlib_, "compile_config") else (None, None)
    if hdr:
        assert zlib_.encoder_plan(hdr)["applies"] == 0


def test_find_blocks_locates_segment_ends_at_every_alignment(zlib_, oracle):
    """The host framing parse probes every fourth byte for the four zero bytes that end coded data (Decoder.skip, Decoder.cs:70-98);
    archives shifted by 0..7 junk bytes, with short zero runs inside the coded data, must give the same block boundaries."""
    from tools import synth
    arcs = []
    for i in range(12):
        d = synth.blocks("mixed", 2000 + i, 1, 3000 + 37 * i).tobytes() + b"\x00" * (i % 5) + bytes(range(i))
        arcs.append(oracle.compress_block_level(d, 1) if i % 3 else oracle.compress_block(d, "x0,0c0,0,255"))
    for shift in range(8):
        junk = bytes([7] * shift)
        blob = junk + b"".join(arcs)
        got = zlib_.find_blocks(blob)
        want, pos = [], shift
        for a in arcs:
            want.append((pos + 13, pos + len(a)))            # the block starts at "zPQ", behind the 13-byte tag
            pos += len(a)
        assert got == want
    # stored (n = 0) blocks: lengths, not zero runs, delimit the data
    d = b"\x00" * 5000 + synth.blocks("text", 1, 1, 2000).tobytes()
    a = oracle.compress_block(d, "0")
    assert zlib_.find_blocks(a + a) == [(13, len(a)), (len(a) + 13, 2 * len(a))]
