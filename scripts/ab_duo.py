"""Parity + timing of the encoders on N blocks of S bytes of the synthetic mixed stream.
usage: ab_duo.py NBLK SIZE LEVEL_OR_METHOD [check]   (ZPQ_DUO=0 / ZPQ_PIPE=0 select the older encoders)
`check` compares every block with the CPU oracle and round-trips through the GPU decoder."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from tools import synth
nblk = int(sys.argv[1]); size = int(sys.argv[2]); what = sys.argv[3]
check = len(sys.argv) > 4 and sys.argv[4] == "check"
ragged = len(sys.argv) > 5 and sys.argv[5] == "ragged"
data = synth.blocks("mixed", 0, nblk, size).tobytes()
if ragged:
    rng = np.random.default_rng(7)
    lens = rng.integers(0, size + 1, nblk); lens[0] = 0; lens[-1] = 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = data[:int(offs[-1])]
else:
    offs = np.arange(0, nblk * size + 1, size, dtype=np.uint64)
ctx = z.Context()
level = int(what) if what in ("1", "2", "3") and len(what) == 1 else None
for it in range(2):
    t = time.time()
    if level: arc, ooff = ctx.compress_blocks_level(data, offs, level)
    else: arc, ooff = ctx.compress_blocks(data, offs, what)
    st = ctx.stats()
    print("duo=%s pipe=%s it %d %d x %d %s -> %d  codec_ms %.1f  MB/s %.1f  resident %d  kernel %s" % (
        os.environ.get("ZPQ_DUO", "1"), os.environ.get("ZPQ_PIPE", "1"), it, nblk, size, what, arc.size, st.codec_kernel_ms,
        len(data) / max(st.codec_kernel_ms, 1e-9) / 1e3, st.resident_blocks, st.kernel), flush=True)
if check:
    from oracle import pyoracle as po
    bad = 0
    a = arc.tobytes()
    for i in range(nblk):
        blk = data[int(offs[i]):int(offs[i + 1])]
        ref = po.compress_block_level(blk, level) if level else po.compress_block(blk, what)
        got = a[int(ooff[i]):int(ooff[i + 1])]
        if got != ref:
            bad += 1
            k = next((j for j in range(min(len(got), len(ref))) if got[j] != ref[j]), min(len(got), len(ref)))
            print("  block %d differs: len %d vs %d, first at %d" % (i, len(got), len(ref), k), flush=True)
            if bad > 4: break
    print("identical to oracle: %s (%d blocks)" % (bad == 0, nblk), flush=True)
    out, _, sha, bst = ctx.decompress_blocks(arc, ooff)
    print("round trip:", out.tobytes() == data, "sha ok:", set(sha.tolist()) <= {1}, flush=True)
