"""A/B of the encoders on the bench workload shape: N blocks of S bytes of the synthetic mixed stream at built-in level L.
ZPQ_PIPE=0 selects the bit-by-bit lane encoder.  Prints codec kernel time and checks block 0 against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from tools import synth
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 296
size = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
level = int(sys.argv[3]) if len(sys.argv) > 3 else 2
check = len(sys.argv) > 4 and sys.argv[4] == "check"
data = synth.blocks("mixed", 0, nblk, size).tobytes()
offs = np.arange(0, nblk * size + 1, size, dtype=np.uint64)
ctx = z.Context()
for it in range(2):
    t = time.time()
    arc, ooff = ctx.compress_blocks_level(data, offs, level)
    st = ctx.stats()
    print("pipe=%s it %d %d x %d level %d -> %d  codec_ms %.1f  MB/s %.1f  resident %d  kernel %s" % (
        os.environ.get("ZPQ_PIPE", "1"), it, nblk, size, level, arc.size, st.codec_kernel_ms, len(data) / st.codec_kernel_ms / 1e3,
        st.resident_blocks, st.kernel), flush=True)
if check:
    from oracle import pyoracle as po
    ref = po.compress_block_level(data[:size], level)
    got = arc.tobytes()[:int(ooff[1])]
    print("block0 identical to oracle:", got == ref, flush=True)
    out, _, sha, bst = ctx.decompress_blocks(arc, ooff)
    print("round trip:", out.tobytes() == data, "sha ok:", set(sha.tolist()) == {1}, flush=True)
