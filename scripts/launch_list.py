"""Every kernel of the path once, for an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv ...`): compress + decompress
of a small batch of every BASELINE config, a level-5 method (byte-gap histograms) and a foreign PCOMP program (interpreter pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from tools import synth

CASES = [("C2a mid.cfg", "mixed", synth.BLOCK_1MB, 296, ("level", 2)),
         ("C1 order-2 CM", "text", synth.BLOCK_1MB, 296, ("method", "x0,0c256,0,255,255")),
         ("C2b LZ77-SA stored", "mixed", synth.BLOCK_1MB, 148, ("method", "x0,1,4,0,7,21,1")),
         ("C2c LZ77-SA + CM", "mixed", synth.BLOCK_1MB, 148, ("method", "x0,2,12,0,7,21,1c0,0,511i2m")),
         ("C3 BWT", "text", synth.BLOCK_4MB, 74, ("method", "x2,3ci1")),
         ("C4 max.cfg", "mixed", 200000, 296, ("level", 3)),
         ("method 1 = LZ77 hash", "mixed", synth.BLOCK_1MB, 148, ("method", "1")),
         ("E8E9 + CM", "mixed", synth.BLOCK_1MB, 148, ("method", "x0,4c0,0,255")),
         ("level 5 (gap analysis)", "mixed", 100000, 74, ("method", "5"))]
ctx = z.Context()
for name, kind, bs, nb, (how, arg) in CASES:
    data = synth.blocks(kind, 0, nb, bs)
    offs = np.arange(0, (nb + 1) * bs, bs, dtype=np.uint64)
    arc, ooff = ctx.compress_blocks_level(data, offs, arg) if how == "level" else ctx.compress_blocks(data, offs, arg)
    st = ctx.stats()
    out, _, sha, _ = ctx.decompress_blocks(arc, ooff)
    sd = ctx.stats()
    print("%-24s %4d x %8d  compress kernels %.1f ms (codec %.1f)  decompress kernels %.1f ms (codec %.1f post %.1f)  round trip %s" % (
        name, nb, bs, st.kernel_ms, st.codec_kernel_ms, sd.kernel_ms, sd.codec_kernel_ms, sd.post_kernel_ms,
        bool(np.array_equal(out, data)) and set(sha.tolist()) == {1}), flush=True)
