"""Round trip + timing of the decoders on N blocks of S bytes of the synthetic stream.
usage: ab_dec.py NBLK SIZE LEVEL_OR_METHOD [kind] [reps]      (ZPQ_FDEC=0 selects the lane-resident decoder with the
post-processor inside; default = the speculative decoder of zpq_fdec.cuh where the model allows it)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from tools import synth
nblk = int(sys.argv[1]); size = int(sys.argv[2]); what = sys.argv[3]
kind = sys.argv[4] if len(sys.argv) > 4 else "mixed"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
data = synth.blocks(kind, 0, nblk, size).tobytes()
offs = np.arange(0, nblk * size + 1, size, dtype=np.uint64)
ctx = z.Context()
level = int(what) if what in ("1", "2", "3") else None
t = time.time()
arc, ooff = ctx.compress_blocks_level(data, offs, level) if level else ctx.compress_blocks(data, offs, what)
st = ctx.stats()
print("compress %d x %d %s -> %d  codec_ms %.1f  kernel_ms %.1f  wall %.2fs  kernel %s" % (
    nblk, size, what, arc.size, st.codec_kernel_ms, st.kernel_ms, time.time() - t, st.kernel.decode()), flush=True)
for it in range(reps):
    t = time.time()
    try:
        out, _, sha, bst = ctx.decompress_blocks(arc, ooff)
        ok = out.tobytes() == data
    except Exception as e:
        print("decompress failed:", e, flush=True)
        break
    st = ctx.stats()
    print("fdec=%s it %d decode codec_ms %.1f post_ms %.1f kernel_ms %.1f wall %.2fs  MB/s(codec) %.1f MB/s(wall) %.1f  round trip %s sha %s  kernel %s" % (
        os.environ.get("ZPQ_FDEC", "1"), it, st.codec_kernel_ms, st.post_kernel_ms, st.kernel_ms, time.time() - t,
        len(data) / max(st.codec_kernel_ms, 1e-9) / 1e3, len(data) / 1e6 / (time.time() - t), ok, sorted(set(sha.tolist())),
        st.kernel.decode()), flush=True)
    if not ok:
        o = out.tobytes()
        k = next((j for j in range(min(len(o), len(data))) if o[j] != data[j]), min(len(o), len(data)))
        print("  first difference at byte %d (block %d, offset %d); lengths %d vs %d" % (k, k // size, k % size, len(o), len(data)), flush=True)
        break
