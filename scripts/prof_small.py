"""Small single-model encode for ncu: N blocks of S bytes of synthetic text at built-in level L."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
level = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(1)
words = [bytes(rng.integers(97, 123, size=rng.integers(2, 10)).astype(np.uint8)) for _ in range(500)]
text = b' '.join(words[i] for i in rng.zipf(1.3, size=nblk * size // 4) % 500)[:nblk * size]
offs = np.arange(0, nblk * size + 1, size, dtype=np.uint64)
ctx = z.Context()
for it in range(2):
    t = time.time()
    arc, ooff = ctx.compress_blocks_level(text, offs, level)
    st = ctx.stats()
    print("it", it, "bytes", len(text), "->", arc.size, "wall %.3f codec_ms %.1f resident %d" % (time.time() - t, st.codec_kernel_ms, st.resident_blocks),
          "us/byte(per block) %.2f" % (st.codec_kernel_ms * 1e3 / size), flush=True)
