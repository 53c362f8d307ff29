"""First GPU bring-up: compress/decompress small inputs through the C ABI and compare with the oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from oracle import pyoracle as po, frontend as fe

rng = np.random.default_rng(1)
words = [bytes(rng.integers(97, 123, size=rng.integers(2, 10)).astype(np.uint8)) for _ in range(500)]
text = b' '.join(words[i] for i in rng.zipf(1.3, size=12000) % 500)
rnd = bytes(rng.integers(0, 256, size=30000, dtype=np.uint8))
data = text + rnd + text[:20000]
print("input", len(data), flush=True)
ctx = z.Context()
bad = 0
offs = np.asarray([0, 40000, 40000, 90000, len(data)], dtype=np.uint64)
for lv in (1, 2, 3):
    t = time.time()
    arc, ooff = ctx.compress_blocks_level(data, offs, lv)
    dt = time.time() - t
    ref = b''.join(po.compress_block_level(data[int(offs[i]):int(offs[i+1])], lv) for i in range(len(offs) - 1))
    ok = arc.tobytes() == ref
    st = ctx.stats()
    print("level", lv, "gpu bytes", arc.size, "oracle", len(ref), "MATCH" if ok else "DIFF", "%.2fs" % dt,
          "codec_ms %.1f resident %d state %d" % (st.codec_kernel_ms, st.resident_blocks, st.state_bytes_per_block), flush=True)
    if not ok:
        bad += 1
        a = arc.tobytes()
        for i in range(min(len(a), len(ref))):
            if a[i] != ref[i]:
                print("  first diff at", i, a[i-4:i+8].hex(), ref[i-4:i+8].hex()); break
    # decode on GPU
    t = time.time()
    out, o2, sha, bst = ctx.decompress_blocks(ref, ooff if ok else np.asarray([0, len(ref)], dtype=np.uint64)) if ok else (None,)*4
    if ok:
        good = out.tobytes() == data
        print("   decode", "OK" if good else "BAD", list(sha), list(bst), "%.2fs" % (time.time() - t), flush=True)
        bad += 0 if good else 1

for m in ['0', 'x0,0c256,0,255,255', 's0,0c0,0,255i2', 'x0,4ci1,1,1,1,2awm', 'x0,0c1,0,255,255a24mm16ts19t0w2', '4', '5', '60,200,3']:
    blk = data[:60000]
    try:
        arc, ooff = ctx.compress_blocks(blk, np.asarray([0, 25000, len(blk)], dtype=np.uint64), m)
        ref = po.compress_block(blk[:25000], m) + po.compress_block(blk[25000:], m)
        ok = arc.tobytes() == ref
        print("method", m, arc.size, len(ref), "MATCH" if ok else "DIFF", flush=True)
        bad += 0 if ok else 1
        out, o2, sha, bst = ctx.decompress_blocks(ref, ooff)
        good = out.tobytes() == blk
        print("   decode", "OK" if good else "BAD", list(sha), list(bst), flush=True)
        bad += 0 if good else 1
    except Exception as e:
        print("method", m, "ERROR", e, flush=True)
        bad += 1

# decode-only: archives with LZ77/BWT post-processing produced by the oracle
for m in ['1', '2', '3', '30,128,1', 'x0,5,4,0,3,19', 'x0,6,12,0,3,19c0,0,511', 'x0,7ci1', 'x0,4']:
    blk = data[:50000]
    try:
        ref = po.compress_block(blk, m)
        out, o2, sha, bst = ctx.decompress_blocks(ref, np.asarray([0, len(ref)], dtype=np.uint64))
        good = out.tobytes() == blk
        print("decode-only", m, fe.expand_method(m, blk), "OK" if good else "BAD", list(sha), list(bst), flush=True)
        bad += 0 if good else 1
    except Exception as e:
        print("decode-only", m, "ERROR", e, flush=True)
        bad += 1
print("FAILURES", bad)
sys.exit(1 if bad else 0)
