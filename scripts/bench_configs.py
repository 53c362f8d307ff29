"""Throughput of the other BASELINE configs (C1, C2b, C2c, C3, C4) next to the headline
C2a: compress and decompress e2e through the C ABI, plus the CPU oracle on one core for scale."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zpaqsharp_b200 import libzpaq as z
from oracle import pyoracle as po
from tools import synth

CASES = [
    ("C1 order-2 CM x0,0c256,0,255,255 (1 MB blocks, text)", "text", synth.BLOCK_1MB, 1024, ("method", "x0,0c256,0,255,255")),
    ("C2a mid.cfg startBlock(2) (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 1024, ("level", 2)),
    ("C2b method 20 = LZ77-SA, stored (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 256, ("method", "20")),
    ("C2c x0,2,12,0,7,21,1c0,0,511i2m (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 256, ("method", "x0,2,12,0,7,21,1c0,0,511i2m")),
    ("C3 method 32,128,1 = BWT + icm/isse (4 MB, text)", "text", synth.BLOCK_4MB, 64, ("method", "32,128,1")),
    ("C4 max.cfg startBlock(3) (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 700, ("level", 3)),
    ("method 10 = LZ77 hash, stored (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 512, ("method", "10")),
    ("min.cfg startBlock(1) (1 MB, mixed)", "mixed", synth.BLOCK_1MB, 2048, ("level", 1)),
]
only = sys.argv[1:]
ctx = z.Context()
rows = []
for name, kind, bs, nb, (how, arg) in CASES:
    if only and not any(o in name for o in only):
        continue
    data = synth.blocks(kind, 0, nb, bs)
    offs = np.arange(0, (nb + 1) * bs, bs, dtype=np.uint64)
    run = (lambda: ctx.compress_blocks_level(data, offs, arg)) if how == "level" else (lambda: ctx.compress_blocks(data, offs, arg))
    arc, ooff = run()                       # warm-up (NVRTC, allocations)
    t = time.perf_counter(); arc, ooff = run(); tc = time.perf_counter() - t
    st = ctx.stats()
    kernel = st.kernel.decode(errors="replace"); resident = st.resident_blocks
    out, _, sha, _ = ctx.decompress_blocks(arc, ooff)
    t = time.perf_counter(); out, _, sha, _ = ctx.decompress_blocks(arc, ooff); td = time.perf_counter() - t
    ok = bool(np.array_equal(out, data)) and set(sha.tolist()) == {1}
    blk = data[:bs].tobytes()
    t = time.perf_counter()
    ref = po.compress_block_level(blk, arg) if how == "level" else po.compress_block(blk, arg)
    tcpu = time.perf_counter() - t
    same = arc[:int(ooff[1])].tobytes() == ref
    row = {"config": name, "blocks": nb, "block_bytes": bs, "compress_MBps": nb * bs / 1e6 / tc, "decompress_MBps": nb * bs / 1e6 / td,
           "ratio": arc.size / data.size, "round_trip": ok, "block0_identical_to_oracle": same, "kernel": kernel,
           "resident_blocks": int(resident), "cpu_oracle_1core_MBps": bs / 1e6 / tcpu}
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs.json"), "w"), indent=1)
