"""The C5 sweep of bench.py alone (mixed-method archives, default legs; `full` adds the 16,773,120-byte leg)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rig = bench.Rig()
full = len(sys.argv) > 1 and sys.argv[1] == "full"
rows = bench.finish_c5(bench.c5_sweep(rig, full=full), None, rig.world, 4 if full else 3, full, lambda t: t)["sweep"]
for r in rows:
    print(json.dumps({k: r[k] for k in ("block_bytes", "blocks_per_gpu", "decompress_e2e_value", "round_trip_identical")}), flush=True)
