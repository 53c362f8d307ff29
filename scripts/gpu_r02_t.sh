# bench.py main path with the C5 sweep in the summary (small headline batch)
cd /root/repo
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch-blocks 148 --configs C5 > gpurun_out/r02t_bench_c5.json 2> gpurun_out/r02t_bench_c5.err; tail -2 gpurun_out/r02t_bench_c5.err
python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_c5.json')); print(d['value'], d['decompress']['e2e_value'], [ (r['block_bytes'], round(r['decompress_e2e_value'],1), r['round_trip_identical']) for r in d['configs']['C5']['sweep']], sorted(d['configs']))"
