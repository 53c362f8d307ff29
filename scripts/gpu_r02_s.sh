# last check of bench.py's code paths after the edits: a small headline-only line, then the C5 sweep through finish_c5
cd /root/repo
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --batch-blocks 148 > gpurun_out/r02s_bench_small.json 2> gpurun_out/r02s_bench_small.err; tail -2 gpurun_out/r02s_bench_small.err; head -c 300 gpurun_out/r02s_bench_small.json; echo
timeout 300 python scripts/c5_quick.py > gpurun_out/r02s_c5.log 2>&1; tail -3 gpurun_out/r02s_c5.log
