# facade batching + large-block parity tests, ncu of the speculative decoder at bench residency, the new default bench
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_facade.py -x -q > gpurun_out/r02e_facade.log 2>&1; tail -5 gpurun_out/r02e_facade.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -k "large_blocks or max_cfg or every_block or 16mb or nvrtc" --durations=10 > gpurun_out/r02e_parity.log 2>&1; tail -25 gpurun_out/r02e_parity.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_aot2_f -c 1 -o gpurun_out/r02e_fdec_1607x200k -f python scripts/ab_dec.py 1607 200000 2 mixed 1 > gpurun_out/r02e_ncu.log 2>&1; tail -3 gpurun_out/r02e_ncu.log
( time python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ) 2> gpurun_out/r02e_bench.time; cat gpurun_out/r02e_bench.time; tail -5 gpurun_out/r02e_bench.err; head -c 1500 gpurun_out/r02e_bench.json
