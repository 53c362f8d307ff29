# is the speculative decoder latency- or issue-bound?  same block size, 4 / 8 / 11 blocks per SM; + SHA-1 kernel check
cd /root/repo
for n in 592 1184 1628; do timeout 300 python scripts/ab_dec.py $n 200000 2 mixed 1 2>&1 | grep -v "^$"; done > gpurun_out/r02g_scaling.log 2>&1
cat gpurun_out/r02g_scaling.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or builtin_levels" > gpurun_out/r02g_tests.log 2>&1; tail -3 gpurun_out/r02g_tests.log
