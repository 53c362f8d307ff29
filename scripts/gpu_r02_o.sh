# SA matcher restructured (all pairs in one warp round): parity tests + timing; concurrent decode groups with the allocation barrier: C5 legs
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lz77_bwt or preprocessing_full or large_blocks or mixed_method or golden" > gpurun_out/r02o_tests.log 2>&1; tail -4 gpurun_out/r02o_tests.log
for m in "x0,1,4,0,7,21,1" "x0,2,12,0,7,21,1c0,0,511i2m"; do timeout 300 python scripts/ab_dec.py 592 1044480 "$m" mixed 1 2>&1 | grep "^compress"; done > gpurun_out/r02o_lz.log 2>&1; cat gpurun_out/r02o_lz.log | cut -c1-140
timeout 600 python scripts/c5_quick.py > gpurun_out/r02o_c5.log 2>&1; tail -4 gpurun_out/r02o_c5.log
