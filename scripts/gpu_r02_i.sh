# A/B: L2 prefetch of MIX rows in the speculative decoder (1776 x 200 KB, and full-size blocks once)
cd /root/repo
timeout 300 python scripts/ab_dec.py 1776 200000 2 mixed 2 > gpurun_out/r02i_1776.log 2>&1; cat gpurun_out/r02i_1776.log
timeout 300 python scripts/ab_dec.py 592 100000 "x0,2,12,0,7,21,1c0,0,511i2m" mixed 1 >> gpurun_out/r02i_1776.log 2>&1; tail -1 gpurun_out/r02i_1776.log
