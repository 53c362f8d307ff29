# decoder without the 8-candidate L2 prefetch: timing at 12 blocks per SM, round trip on three models, ncu --set full for the traffic
cd /root/repo
timeout 200 python scripts/ab_dec.py 1776 200000 2 mixed 2 2>&1 | grep "^fdec=" > gpurun_out/r02p_nopf8.log
timeout 200 python scripts/ab_dec.py 148 60000 1 mixed 1 2>&1 | grep "^fdec=" >> gpurun_out/r02p_nopf8.log
timeout 200 python scripts/ab_dec.py 148 60000 "x0,2,12,0,7,21,1c0,0,511i2m" mixed 1 2>&1 | grep "^fdec=" >> gpurun_out/r02p_nopf8.log
cut -c1-170 gpurun_out/r02p_nopf8.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_aot2_f -c 1 -o gpurun_out/r02p_fdec_296x50k -f python scripts/ab_dec.py 296 50000 2 mixed 1 > gpurun_out/r02p_ncu.log 2>&1; tail -2 gpurun_out/r02p_ncu.log
