cd /root/repo
for m in 2 1 "x0,0c256,0,255,255" "x0,2,12,0,7,21,1c0,0,511i2m" "32,128,1"; do python scripts/ab_dec.py 16 60000 "$m" text 1; done > gpurun_out/r02b_small.log 2>&1
grep -c "round trip True" gpurun_out/r02b_small.log; grep "round trip False\|failed" gpurun_out/r02b_small.log
python scripts/ab_dec.py 1607 1044480 2 mixed 2 > gpurun_out/r02b_full.log 2>&1; cat gpurun_out/r02b_full.log
python scripts/ab_dec.py 592 200000 2 mixed 1 > gpurun_out/r02b_592.log 2>&1; cat gpurun_out/r02b_592.log
