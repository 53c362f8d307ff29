set -x
cd /root/repo
python scripts/ab_dec.py 8 3000 2 > gpurun_out/r02a_small.log 2>&1
python scripts/ab_dec.py 64 60000 2 >> gpurun_out/r02a_small.log 2>&1
python scripts/ab_dec.py 64 60000 1 >> gpurun_out/r02a_small.log 2>&1
python scripts/ab_dec.py 16 60000 "x0,0c256,0,255,255" text >> gpurun_out/r02a_small.log 2>&1
python scripts/ab_dec.py 16 60000 "x0,2,12,0,7,21,1c0,0,511i2m" >> gpurun_out/r02a_small.log 2>&1
python scripts/ab_dec.py 16 60000 "32,128,1" text >> gpurun_out/r02a_small.log 2>&1
tail -40 gpurun_out/r02a_small.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; tail -5 gpurun_out/r02a_pytest.log
python scripts/ab_dec.py 1607 1044480 2 mixed 2 > gpurun_out/r02a_full.log 2>&1; cat gpurun_out/r02a_full.log
ZPQ_FDEC=0 python scripts/ab_dec.py 592 200000 2 mixed 1 > gpurun_out/r02a_old.log 2>&1; cat gpurun_out/r02a_old.log
python scripts/ab_dec.py 592 200000 2 mixed 1 > gpurun_out/r02a_new592.log 2>&1; cat gpurun_out/r02a_new592.log
