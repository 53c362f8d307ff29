# native post-processors: new tests, the whole GPU suite, decode timings per config
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_postproc.py -x -q > gpurun_out/r02d_post.log 2>&1; tail -15 gpurun_out/r02d_post.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; tail -8 gpurun_out/r02d_pytest.log
for m in "x0,1,4,0,7,21,1" "x0,2,12,0,7,21,1c0,0,511i2m" "x0,0c256,0,255,255"; do timeout 600 python scripts/ab_dec.py 592 1044480 "$m" mixed 1; done > gpurun_out/r02d_cfg.log 2>&1
timeout 600 python scripts/ab_dec.py 148 4190208 "x2,3ci1" text 1 >> gpurun_out/r02d_cfg.log 2>&1
cat gpurun_out/r02d_cfg.log
