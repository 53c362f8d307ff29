# round 2, first call of this session: state check (GPU tests, full-size decode A/B line, default bench)
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; tail -5 gpurun_out/r02c_pytest.log
python scripts/ab_dec.py 1607 1044480 2 mixed 2 > gpurun_out/r02c_full.log 2>&1; cat gpurun_out/r02c_full.log
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json
