# whole GPU suite on the current library, then the default bench (the line the driver will take)
cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02l_pytest.log 2>&1 ) 2> gpurun_out/r02l_pytest.time; tail -14 gpurun_out/r02l_pytest.log; cat gpurun_out/r02l_pytest.time
( time timeout 1500 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err ) 2> gpurun_out/r02l_bench.time; cat gpurun_out/r02l_bench.time; tail -3 gpurun_out/r02l_bench.err; head -c 400 gpurun_out/r02l_bench.json
