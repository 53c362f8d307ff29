# multi-segment decode + per-segment SHA-1, then the launch list of every kernel of the path
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_segment or two_segment or golden or reference" > gpurun_out/r02k_tests.log 2>&1; tail -4 gpurun_out/r02k_tests.log
timeout 900 python -m pytest tests/test_gpu_reference_text.py tests/test_gpu_postproc.py -x -q > gpurun_out/r02k_tests2.log 2>&1; tail -3 gpurun_out/r02k_tests2.log
timeout 300 python scripts/launch_list.py > gpurun_out/r02k_launch_plain.log 2>&1; cat gpurun_out/r02k_launch_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv python scripts/launch_list.py > gpurun_out/r02k_launch_ncu.log 2>&1; tail -3 gpurun_out/r02k_launch_ncu.log; wc -l gpurun_out/r02_launches.csv
