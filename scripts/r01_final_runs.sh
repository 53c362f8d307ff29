cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generations or many_blocks" > gpurun_out/pytest_new.log 2>&1; tail -3 gpurun_out/pytest_new.log
timeout 900 python scripts/bench_configs.py > gpurun_out/configs_run.log 2>&1; tail -2 gpurun_out/configs_run.log | cut -c1-200
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --batch-blocks 592 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1; echo launchlist rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zpq_enc -c 1 -o gpurun_out/prof_final_enc -f python bench.py --steps 1 --warmup 0 --batch-blocks 592 --no-cpu-baseline --no-decompress > gpurun_out/ncu_full_final.log 2>&1; echo ncufull rc=$?
