# level-5 analysis test, ncu (application replay, bench residency) of decoder + encoder, default bench with fresh-context configs
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "level5 or every_block" > gpurun_out/r02f_tests.log 2>&1; tail -5 gpurun_out/r02f_tests.log
timeout 1200 ncu --replay-mode application --clock-control none --import-source on --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  -k regex:"zpq_dec_aot2_f|zpq_enc_aot2_d" -c 2 -o gpurun_out/r02f_codec_1700x200k -f python scripts/ab_dec.py 1700 200000 2 mixed 1 > gpurun_out/r02f_ncu.log 2>&1; tail -4 gpurun_out/r02f_ncu.log
( time timeout 1500 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err ) 2> gpurun_out/r02f_bench.time; cat gpurun_out/r02f_bench.time; tail -5 gpurun_out/r02f_bench.err; head -c 600 gpurun_out/r02f_bench.json
