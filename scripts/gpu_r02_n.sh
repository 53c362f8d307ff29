# groups of a mixed batch side by side: tests, then the full C5 sweep (with the 16,773,120-byte leg)
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_postproc.py -x -q -k "mixed_method or multi_segment or golden or restore_what" > gpurun_out/r02n_tests.log 2>&1; tail -3 gpurun_out/r02n_tests.log
( time timeout 1500 python bench.py --config C5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n_c5.json 2> gpurun_out/r02n_c5.err ) 2> gpurun_out/r02n_c5.time; cat gpurun_out/r02n_c5.time; tail -3 gpurun_out/r02n_c5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02n_c5.json"))
for r in d["configs"]["C5"]["sweep"]:
    print(r["block_bytes"], r["blocks_per_gpu"], round(r["decompress_e2e_value"], 1), r["round_trip_identical"])
PY
