# two GPUs: the in-library multi-device path (tests), then bench.py under torchrun with the library_multi_gpu leg
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02m_tests.log 2>&1; tail -5 gpurun_out/r02m_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-configs > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err; tail -3 gpurun_out/r02m_bench_n2.err; head -c 3000 gpurun_out/r02m_bench_n2.json
