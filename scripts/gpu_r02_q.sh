# NVRTC-compiled PCOMP: the post-processor tests (native / compiled / interpreted must agree), the foreign program, mixed + multi-segment
cd /root/repo
timeout 420 python -m pytest tests/test_gpu_postproc.py -x -q > gpurun_out/r02q_post.log 2>&1; tail -6 gpurun_out/r02q_post.log
