# after the compiled-PCOMP change: every GPU test except the long large-block ones, and the smoke entry
cd /root/repo
timeout 400 python -m pytest tests -m gpu -x -q -k "not large_blocks and not max_cfg_full and not mid_cfg_4mb and not every_block and not level5 and not postproc" > gpurun_out/r02r_tests.log 2>&1; tail -4 gpurun_out/r02r_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; tail -2 gpurun_out/r02r_smoke.log
