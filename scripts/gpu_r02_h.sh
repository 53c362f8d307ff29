# compact ICM map: 12 decoder warps per SM (1776 blocks); ncu --set full of the decoder at 2 blocks per SM
cd /root/repo
timeout 300 python scripts/ab_dec.py 1776 200000 2 mixed 1 > gpurun_out/r02h_1776.log 2>&1; cat gpurun_out/r02h_1776.log
timeout 300 python scripts/ab_dec.py 64 30000 1 mixed 1 >> gpurun_out/r02h_1776.log 2>&1; tail -2 gpurun_out/r02h_1776.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_aot2_f -c 1 -o gpurun_out/r02h_fdec_296x50k -f python scripts/ab_dec.py 296 50000 2 mixed 1 > gpurun_out/r02h_ncu.log 2>&1; tail -4 gpurun_out/r02h_ncu.log; ls -la gpurun_out/*.ncu-rep
