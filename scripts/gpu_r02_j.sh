# A/B of decoder variants at 12 blocks per SM: MATCH work fused into the look-up wait; without the 8-candidate L2 prefetch
cd /root/repo
cp zpaqsharp_b200/libzpaqb200.so /tmp/default.so
for v in fuse fuse_nopf8; do
  cp build_variants/$v.so zpaqsharp_b200/libzpaqb200.so
  echo "== $v" >> gpurun_out/r02j_variants.log
  timeout 200 python scripts/ab_dec.py 1776 200000 2 mixed 2 2>&1 | grep "^fdec=" >> gpurun_out/r02j_variants.log
done
cp /tmp/default.so zpaqsharp_b200/libzpaqb200.so
timeout 200 python scripts/ab_dec.py 64 30000 1 mixed 1 2>&1 | grep "^fdec=" >> gpurun_out/r02j_variants.log
cat gpurun_out/r02j_variants.log | cut -c1-200
