#!/bin/bash
# A/B of prebuilt library variants (build_variants/*.so) on the bench workload: 1607 blocks x 1,044,480 B, mid.cfg.
for v in "$@"; do
  cp build_variants/$v.so zpaqsharp_b200/libzpaqb200.so
  echo "== $v" >> gpurun_out/variants.log
  timeout 120 python scripts/ab_duo.py 1607 1044480 2 2>&1 | grep "^duo=" | tail -1 >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
