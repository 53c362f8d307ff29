#!/bin/bash
# Build library variants for scripts/ab_variants.sh: each argument is NAME=FLAGS (FLAGS go to nvcc through
# ZPQ_EXTRA_NVCC_FLAGS), e.g.   scripts/build_variants.sh base= rot=-DZPQ_DUO_ROTATE "o40123=-DZPQ_DUO_ORDER=40123"
# The default library is rebuilt at the end so that the tree is left as committed.
set -u
cd "$(dirname "$0")/.."
mkdir -p build_variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  if ZPQ_EXTRA_NVCC_FLAGS="$flags" timeout 900 python zpaqsharp_b200/build.py --force > /tmp/zpq_variant_build.log 2>&1; then
    cp zpaqsharp_b200/libzpaqb200.so "build_variants/$name.so"; echo "built $name ($flags)"
  else
    echo "FAILED $name"; grep -i -m5 -A3 "error" /tmp/zpq_variant_build.log
  fi
done
timeout 900 python zpaqsharp_b200/build.py --force > /dev/null 2>&1 && echo "default library restored"
