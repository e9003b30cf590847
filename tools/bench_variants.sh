#!/bin/bash
# Time every build/variants/*.so on the GPU box: smoke (parity vs oracle) + add_kernel Mkeys/s over 2^LOG2 keys.
set -u
mkdir -p gpurun_out
LOG2=${LOG2:-31}
for so in build/variants/libecloop_b200_*.so; do
  tag=$(basename $so .so); tag=${tag#libecloop_b200_}
  echo "== $tag"
  ECLOOP_B200_LIB=$so timeout 300 python __graft_entry__.py --smoke-only 2>&1 | tail -1
  ECLOOP_B200_LIB=$so timeout 300 python tools/prof_add.py $LOG2 2>&1 | tail -2
done 2>&1 | tee gpurun_out/variants.txt
