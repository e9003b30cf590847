#!/bin/bash
# GPU visit I: K2a window-entry prefetch A/B, W = 22 / 24 / 26 with prefetch.
set -u
mkdir -p gpurun_out
for v in "" k2nopf w22 w26; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== mul ${v:-default(W24,prefetch)}"; timeout 300 python tools/prof_mul.py 22 2>&1 | tail -3 | tee gpurun_out/i_mul_${v:-default}.txt
done
unset ECLOOP_B200_LIB
echo "== pytest mul"; timeout 1200 python -m pytest tests/test_gpu_mul.py tests/test_gpu_prims.py -q --timeout 900 2>&1 | tail -2 | tee gpurun_out/i_pytest.txt
