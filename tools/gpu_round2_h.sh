#!/bin/bash
# GPU visit H: K2a with branch-free products A/B, -a cu pipelined (branch-free) vs NW=1, new mul tests.
set -u
mkdir -p gpurun_out
for v in "" k2bf0; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== mul ${v:-default(bf)}"; timeout 300 python tools/prof_mul.py 22 2>&1 | tail -2 | tee gpurun_out/h_mul_${v:-bf}.txt
done
for v in "" spcu; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== -a cu ${v:-default(NW=1)}"; timeout 600 python tools/prof_add.py 31 3 2>&1 | tail -1 | tee gpurun_out/h_cu_${v:-nw1}.txt
done
unset ECLOOP_B200_LIB
echo "== pytest mul + prims"; timeout 1200 python -m pytest tests/test_gpu_mul.py tests/test_gpu_prims.py tests/test_gpu_cli.py -k "mul or prim or scalar or fp_" -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/h_pytest.txt
