#!/bin/bash
# GPU visit A of round 2: parity of the new planner / filters / mul path, headline bench, add-kernel variants.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_smi.txt 2>&1; nproc >> gpurun_out/a_smi.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke-only 2>&1 | tail -3 | tee gpurun_out/a_smoke.txt
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x -rf 2>&1 | tail -40 | tee gpurun_out/a_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/a_bench.txt
echo "== variants default"; timeout 900 python tools/prof_variants.py 32 2>&1 | tee gpurun_out/a_variants_default.txt
for v in spall nw1; do
  echo "== variants $v"; ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so timeout 900 python tools/prof_variants.py 32 2 3 5 7 2>&1 | tee gpurun_out/a_variants_$v.txt
done
echo "== mul"; timeout 300 python tools/prof_mul.py 22 2>&1 | tee gpurun_out/a_mul.txt
