#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch list + full capture. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke-only 2>&1 | tail -3 | tee gpurun_out/smoke.txt
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rf 2>&1 | tail -30 | tee gpurun_out/pytest.txt
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.txt
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.txt
if [ "${WITH_PROF:-1}" = "1" ]; then
  PROF_LOG2=${PROF_LOG2:-29} bash tools/gpu_prof.sh
  ncu -i gpurun_out/prof_add.ncu-rep --page raw --csv > gpurun_out/prof_add_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_add.ncu-rep --page source --csv > gpurun_out/prof_add_source.csv 2>/dev/null
fi
ls -la gpurun_out
