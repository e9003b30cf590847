#!/bin/bash
# One GPU-box visit for profiles: ncu launch list of the bench command + one --set full capture of add_kernel.
set -u
mkdir -p gpurun_out
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log
echo "== full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:add_kernel -s 1 -c 1 -f -o gpurun_out/prof_add \
  python tools/prof_add.py ${PROF_LOG2:-28} > gpurun_out/prof_add.log 2>&1
tail -5 gpurun_out/prof_add.log
ls -la gpurun_out
