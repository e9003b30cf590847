#!/bin/bash
# One GPU-box visit: parity tests, peaks, bench, launch list. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== peaks"; timeout 300 python -c "
import json, ecloop_b200 as E
d = E.Device(0); print(json.dumps(d.peak_bench()))" 2>&1 | tee gpurun_out/peak.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rf 2>&1 | tail -60 | tee gpurun_out/pytest.txt
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.txt
