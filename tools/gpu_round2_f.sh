#!/bin/bash
# GPU visit F of round 2: pinned field multiplications A/B, the drop-in binary's mul end to end (mapped stdin).
set -u
mkdir -p gpurun_out
for v in "" pins; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== add ${v:-default}"
  for f in 1 2; do timeout 600 python tools/prof_add.py $((33 - f)) $f 2>&1 | tail -1; done | tee gpurun_out/f_add_${v:-default}.txt
  timeout 600 python tools/prof_variants.py 32 1 2 2>&1 | grep hbm | tee -a gpurun_out/f_add_${v:-default}.txt
done
export ECLOOP_B200_LIB=build/variants/libecloop_b200_pins.so
echo "== parity of the pinned build"; timeout 1200 python -m pytest tests/test_gpu_add.py -q --timeout 900 -x 2>&1 | tail -3 | tee gpurun_out/f_pytest_pins.txt
unset ECLOOP_B200_LIB
echo "== cli bench"; bash tools/cli_bench.sh > /dev/null 2>&1; cp gpurun_out/cli_bench.txt gpurun_out/f_cli_bench.txt; cat gpurun_out/f_cli_bench.txt
echo "== mul through a pipe"
( export ECLOOP_VERBOSE=1; time (cat /tmp/mul_keys.txt | ecloop_b200/host/ecloop mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_pipe.txt -gpus 1) ) 2>&1 | tr '\r' '\n' | grep -E "Mkeys|real|stages" | tail -3 | tee gpurun_out/f_cli_mul_pipe.txt
