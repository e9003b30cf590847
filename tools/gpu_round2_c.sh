#!/bin/bash
# GPU visit C of round 2: branch-free pipelined kernels A/B, L2 fetch granularity A/B, K2a tuning, CLI mul end to end,
# BASELINE configs[3] against the reference (6 GiB filter).
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x -rf 2>&1 | tail -8 | tee gpurun_out/c_pytest.txt
for v in "" nobf; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== add ${v:-default(bf)}"
  for f in 1 2; do timeout 600 python tools/prof_add.py $((33 - f)) $f 2>&1 | tail -1; done | tee gpurun_out/c_add_${v:-bf}.txt
done
unset ECLOOP_B200_LIB
echo "== hbm, L2 granularity set when the filter is loaded (default)"
timeout 900 python tools/prof_variants.py 32 1 3 5 7 2>&1 | grep hbm | tee gpurun_out/c_hbm_gran_filter.txt
echo "== hbm, L2 granularity set at ecl_open (round-1 behaviour)"
ECLOOP_B200_L2GRAN_AT_OPEN=1 timeout 900 python tools/prof_variants.py 32 1 3 5 7 2>&1 | grep hbm | tee gpurun_out/c_hbm_gran_open.txt
for g in "" 1; do
  echo "== dram bytes per launch, flags 5, gran_at_open=${g:-0}"
  ECLOOP_B200_L2GRAN_AT_OPEN=$g timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"add_kernel|cand_verify" -s 2 -c 2 --csv \
    python tools/prof_bloom.py 32 30 5 2>/dev/null | grep -E '^"' | cut -d, -f5,13- | tee gpurun_out/c_dram_gran${g:-0}.txt
done
for v in "" mb3 mb4 msync mb3sync; do
  echo "== mul ${v:-default}"
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  timeout 300 python tools/prof_mul.py 22 2>&1 | tail -2 | tee gpurun_out/c_mul_${v:-default}.txt
done
unset ECLOOP_B200_LIB
echo "== cli bench"; bash tools/cli_bench.sh > /dev/null 2>&1; cp gpurun_out/cli_bench.txt gpurun_out/c_cli_bench.txt; cat gpurun_out/c_cli_bench.txt
echo "== mul through a pipe"
( export ECLOOP_VERBOSE=1; time (cat /tmp/mul_keys.txt | ecloop_b200/host/ecloop mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_pipe.txt -gpus 1) ) 2>&1 | tr '\r' '\n' | grep -E "Mkeys|real|stages" | tail -3 | tee gpurun_out/c_cli_mul_pipe.txt
echo "== config 4 parity"; timeout 1500 python tools/config4_parity.py 30 6 2>&1 | tee gpurun_out/c_config4_parity.txt
