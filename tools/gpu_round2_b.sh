#!/bin/bash
# GPU visit B of round 2: parity with the new table / squaring / NW=1 probe pipe, variants, full bench line, ncu.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x -rf 2>&1 | tail -15 | tee gpurun_out/b_pytest.txt
echo "== variants default"; timeout 900 python tools/prof_variants.py 32 2>&1 | tee gpurun_out/b_variants_default.txt
echo "== nosqr"; ECLOOP_B200_LIB=build/variants/libecloop_b200_nosqr.so timeout 600 python tools/prof_add.py 32 1 2>&1 | tee gpurun_out/b_nosqr.txt
echo "== sqr"; timeout 600 python tools/prof_add.py 32 1 2>&1 | tee gpurun_out/b_sqr.txt
for v in "" w16 w24 w26 mulnw1; do
  echo "== mul ${v:-default}"
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  timeout 300 python tools/prof_mul.py 22 2>&1 | tee gpurun_out/b_mul_${v:-default}.txt
done
unset ECLOOP_B200_LIB
echo "== bench full"; timeout 1500 python bench.py --steps 5 --warmup 3 2> gpurun_out/b_bench_err.txt | tail -1 | tee gpurun_out/b_bench.json | cut -c1-3000
tail -5 gpurun_out/b_bench_err.txt
echo "== ncu launch list (headline + mul + endo legs)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rnd-windows 0 --mul-keys 4194304 --endo-steps 1 > gpurun_out/b_bench_under_ncu.log 2>&1
tail -2 gpurun_out/b_bench_under_ncu.log | cut -c1-400
echo "== ncu full: headline at the bench's launch shape"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:add_kernel -s 1 -c 1 -f -o gpurun_out/b_prof_add python tools/prof_add.py 32 > gpurun_out/b_prof_add.log 2>&1; tail -2 gpurun_out/b_prof_add.log
echo "== ncu full: mul kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mul_ -s 2 -c 2 -f -o gpurun_out/b_prof_mul python tools/prof_mul.py 22 > gpurun_out/b_prof_mul.log 2>&1; tail -2 gpurun_out/b_prof_mul.log
echo "== ncu full: -endo instance, 4 GiB filter in HBM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:add_kernel -s 3 -c 1 -f -o gpurun_out/b_prof_endo python tools/prof_bloom.py 32 30 5 > gpurun_out/b_prof_endo.log 2>&1; tail -2 gpurun_out/b_prof_endo.log
for n in add mul endo; do ncu -i gpurun_out/b_prof_$n.ncu-rep --page raw --csv > gpurun_out/b_prof_${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -30
