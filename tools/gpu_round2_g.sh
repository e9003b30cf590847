#!/bin/bash
# GPU visit G of round 2: validation of the final build (tests, smoke, both bench arms), CTA-size / barrier-cadence
# variants of the headline kernel, ncu launch list + full captures of the final kernels.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke-only 2>&1 | tail -2 | tee gpurun_out/g_smoke.txt
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -rf 2>&1 | tail -6 | tee gpurun_out/g_pytest.txt
for v in "" t576 t640 sync2 sync4; do
  if [ -n "$v" ]; then export ECLOOP_B200_LIB=build/variants/libecloop_b200_$v.so; else unset ECLOOP_B200_LIB; fi
  echo "== add ${v:-default}"; timeout 600 python tools/prof_add.py 32 1 2>&1 | tail -1 | tee gpurun_out/g_add_${v:-default}.txt
done
unset ECLOOP_B200_LIB
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/g_bench_ref.json | cut -c1-400
echo "== bench"; timeout 1500 python bench.py 2>gpurun_out/g_bench_err.txt | tail -1 > gpurun_out/g_bench.json; cut -c1-400 gpurun_out/g_bench.json; tail -3 gpurun_out/g_bench_err.txt
echo "== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/g_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rnd-windows 0 --endo-steps 1 > gpurun_out/g_bench_under_ncu.log 2>&1
echo "== ncu full: headline"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:add_kernel -s 1 -c 1 -f -o gpurun_out/g_prof_add python tools/prof_add.py 32 > gpurun_out/g_prof_add.log 2>&1; tail -1 gpurun_out/g_prof_add.log
echo "== ncu full: mul"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:mul_ -s 2 -c 2 -f -o gpurun_out/g_prof_mul python tools/prof_mul.py 22 > gpurun_out/g_prof_mul.log 2>&1; tail -1 gpurun_out/g_prof_mul.log
echo "== ncu full: -a cu -endo, 4 GiB filter"; timeout 900 ncu --set full --clock-control none -k regex:add_kernel -s 3 -c 1 -f -o gpurun_out/g_prof_cuendo python tools/prof_bloom.py 32 28 7 > gpurun_out/g_prof_cuendo.log 2>&1; tail -1 gpurun_out/g_prof_cuendo.log
for n in add mul cuendo; do ncu -i gpurun_out/g_prof_$n.ncu-rep --page raw --csv > gpurun_out/g_prof_${n}_raw.csv 2>/dev/null; done
ncu -i gpurun_out/g_prof_add.ncu-rep --page source --csv > gpurun_out/g_prof_add_source.csv 2>/dev/null
rm -f gpurun_out/g_prof_mul.ncu-rep gpurun_out/g_prof_cuendo.ncu-rep
ls -la gpurun_out | grep " g_"
