#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the torchrun bench at N ranks and the drop-in binary with N rank threads.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n$N.txt
echo "== bench N=1 (same box)"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1.txt
echo "== ecloop add, 2^38 keys, -gpus $N"
timeout 300 ecloop_b200/host/ecloop add -f tests/golden/btc-puzzles-hash -r 400000000000000000:400000003fffffffff -q -o /dev/null -gpus $N 2>&1 | tr '\r' '\n' | tail -1 | tee gpurun_out/cli_add_n$N.txt
echo "== ecloop add 8000:fffffff -gpus $N (13 keys)"
rm -f /tmp/o.txt; timeout 120 ecloop_b200/host/ecloop add -f tests/golden/btc-puzzles-hash -r 8000:fffffff -q -o /tmp/o.txt -gpus $N 2>&1 | tr '\r' '\n' | tail -1
sort /tmp/o.txt | md5sum; wc -l /tmp/o.txt
