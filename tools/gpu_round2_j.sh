#!/bin/bash
# GPU visit J (gpurun --gpus N): the driver's scaling configuration — torchrun bench at N ranks with all legs — plus the
# drop-in binary on N GPUs against a 6 GiB .blf (streamed load timing) and the multi-GPU tests.
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/j_gpus.txt; nproc | tee -a gpurun_out/j_gpus.txt
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/j_bench_n${N}_err.txt | tail -1 > gpurun_out/j_bench_n$N.json
cut -c1-300 gpurun_out/j_bench_n$N.json; tail -2 gpurun_out/j_bench_n${N}_err.txt | cut -c1-300
echo "== 6 GiB .blf on $N GPUs"
python - <<'PY'
import struct, sys
sys.path.insert(0, ".")
import ecloop_b200 as E
size = 6 * 2**30 // 8
with E.Device(0) as d:
    d.filter_generate(size, 0.37, 4)
    with open("/dev/shm/big.blf", "wb") as f:
        f.write(struct.pack("<IIQ", 0x45434246, 1, size))
        for off in range(0, size, 1 << 25):
            f.write(d.filter_read(off, min(1 << 25, size - off)).tobytes())
print("written")
PY
for mode in "" 1; do
  ( if [ -n "$mode" ]; then export ECLOOP_BLF_PEER=1; fi; export ECLOOP_VERBOSE=1
    time ecloop_b200/host/ecloop add -f /dev/shm/big.blf -endo -r 400000000000000000:4000001fffffffffff -q -o /dev/null -gpus $N ) 2>&1 | tr '\r' '\n' | grep -E "filter:|Mkeys/s|real" | tail -4 | tee gpurun_out/j_blf_load_n${N}_peer${mode:-0}.txt
done
rm -f /dev/shm/big.blf
echo "== multi-GPU tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "all_gpus or peer_copy" 2>&1 | tail -2 | tee gpurun_out/j_pytest.txt
