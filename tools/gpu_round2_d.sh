#!/bin/bash
# GPU visit D of round 2 (gpurun --gpus N, N >= 2): every test incl. the multi-GPU ones, torchrun bench at N ranks with
# the secondary legs, the drop-in binary on N GPUs, the streamed .blf load on N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/d_gpus.txt; nproc | tee -a gpurun_out/d_gpus.txt
echo "== pytest (all GPUs visible)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -rf 2>&1 | tail -8 | tee gpurun_out/d_pytest.txt
echo "== bench N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/d_bench_n${N}_err.txt | tail -1 > gpurun_out/d_bench_n$N.json
cut -c1-600 gpurun_out/d_bench_n$N.json; tail -3 gpurun_out/d_bench_n${N}_err.txt
echo "== bench N=1 skipped"; if false; then
echo "== bench N=1 (same box)"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/d_bench_n1.json; cut -c1-300 gpurun_out/d_bench_n1.json
fi
echo "== ecloop add, 2^38 keys, -gpus $N"
timeout 300 ecloop_b200/host/ecloop add -f tests/golden/btc-puzzles-hash -r 400000000000000000:400000003fffffffff -q -o /dev/null -gpus $N 2>&1 | tr '\r' '\n' | tail -1 | tee gpurun_out/d_cli_add_n$N.txt
echo "== 6 GiB .blf on $N GPUs: streamed load (all GPUs at once) vs GPU 0 + peer copies; add -endo 2^32 keys"
python - <<'PY'
import struct, sys, time
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
    time ecloop_b200/host/ecloop add -f /dev/shm/big.blf -endo -r 400000000000000000:40000000ffffffffff -q -o /dev/null -gpus $N ) 2>&1 | tr '\r' '\n' | grep -E "filter:|Mkeys/s|real" | tail -3 | tee gpurun_out/d_blf_load_n${N}_peer${mode:-0}.txt
done
rm -f /dev/shm/big.blf
echo "== rnd 24 windows -gpus $N"
ECLOOP_RND_WINDOWS=24 timeout 300 ecloop_b200/host/ecloop rnd -f tests/golden/btc-puzzles-hash -d 128:32 -a cu -gpus $N 2>&1 | tr '\r' '\n' | grep -E "Mkeys/s" | tail -1 | tee gpurun_out/d_rnd_n$N.txt
