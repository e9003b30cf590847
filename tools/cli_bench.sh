#!/bin/bash
# The drop-in binary end to end on the GPU box: add over 2^37 keys (configs[1] prefix) and mul over N random keys
# (BASELINE configs[2] shape: hex private keys on stdin, -a cu), with the unmodified reference beside it on a subset.
set -u
mkdir -p gpurun_out
BIN=ecloop_b200/host/ecloop
REF=oracle/_ref/ecloop_ref
F=tests/golden/btc-puzzles-hash
N=${MUL_KEYS:-10000000}
{
echo "== add 2^37 keys, 1 GPU"
timeout 300 $BIN add -f $F -r 400000000000000000:400000001fffffffff -q -o /dev/null -gpus 1 2>&1 | tr '\r' '\n' | tail -1
echo "== mul: generating $N keys"
python - <<PY
import random, sys
sys.path.insert(0, ".")
import ecloop_b200 as E
r = random.Random(3)
picked = []
with open("/tmp/mul_keys.txt", "w") as f:
    for i in range(0, $N, 100000):
        ks = [r.getrandbits(256) for _ in range(min(100000, $N - i))]
        picked += ks[::10000]
        f.write("".join("%064x\n" % k for k in ks))
# filter = hash160 of every 10000th key (even pick -> compressed, odd pick -> uncompressed): SURVEY 8d config 3
with E.Device(0) as d:
    h33, h65 = d.hash160(d.scalar_mul(picked))
with open("/tmp/mul_filter.txt", "w") as f:
    f.write("".join((h33[i] if i % 2 == 0 else h65[i]) + "\n" for i in range(len(picked))))
print("planted", len(picked))
PY
ls -la /tmp/mul_keys.txt
echo "== ours: mul -a cu, all $N keys"
( export ECLOOP_VERBOSE=1; time timeout 600 $BIN mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_ours.txt -gpus 1 < /tmp/mul_keys.txt ) 2>&1 | tr '\r' '\n' | grep -E "Mkeys|real|stages" | tail -3
echo "== ours: mul -a cu -t 1 (one parser thread)"
( time timeout 600 $BIN mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_ours1.txt -gpus 1 -t 1 < /tmp/mul_keys.txt ) 2>&1 | tr '\r' '\n' | grep -E "Mkeys|real" | tail -2
echo "== reference: mul -a cu, first 1000000 keys, -t $(nproc)"
head -1000000 /tmp/mul_keys.txt > /tmp/mul_keys_1m.txt
( time timeout 600 $REF mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_ref.txt < /tmp/mul_keys_1m.txt ) 2>&1 | tr '\r' '\n' | grep -E "Mkeys|real" | tail -2
head -1000000 /tmp/mul_keys.txt | timeout 300 $BIN mul -f /tmp/mul_filter.txt -a cu -q -o /tmp/mul_ours_1m.txt -gpus 1 >/dev/null 2>&1
echo "found lines (ours 1M / ref 1M): $(wc -l < /tmp/mul_ours_1m.txt) / $(wc -l < /tmp/mul_ref.txt); identical sorted: $(cmp <(sort /tmp/mul_ours_1m.txt) <(sort /tmp/mul_ref.txt) && echo yes)"
} 2>&1 | tee gpurun_out/cli_bench.txt
