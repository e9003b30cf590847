#!/bin/bash
# `make blf` (Makefile:35-44 of the reference) end to end with a filter large enough to live in HBM (270 MB):
# blf-gen with our binary, then add / mul against it with ours and with the unmodified reference.
set -u
mkdir -p gpurun_out
B=ecloop_b200/host/ecloop; R=oracle/_ref/ecloop_ref; G=tests/golden
{
rm -f /tmp/big.blf /tmp/o_add.txt /tmp/r_add.txt /tmp/o_mul.txt /tmp/r_mul.txt
$B blf-gen -n 50000000 -o /tmp/big.blf < $G/btc-puzzles-hash | tail -2
$B blf-gen -n 50000000 -o /tmp/big.blf < $G/btc-bw-hash | tail -1
ls -la /tmp/big.blf
$B add -f /tmp/big.blf -r 8000:ffffff -q -o /tmp/o_add.txt -gpus 1 2>&1 | tr '\r' '\n' | tail -1
$R add -f /tmp/big.blf -r 8000:ffffff -q -o /tmp/r_add.txt 2>&1 | tr '\r' '\n' | tail -1
echo "add: ours $(sort /tmp/o_add.txt | md5sum | cut -c1-32) ref $(sort /tmp/r_add.txt | md5sum | cut -c1-32) lines $(wc -l < /tmp/o_add.txt)"
$B mul -f /tmp/big.blf -a cu -q -o /tmp/o_mul.txt -gpus 1 < $G/btc-bw-priv 2>&1 | tr '\r' '\n' | tail -1
$R mul -f /tmp/big.blf -a cu -q -o /tmp/r_mul.txt < $G/btc-bw-priv 2>&1 | tr '\r' '\n' | tail -1
echo "mul: ours $(sort /tmp/o_mul.txt | md5sum | cut -c1-32) ref $(sort /tmp/r_mul.txt | md5sum | cut -c1-32) lines $(wc -l < /tmp/o_mul.txt)"
} 2>&1 | tee gpurun_out/blf_flow.txt
