#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
run() { RPGO_LIB_PATH=$1 timeout 300 python tools/k3_probe.py $2 $3 $4 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; }
P=kimera-rpgo_b200/librpgo_b200.so
V=kimera-rpgo_b200/variants
for L in $P $V/librpgo_b200_g2.so $V/librpgo_b200_g4.so $V/librpgo_b200_g6.so $V/librpgo_b200_seg1008.so $V/librpgo_b200_seg252.so; do run $L 3 0 20000; done
cat gpurun_out/variants.jsonl; tail -3 gpurun_out/variants.err
