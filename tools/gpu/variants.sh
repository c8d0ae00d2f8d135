#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
run() { RPGO_LIB_PATH=$1 timeout 300 python tools/k3_probe.py $2 $3 $4 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; }
P=kimera-rpgo_b200/librpgo_b200.so
B=kimera-rpgo_b200/variants/librpgo_b200_base.so
for L in $B $P; do run $L 3 0 20000; run $L 3 0 50000; run $L 2 0 20000; run $L 3 1 20000; run $L 2 1 20000; done
cat gpurun_out/variants.jsonl; tail -3 gpurun_out/variants.err
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "not stated and not config5 and not config4 and not clique" > gpurun_out/n1e_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/n1e_pytest.log
