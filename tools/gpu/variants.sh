#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
run() { RPGO_LIB_PATH=$1 timeout 300 python tools/k3_probe.py $2 $3 $4 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; }
P=kimera-rpgo_b200/librpgo_b200.so
V=kimera-rpgo_b200/variants
run $P 3 0 20000; run $P 3 0 50000; run $P 2 0 20000; run $P 3 1 20000; run $P 2 1 20000; run $P 2 0 50000; run $P 3 1 50000
for L in $V/librpgo_b200_d2m2.so $V/librpgo_b200_d2m4.so $V/librpgo_b200_d2nou.so; do run $L 2 0 20000; done
run $V/librpgo_b200_s3m2.so 3 1 20000
cat gpurun_out/variants.jsonl; tail -3 gpurun_out/variants.err
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "not stated and not config5 and not config4 and not clique" > gpurun_out/n1e_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/n1e_pytest.log
