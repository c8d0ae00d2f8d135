#!/bin/bash
# A/B of the K1 fold kernels: old batched kernel, pipelined kernel with the exchange variants
mkdir -p gpurun_out
V=kimera-rpgo_b200/variants
{
echo "== batched (round-1 kernel)"; RPGO_FOLD_V2=1 timeout 120 python tools/fold_probe.py 50000
echo "== pipelined, XCHG=1 (product)"; timeout 120 python tools/fold_probe.py 50000
for x in $V/librpgo_b200_fx*.so; do echo "== pipelined, $x"; RPGO_LIB_PATH=$x timeout 120 python tools/fold_probe.py 50000; done
} > gpurun_out/fold_ab.log 2>&1
cat gpurun_out/fold_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or scenarios or scan_mode or closure_before or landmark or config3" 2>&1 | tail -3
