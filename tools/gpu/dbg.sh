#!/bin/bash
mkdir -p gpurun_out
RPGO_CLIQUE_HOSTLOOP=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "clique" > gpurun_out/dbg1.log 2>&1
tail -15 gpurun_out/dbg1.log
timeout 3000 python -m pytest tests -m gpu -q -x --deselect tests/test_multigpu.py > gpurun_out/n1b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n1b_pytest.log
tail -30 gpurun_out/n1b_pytest.log
