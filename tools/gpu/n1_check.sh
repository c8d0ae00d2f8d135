#!/bin/bash
# single-GPU: full gpu test suite, bench, ncu capture of the FP64 peak micro-benchmark
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/n1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n1_pytest.log
tail -8 gpurun_out/n1_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err
echo "bench rc=$?" | tee -a gpurun_out/n1_bench.err
tail -c 1500 gpurun_out/n1_bench.json
timeout 600 ncu --set full --clock-control none -k regex:dfma_peak -c 2 -o gpurun_out/r2_dfma_peak -f python -c "
import importlib, ctypes as C
pkg = importlib.import_module('kimera-rpgo_b200')
lib = pkg._capi.load(); tf = C.c_double(); lib.rpgo_fp64_peak(0, C.byref(tf)); print('peak', tf.value)" > gpurun_out/n1_ncu_dfma.log 2>&1
tail -3 gpurun_out/n1_ncu_dfma.log
