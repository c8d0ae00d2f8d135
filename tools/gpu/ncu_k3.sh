#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:pairwise_grouped -c 1 -o gpurun_out/r2_k3_pcm3d_50k_v3 python tools/k3_probe.py 3 0 50000 1 > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
timeout 600 $NCU -k regex:pairwise_grouped -c 1 -o gpurun_out/r2_k3_pcm2d_50k_v3 python tools/k3_probe.py 2 0 50000 1 > gpurun_out/ncu_k3_2d.log 2>&1; tail -1 gpurun_out/ncu_k3_2d.log
ls -la gpurun_out/*.ncu-rep
