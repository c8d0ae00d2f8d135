#!/bin/bash
# ncu --set full captures of every kernel family (one launch each), reports into gpurun_out/
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:pairwise_grouped -c 1 -o gpurun_out/r2_k3_pcm3d_50k python tools/k3_probe.py 3 0 50000 1 > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
timeout 600 $NCU -k regex:pairwise_grouped -c 1 -o gpurun_out/r2_k3_pcm2d_20k python tools/k3_probe.py 2 0 20000 1 > gpurun_out/ncu_k3_2d.log 2>&1; tail -1 gpurun_out/ncu_k3_2d.log
timeout 600 $NCU -k regex:pairwise_grouped -c 1 -o gpurun_out/r2_k3_simple3d_20k python tools/k3_probe.py 3 1 20000 1 > gpurun_out/ncu_k3_s3.log 2>&1; tail -1 gpurun_out/ncu_k3_s3.log
timeout 600 $NCU -k regex:"mirror_tile|degree_kernel|heu_persistent" -c 6 -o gpurun_out/r2_bitset_clique_50k python tools/clique_probe.py 50000 > gpurun_out/ncu_clique.log 2>&1; tail -3 gpurun_out/ncu_clique.log
ls -la gpurun_out/*.ncu-rep
