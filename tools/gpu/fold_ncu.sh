#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:traj_fold -c 1 -o gpurun_out/r2_fold_pipelined_3d -f python tools/fold_probe.py 50000 3 > gpurun_out/fold_ncu.log 2>&1
tail -3 gpurun_out/fold_ncu.log
ncu --set full --import-source on --clock-control none -k regex:traj_fold -c 1 -o gpurun_out/r2_fold_pipelined_2d -f python tools/fold_probe.py 50000 2 >> gpurun_out/fold_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
