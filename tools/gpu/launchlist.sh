#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -c 300 gpurun_out/launch_bench.log; wc -l gpurun_out/r2_launches.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:traj_fold -c 1 -o gpurun_out/r2_fold_pipelined_3d_v2 -f python tools/fold_probe.py 50000 3 > gpurun_out/fold_ncu.log 2>&1; tail -2 gpurun_out/fold_ncu.log
