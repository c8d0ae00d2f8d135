#!/bin/bash
mkdir -p gpurun_out
P=kimera-rpgo_b200/librpgo_b200.so
V=kimera-rpgo_b200/variants
for L in $P $V/librpgo_b200_heu5.so $V/librpgo_b200_heu6.so $V/librpgo_b200_heu8.so $V/librpgo_b200_mir100.so $V/librpgo_b200_mir75.so; do
  echo "== $L"; RPGO_LIB_PATH=$L timeout 300 python tools/clique_probe.py 50000
done
for L in $P $V/librpgo_b200_heu6.so $V/librpgo_b200_heu8.so; do echo "== $L 200k"; RPGO_LIB_PATH=$L timeout 300 python tools/clique_probe.py 200000; done
timeout 600 ncu --set full --clock-control none -f -k regex:"mirror_tile|heu_persistent" -c 4 -o gpurun_out/r2_mirror_heu_v2 python tools/clique_probe.py 50000 > gpurun_out/ncu_clique2.log 2>&1; tail -2 gpurun_out/ncu_clique2.log
