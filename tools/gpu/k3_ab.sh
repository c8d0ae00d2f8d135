#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'P'
import ctypes as C, importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module("kimera-rpgo_b200")
lib = pkg._capi.load()
m, c = C.c_ulonglong(), C.c_ulonglong()
rc = lib.rpgo_debug_check_fastmath(20_000_000_000, 987654321, C.byref(m), C.byref(c))
print("fastmath rc", rc, "mismatches", m.value, "checked", c.value)
P
for a in "3 0 20000" "3 0 50000" "2 0 20000"; do timeout 200 python tools/k3_probe.py $a; done
for x in kimera-rpgo_b200/variants/librpgo_b200_k3*.so; do [ -f $x ] && for a in "3 0 20000" "3 0 50000"; do RPGO_LIB_PATH=$x timeout 200 python tools/k3_probe.py $a; done; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q --deselect tests/test_gpu_round2.py::test_config4_stated_size_groups_and_fmc 2>&1 | tail -2
