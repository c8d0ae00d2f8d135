#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
run() { RPGO_LIB_PATH=$1 timeout 300 python tools/k3_probe.py $2 $3 $4 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; }
P=kimera-rpgo_b200/librpgo_b200.so
V=kimera-rpgo_b200/variants
for n in 20000 50000; do run $P 3 0 $n; run $V/librpgo_b200_r1.so 3 0 $n; run $V/librpgo_b200_r2.so 3 0 $n; run $V/librpgo_b200_r12.so 3 0 $n; done
run $P 2 0 20000; run $V/librpgo_b200_d2_minb3.so 2 0 20000; run $V/librpgo_b200_d2_tw12.so 2 0 20000; run $V/librpgo_b200_d2_tw12b1.so 2 0 20000
run $P 3 1 20000; run $V/librpgo_b200_s3_minb3.so 3 1 20000; run $V/librpgo_b200_s3_minb4.so 3 1 20000; run $V/librpgo_b200_s3_tw12.so 3 1 20000
run $P 2 1 20000
cat gpurun_out/variants.jsonl
tail -5 gpurun_out/variants.err
python - <<'PY'
import importlib, sys, json, time, ctypes as C
sys.path.insert(0,'.')
import numpy as np, torch
pkg = importlib.import_module("kimera-rpgo_b200")
h = pkg.PcmGpu(3,0)
rng = np.random.default_rng(1)
n=50000
W=(n+63)//64
rows = rng.integers(0, 2**63, size=(n, W), dtype=np.int64).view(np.uint64)
g = C.c_int32(-1)
h._check(h.lib.rpgo_debug_load_group(h.h, ord('y'), ord('z'), n, rows.ctypes.data_as(pkg._capi.c_u64p), W, C.byref(g)), "load")
st = torch.cuda.ExternalStream(h.stream_ptr())
for which,name in ((0,'mirror'),(1,'degree')):
    best=1e9
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st); h.debug_pass(g.value, which); e1.record(st)
        e1.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(name, best, 'ms', n*n/8/best/1e6, 'GB/s')
PY
