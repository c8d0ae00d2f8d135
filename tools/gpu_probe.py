#!/usr/bin/env python
"""Quick GPU probe (run under gpurun): FP64 DFMA peak and first K3 timings.  Not a bench line."""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
lib = pkg._capi.load()
tf = C.c_double()
print("fp64_peak rc", lib.rpgo_fp64_peak(0, C.byref(tf)), "TFLOP/s", tf.value)
print(torch.cuda.get_device_name(0))
kernels = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1"])]
sizes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["2000", "8000"])]
for n in sizes:
    gph = synth.config2(seed=4, P=max(2500, n), n=n)
    for kern in kernels:
        g = pkg.PcmGpu(3, 0, odom_threshold=-1, lc_threshold=5.0, kernel=kern)
        t0 = time.time()
        g.update(gph["odom"], gph["values"])
        t1 = time.time()
        g.update(gph["lcs"], [])
        t2 = time.time()
        st = torch.cuda.ExternalStream(g.stream_ptr())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        ms = []
        for rep in range(3):
            with torch.cuda.stream(st):
                e0.record(st)
                g.recompute(0, 0)
                e1.record(st)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        pairs = n * (n - 1) / 2
        best = min(ms)
        print("n=%d kernel=%d odom %.3fs update %.3fs | recompute ms %s | %.3e pairs/s | %.2f TFLOP/s (6.2 kflop/pair) | inliers %d"
              % (n, kern, t1 - t0, t2 - t1, ["%.2f" % m for m in ms], pairs / (best * 1e-3), pairs * 6.2e3 / (best * 1e-3) / 1e12,
                 g.num_inliers()))
        g.close()
