#!/usr/bin/env python
"""K1 probe: odom_append wall time for 2D / 3D chains (run under gpurun; RPGO_FOLD_V1=1 selects the one-warp kernel)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("kimera-rpgo_b200"); synth = importlib.import_module("kimera-rpgo_b200.synth")
for d, gph in [(2, synth.config3(seed=2, P=10000, n=100)), (3, synth.config2(seed=1, P=10000, n=100))]:
    arr = synth.as_arrays(gph)
    for it in range(3):
        p = pkg.PcmGpu(d, 0, odom_threshold=-1.0, lc_threshold=5.0)
        p.sync(); t0 = time.perf_counter()
        p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"]); p.sync()
        t1 = time.perf_counter()
        print("d=%d iter %d odom_append(%d steps) %.2f ms" % (d, it, len(arr["o_prev"]), (t1 - t0) * 1e3), flush=True)
        p.close()
