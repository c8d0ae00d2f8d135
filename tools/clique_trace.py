#!/usr/bin/env python
"""Trace the heuristic-clique rounds at a given size (run under gpurun with RPGO_CLIQUE_TRACE=1)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("kimera-rpgo_b200"); synth = importlib.import_module("kimera-rpgo_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
p = pkg.PcmGpu(3, 0, odom_threshold=-1.0, lc_threshold=5.0)
p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
p.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"]); p.sync()
for it in range(2):
    t0 = time.perf_counter(); sz, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU); p.sync(); t1 = time.perf_counter()
    print("n=%d clique size %d in %.1f ms" % (n, sz, (t1 - t0) * 1e3), flush=True)
deg = p.degrees(0)
import numpy as np
print("degree percentiles", np.percentile(deg, [1, 25, 50, 75, 99]).tolist(), "max", int(deg.max()))
