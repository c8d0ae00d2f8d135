#!/usr/bin/env python
"""Online use: closures arrive one at a time (RobustSolver::update per keyframe).  Times one update at n0 stored closures
through the array API (lc_append of one closure + incremental clique) — run under gpurun."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("kimera-rpgo_b200"); synth = importlib.import_module("kimera-rpgo_b200.synth")
for n0 in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1000", "10000", "50000"])]:
    extra = 40
    arr = synth.as_arrays(synth.config2(seed=4, P=max(2500, n0), n=n0 + extra))
    p = pkg.PcmGpu(3, 0, odom_threshold=-1.0, lc_threshold=5.0, incremental=True)
    p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    p.lc_append_arrays(arr["l_from"][:n0], arr["l_to"][:n0], arr["l_pose"][:n0], arr["l_cov"][:n0])
    size, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU)
    p.sync()
    ta, tc = [], []
    for k in range(n0, n0 + extra):
        t0 = time.perf_counter()
        p.lc_append_arrays(arr["l_from"][k:k + 1], arr["l_to"][k:k + 1], arr["l_pose"][k:k + 1], arr["l_cov"][k:k + 1])
        p.sync()
        t1 = time.perf_counter()
        s2, ids2, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU_INCREMENTAL, 1, size)
        if s2 > size:
            size = s2
        p.sync()
        t2 = time.perf_counter()
        ta.append((t1 - t0) * 1e3); tc.append((t2 - t1) * 1e3)
    print("n0=%d: lc_append(1 closure) median %.3f ms (min %.3f max %.3f) | incremental clique median %.3f ms (max %.3f) | clique size %d"
          % (n0, np.median(ta), min(ta), max(ta), np.median(tc), max(tc), size), flush=True)
    p.close()
