#!/usr/bin/env python
"""Wall-clock phases of the end-to-end path (run under gpurun)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("kimera-rpgo_b200"); synth = importlib.import_module("kimera-rpgo_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = pkg.PcmGpu(3, 0, odom_threshold=-1.0, lc_threshold=5.0, traj_mode=int(os.environ.get("TRAJ_MODE", "0"))); p.sync(); t1 = time.perf_counter()
    p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"]); p.sync(); t2 = time.perf_counter()
    p.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"]); p.sync(); t3 = time.perf_counter()
    sz, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU); p.sync(); t4 = time.perf_counter()
    p.close(); t5 = time.perf_counter()
    print("iter %d: create %.1f ms | odom_append %.1f | lc_append %.1f | find_inliers %.1f | close %.1f | total %.1f"
          % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3))
