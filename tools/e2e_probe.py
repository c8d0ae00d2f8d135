#!/usr/bin/env python
"""Phase breakdown of the end-to-end path (reset -> odom_append -> lc_append -> find_inliers) on the headline workload.
RPGO_TRACE=1 makes the library print its own phase times to stderr."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
    p = pkg.PcmGpu(3, 0, odom_threshold=-1.0, lc_threshold=5.0)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p.reset()
        t1 = time.perf_counter()
        p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
        t2 = time.perf_counter()
        p.sync()
        t2b = time.perf_counter()
        p.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
        t3 = time.perf_counter()
        k, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU)
        p.sync()
        t4 = time.perf_counter()
        print("iter %d: reset %.2f  odom_append %.2f (+sync %.2f)  lc_append %.2f  find_inliers %.2f  total %.2f ms" %
              (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2b - t2) * 1e3, (t3 - t2b) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
    p.close()


if __name__ == "__main__":
    main()
