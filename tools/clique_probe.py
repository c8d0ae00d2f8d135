#!/usr/bin/env python
"""The bitset passes and the heuristic clique search on the headline workload shape (one group of n closures): timings,
epochs and algorithmic bytes; also the target of the ncu captures of mirror / degree / heu kernels."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
    g = pkg.PcmGpu(3, 0, odom_threshold=-1.0, lc_threshold=5.0)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    st = torch.cuda.ExternalStream(g.stream_ptr())
    out = dict(n=n)

    def timed(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                fn()
                e1.record(st)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    out["mirror_ms"] = timed(lambda: g.debug_pass(0, 0))
    out["degree_ms"] = timed(lambda: g.debug_pass(0, 1))
    out["clique_ms"] = timed(lambda: g.find_inliers_raw(0, pkg.CLIQUE_HEU))
    k, ids, true = g.find_inliers_raw(0, pkg.CLIQUE_HEU)
    out.update(clique_size=k, **g.clique_stats())
    deg = g.degrees(0)
    out.update(deg_min=int(deg.min()), deg_max=int(deg.max()), deg_mean=float(deg.mean()))
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
