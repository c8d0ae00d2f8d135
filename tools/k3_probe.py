#!/usr/bin/env python
"""Kernel-only timing of the pairwise kernel for one (dim, mode, n): prints one JSON line.
Used for A/B measurements of build variants (RPGO_LIB_PATH=kimera-rpgo_b200/variants/librpgo_b200_<tag>.so)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
FLOP = {(3, 0): 6.2e3, (2, 0): 9.6e2, (3, 1): 7.4e2, (2, 1): 7.4e2}


def main():
    d, mode, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n) if d == 3 else synth.config3(seed=2, P=n, n=n))
    if mode == 1:
        params = dict(odom_trans=-1, odom_rot=-1, dist_trans=0.5 if d == 3 else 0.3, dist_rot=0.1 if d == 3 else 0.05)
    else:
        params = dict(odom_threshold=-1.0, lc_threshold=5.0 if d == 3 else 3.0)
    g = pkg.PcmGpu(d, mode, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    g.sync()
    st = torch.cuda.ExternalStream(g.stream_ptr())
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            g.pairwise_only(0, 0)
            e1.record(st)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    pairs = n * (n - 1) // 2
    sig = int(g.degrees(0).astype(np.int64).sum())
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("RPGO_LIB_PATH", "product")), d=d, mode=mode, n=n, k3_ms=round(best, 3),
                          pairs_per_s=pairs / (best * 1e-3), tflops=FLOP[(d, mode)] * pairs / (best * 1e-3) / 1e12, degree_sum=sig)))
    g.close()


if __name__ == "__main__":
    main()
