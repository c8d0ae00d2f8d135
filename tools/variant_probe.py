#!/usr/bin/env python
"""K3 variant probe (run under gpurun): time kernel selectors on one synthetic group and check that every variant
produces the same adjacency bitset as the first one.  usage: variant_probe.py 24,25,26 8000,20000"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
kernels = [int(x) for x in sys.argv[1].split(",")]
sizes = [int(x) for x in sys.argv[2].split(",")]
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
print("stagger env", os.environ.get("RPGO_STAGGER_NS"))
for n in sizes:
    gph = synth.config2(seed=4, P=max(2500, n), n=n) if dim == 3 else synth.config3(seed=4, P=max(2500, n), n=n)
    arr = synth.as_arrays(gph)
    ref = None
    for kern in kernels:
        g = pkg.PcmGpu(dim, 0, odom_threshold=-1, lc_threshold=5.0, kernel=kern, traj_mode=pkg.TRAJ_SCAN)
        g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
        g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
        st = torch.cuda.ExternalStream(g.stream_ptr())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        ms = []
        for rep in range(4):
            with torch.cuda.stream(st):
                e0.record(st)
                g.pairwise_only(0, 0)
                e1.record(st)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        g.finalize(0)
        bits = g.group_bits(0)
        same = True if ref is None else bool(np.array_equal(ref, bits))
        if ref is None:
            ref = bits
        pairs = n * (n - 1) / 2
        best = min(ms[1:])
        print("n=%d kernel=%d ms %s | %.3e pairs/s | %.2f TFLOP/s | bits_equal=%s"
              % (n, kern, ["%.2f" % m for m in ms], pairs / (best * 1e-3), pairs * 6.2e3 / (best * 1e-3) / 1e12, same), flush=True)
        g.close()
