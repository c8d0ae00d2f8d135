#!/usr/bin/env python
"""K1 probe: odom_append (+ sync) wall time for 2D / 3D chains and a checksum of sampled trajectory entries, so that kernel
variants (RPGO_LIB_PATH, RPGO_FOLD_V2=1 = the two-warp batched kernel) can be compared
bit for bit.  Run under gpurun."""
import hashlib, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("kimera-rpgo_b200"); synth = importlib.import_module("kimera-rpgo_b200.synth")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
DIMS = sys.argv[2] if len(sys.argv) > 2 else "23"
for d in [int(c) for c in DIMS]:
    gph = synth.config3(seed=2, P=P, n=100) if d == 2 else synth.config2(seed=1, P=P, n=100)
    arr = synth.as_arrays(gph)
    best = 1e9
    for it in range(4):
        p = pkg.PcmGpu(d, 0, odom_threshold=-1.0, lc_threshold=5.0)
        p.sync(); t0 = time.perf_counter()
        p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"]); t1 = time.perf_counter(); p.sync()
        t2 = time.perf_counter()
        best = min(best, t2 - t0)
        if it == 3:
            hsh = hashlib.sha1()
            vals = gph["values"]
            for key, _ in vals[::max(1, len(vals) // 997)] + vals[-1:]:
                pg, cg, ng, rg = p.traj_get(key)
                hsh.update(pg.tobytes()); hsh.update(cg.tobytes()); hsh.update(bytes([int(rg)]))
            print("d=%d steps=%d odom_append host %.2f ms + sync %.2f ms; best total %.2f ms; sha1 %s" %
                  (d, len(arr["o_prev"]), (t1 - t0) * 1e3, (t2 - t1) * 1e3, best * 1e3, hsh.hexdigest()[:16]), flush=True)
        p.close()
