#!/usr/bin/env python
"""K5 probe (run under gpurun): exact clique on dense random graphs, warp/colouring kernel vs RPGO_EXACT_BLOCK=1."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("kimera-rpgo_b200")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 31)
g = pkg.PcmGpu(3, 0)
for n, p in [(78, 0.95), (90, 0.9), (200, 0.7), (500, 0.5), (1000, 0.3), (1000, 0.5)]:
    a = np.triu((rng.random((n, n)) < p).astype(np.uint8), 1); a = a + a.T
    gi = g.load_adjacency(a)
    t0 = time.perf_counter(); k, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_EXACT); t1 = time.perf_counter()
    ok = all(a[x, y] for x in ids for y in ids if x != y)
    print("n=%d p=%.2f exact size %d in %.1f ms clique_ok=%s ids_hash=%d" % (n, p, k, (t1 - t0) * 1e3, ok, hash(tuple(ids.tolist())) % 100000), flush=True)
