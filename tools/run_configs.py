#!/usr/bin/env python
"""Run the five BASELINE.json configurations on one GPU (under gpurun) and print one JSON line each.
Parity: configs 1-3 are compared with the CPU oracle (bit-exact adjacency + inlier ids) at the sizes the oracle
finishes in seconds; configs 4-5 are checked through size-independent properties (tiled kernel == direct
kernel bitsets on a slice, symmetry, zero diagonal, degrees == popcount)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import orc  # noqa: E402  (checker only)
import scenarios  # noqa: E402

pkg = importlib.import_module("kimera-rpgo_b200")
synth = importlib.import_module("kimera-rpgo_b200.synth")
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "3", "4", "5"]


def timed_update(g, arr):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    t1 = time.perf_counter()
    num_new, acc = g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    g.sync()
    t2 = time.perf_counter()
    sizes = {}
    for gi in sorted(num_new):
        k, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_HEU)
        sizes[gi] = k
    g.sync()
    t3 = time.perf_counter()
    batch = g.find_inliers_batch(sorted(num_new), pkg.CLIQUE_HEU)   # the same searches, concurrently
    g.sync()
    t4 = time.perf_counter()
    assert [b[0] for b in batch] == [sizes[gi] for gi in sorted(num_new)]
    return dict(odom_ms=(t1 - t0) * 1e3, lc_append_ms=(t2 - t1) * 1e3, clique_ms=(t3 - t2) * 1e3,
                clique_batch_ms=(t4 - t3) * 1e3, accepted=int(acc.sum()),
                inliers=int(sum(sizes.values())), groups=len(sizes))


def kernel_ms(g, gi):
    st = torch.cuda.ExternalStream(g.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        with torch.cuda.stream(st):
            e0.record(st)
            g.pairwise_only(gi, 0)
            e1.record(st)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    g.finalize(gi)
    return best


def props(g, gi, rows=256):
    b = g.group_bits(gi)
    n = b.shape[0]
    a = np.unpackbits(b[:rows].view(np.uint8), axis=1, bitorder="little")[:, :n]
    sub = a[:, :rows]
    deg = g.degrees(gi)
    return bool(np.array_equal(sub, sub.T) and not sub.diagonal().any() and np.array_equal(deg[:rows], a.sum(1)))


out = []
if "1" in which:
    values, edges = scenarios.g2o_fixture("ordered")
    for mode, params, name in [(0, dict(odom_threshold=1.0, lc_threshold=1.0), "Pcm3D(1,1)"),
                               (1, dict(odom_trans=1.0, odom_rot=1.0, dist_trans=1.0, dist_rot=1.0), "PcmSimple3D(1,1)"),
                               (1, dict(odom_trans=0.05, odom_rot=0.01, dist_trans=0.05, dist_rot=0.01), "PcmSimple3D(0.05,0.01)")]:
        g = pkg.PcmGpu(3, mode, **params)
        o = orc.OraclePcm(3, mode, **params)
        g.update(edges, values)
        o.update(edges, values)
        fids = list(g.group_factor_ids(0))
        out.append(dict(config=1, what="ordered g2o fixture, " + name, factors=g.nfg_size(), num_lc=g.num_lc(), inliers=g.num_inliers(),
                        inlier_ids=[fids.index(x) for x in g.group_inlier_ids(0)],
                        equals_oracle=bool(np.array_equal(g.group_adj(0, False)[0], o.group_adj(0)[0]) and
                                           g.group_inlier_ids(0).tolist() == o.group_inlier_ids(0).tolist())))
if "2" in which:
    gph = synth.config2(seed=1, P=2500, n=1000)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    g = pkg.PcmGpu(3, 0, **params)
    r = timed_update(g, arr)
    o = orc.OraclePcm(3, 0, **params)
    t0 = time.perf_counter()
    o.update(gph["odom"], gph["values"]); o.update(gph["lcs"], [])
    r.update(config=2, what="3D sphere/helix P=2500 n=1000 50% outliers Pcm3D(-1,3)", pairs=499500, k3_ms=kernel_ms(g, 0),
             oracle_s=time.perf_counter() - t0,
             equals_oracle=bool(np.array_equal(g.group_adj(0, False)[0], o.group_adj(0)[0]) and
                                g.find_inliers_raw(0)[1].tolist() == [list(o.group_factor_ids(0)).index(x) for x in o.group_inlier_ids(0)]),
             flagged=g.flagged(0)[0], flagged_oracle=len(o.flagged()))
    out.append(r)
if "3" in which:
    gph = synth.config3(seed=2, P=10000, n=10000)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    g = pkg.PcmGpu(2, 0, **params)
    r = timed_update(g, arr)
    # oracle on the first 1500 closures (the adjacency of a prefix is a prefix of the adjacency)
    o = orc.OraclePcm(2, 0, **params)
    o.set_reference_shaped(False)
    o.update(gph["odom"], gph["values"]); o.update(gph["lcs"][:1500], [])
    ao = o.group_adj(0)[0]
    ag = np.unpackbits(g.group_bits(0)[:1500].view(np.uint8), axis=1, bitorder="little")[:, :1500]
    r.update(config=3, what="2D Manhattan P=10000 n=10000 30% outliers Pcm2D(-1,3)", pairs=10000 * 9999 // 2, k3_ms=kernel_ms(g, 0),
             prefix1500_equals_oracle=bool(np.array_equal(ao, ag)), props_ok=props(g, 0))
    out.append(r)
if "4" in which:
    gph = synth.config4(seed=3, robots=8, P=20000, n=50000, outlier_frac=0.3)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=50.0, lc_threshold=5.0)
    g = pkg.PcmGpu(3, 0, **params)
    r = timed_update(g, arr)
    sizes = [x[2] for x in g.groups()]
    pairs = int(sum(s * (s - 1) // 2 for s in sizes))
    r.update(config=4, what="8 robots x 20000 poses, 50000 closures over 36 prefix pairs, mixed directions, Pcm3D(50,5)",
             group_sizes=sizes, pairs=pairs, props_ok=all(props(g, gi) for gi in range(0, len(sizes), 7)))
    out.append(r)
if "5" in which:
    n = int(os.environ.get("CONFIG5_N", "200000"))
    gph = synth.config2(seed=4, P=n, n=n)
    arr = synth.as_arrays(gph)
    del gph
    params = dict(odom_threshold=-1.0, lc_threshold=5.0)
    g = pkg.PcmGpu(3, 0, **params)
    r = timed_update(g, arr)
    r.update(config=5, what="3D single group n=%d (%.2e pairs) Pcm3D(-1,5)" % (n, n * (n - 1) / 2), pairs=n * (n - 1) // 2,
             k3_ms=kernel_ms(g, 0), props_ok=props(g, 0))
    r["pairs_per_s_k3"] = r["pairs"] / (r["k3_ms"] * 1e-3)
    out.append(r)
for r in out:
    print(json.dumps(r))
