#!/usr/bin/env python
"""Multi-GPU parity check (run under torch.distributed.run on a multi-GPU box):
the row-sharded + all-gathered adjacency and the inlier ids must equal the single-GPU result bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import faulthandler
import importlib

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_common import PcmGpu, synth  # noqa: E402

par = importlib.import_module("kimera-rpgo_b200.parallel")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    faulthandler.enable()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    handles = []
    for name, gph, params in [
        ("single-group 3D n=3000", synth.config2(seed=4, P=3000, n=3000), dict(odom_threshold=-1, lc_threshold=5.0)),
        ("3 robots mixed n=2500", synth.config4(seed=3, robots=3, P=800, n=2500, outlier_frac=0.3),
         dict(odom_threshold=20.0, lc_threshold=5.0)),
    ]:
        single = PcmGpu(3, 0, device=local, **params)
        shard = par.attach_comm(PcmGpu(3, 0, device=local, rank=rank, world=world, **params))  # NCCL inside the library
        handles += [single, shard]
        half = len(gph["lcs"]) // 2
        for x in (single, shard):
            x.update(gph["odom"], gph["values"])
            x.update(gph["lcs"][:half], [])
            x.update(gph["lcs"][half:half + 5], [])       # incremental growth across ranks
            x.update(gph["lcs"][half + 5:], [])
        assert single.groups() == shard.groups()
        for gi in range(len(single.groups())):
            same = np.array_equal(single.group_bits(gi), shard.group_bits(gi))
            inl = single.group_inlier_ids(gi).tolist() == shard.group_inlier_ids(gi).tolist()
            deg = np.array_equal(single.degrees(gi), shard.degrees(gi))
            n1, p1 = single.flagged(gi, cap=1 << 16)
            n2, p2 = shard.flagged(gi, cap=1 << 16)   # collective: the union over the ranks' row chunks
            fl = n1 == n2 and sorted(map(tuple, p1.tolist())) == sorted(map(tuple, p2.tolist()))
            if not (same and inl and deg and fl):
                ok = False
                print("rank %d MISMATCH %s group %d bits=%s inliers=%s deg=%s flagged=%s" % (rank, name, gi, same, inl, deg, fl))
        if rank == 0:
            print("%s: %d groups, %d closures, %d inliers: %s" % (name, len(single.groups()), single.num_lc(),
                                                                   single.num_inliers(), "OK" if ok else "FAIL"))
    # rank-partitioned clique searches (heuristic, incremental, exact) with the NCCL incumbent exchange
    pkg = sys.modules[PcmGpu.__module__.rsplit(".", 1)[0]]
    rng = np.random.default_rng(77)
    single = PcmGpu(3, 0, device=local)
    shard = par.attach_comm(PcmGpu(3, 0, device=local, rank=rank, world=world))
    handles += [single, shard]
    assert shard.has_comm
    for t_ in range(12):
        n = int(rng.integers(2, 120))
        p = rng.uniform(0.1, 0.9)
        a = np.triu((rng.random((n, n)) < p).astype(np.uint8), 1)
        a = a + a.T
        modes = [(pkg.CLIQUE_HEU, 0, 0), (pkg.CLIQUE_HEU_INCREMENTAL, int(rng.integers(1, n)), int(rng.integers(0, 5)))]
        if n <= 80 and p <= 0.8:
            modes.append((pkg.CLIQUE_EXACT, 0, 0))
        g1, g2 = single.load_adjacency(a), shard.load_adjacency(a)
        for m, nn, pv in modes:
            k1, i1, _ = single.find_inliers_raw(g1, m, nn, pv)
            k2, i2, _ = shard.find_inliers_raw(g2, m, nn, pv)
            if k1 != k2 or i1.tolist() != i2.tolist():
                ok = False
                print("rank %d CLIQUE MISMATCH case %d mode %d: %d vs %d" % (rank, t_, m, k1, k2))
    if rank == 0:
        print("sharded clique searches: %s" % ("OK" if ok else "FAIL"))
    ok = big_sample(rank, world, local, handles) and ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res = int(t.item())
    del t
    for h in handles:   # library handles (and their communicators) go first, then the process group
        h.close()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if res == 1 else "FAIL")
    sys.exit(0 if res == 1 else 1)


def big_sample(rank, world, local, handles):
    """the headline workload shape (one group, n closures) sharded over the ranks: sampled pairs of the all-gathered
    adjacency against the CPU oracle (Pcm.h:670-718 restated), symmetry / degrees on every rank, same inliers everywhere"""
    import parity_tools as pt
    n = int(os.environ.get("MGPU_BIG_N", "50000"))
    m = int(os.environ.get("MGPU_BIG_SAMPLE", "1000000"))
    params = dict(odom_threshold=-1.0, lc_threshold=5.0)
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
    h = par.attach_comm(PcmGpu(3, 0, device=local, rank=rank, world=world, **params))
    handles.append(h)
    h.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    h.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    k, ids, true = h.find_inliers_raw(0)
    rows = h.group_bits(0)
    deg = h.degrees(0)
    nfl, fl = h.flagged(0, cap=1 << 16)
    ok = True
    # every rank holds the same full matrix and the same answer
    sig = torch.tensor([int(rows.sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF)), int(deg.sum()), k, int(ids.sum()), nfl],
                       dtype=torch.int64, device="cuda")
    lo, hi = sig.clone(), sig.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        ok = False
        print("rank %d: ranks disagree on the all-gathered state: %s vs %s" % (rank, lo.tolist(), hi.tolist()))
    if rank == 0:
        rng = np.random.default_rng(5)
        pi, pj = pt.sample_pairs(rng, n, m)
        # add the flagged pairs: they must lie inside the oracle's band too
        if nfl:
            fi, fj = np.minimum(fl[:, 0], fl[:, 1]), np.maximum(fl[:, 0], fl[:, 1])
            pi, pj = np.concatenate([pi, fi]), np.concatenate([pj, fj])
        want, _, band = pt.oracle_pairs(3, 0, params, arr, pi, pj)
        got = pt.bits_at(rows, pi, pj)
        sym = pt.bits_at(rows, pj, pi)
        bad = int((want != got).sum())
        flagged_set = set(map(tuple, np.sort(fl, axis=1).tolist()))
        missing = [(int(a), int(b)) for a, b, z in zip(pi, pj, band) if z and (int(a), int(b)) not in flagged_set]
        outside = int((band[len(pi) - nfl:] == 0).sum()) if nfl else 0
        pc = np.array([bin(int(x)).count("1") for x in rows[:64].reshape(-1)]).reshape(64, -1).sum(1)
        print("big sample n=%d world=%d: %d sampled pairs, %d mismatches vs oracle, %d asymmetric, %d band pairs not flagged, "
              "%d flagged pairs outside the band, clique %d" % (n, world, len(pi), bad, int((got != sym).sum()), len(missing), outside, k))
        ok = ok and bad == 0 and (got == sym).all() and not missing and outside == 0 and np.array_equal(pc, deg[:64])
    return ok


if __name__ == "__main__":
    main()
