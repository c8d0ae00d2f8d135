#!/usr/bin/env python
"""Multi-GPU parity check (run under torch.distributed.run on a multi-GPU box):
the row-sharded + all-gathered adjacency and the inlier ids must equal the single-GPU result bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_common import PcmGpu, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, gph, params in [
        ("single-group 3D n=3000", synth.config2(seed=4, P=3000, n=3000), dict(odom_threshold=-1, lc_threshold=5.0)),
        ("3 robots mixed n=2500", synth.config4(seed=3, robots=3, P=800, n=2500, outlier_frac=0.3),
         dict(odom_threshold=20.0, lc_threshold=5.0)),
    ]:
        single = PcmGpu(3, 0, device=local, **params)
        shard = PcmGpu(3, 0, device=local, rank=rank, world=world, **params)
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
            if not (same and inl and deg):
                ok = False
                print("rank %d MISMATCH %s group %d bits=%s inliers=%s deg=%s" % (rank, name, gi, same, inl, deg))
        if rank == 0:
            print("%s: %d groups, %d closures, %d inliers: %s" % (name, len(single.groups()), single.num_lc(),
                                                                   single.num_inliers(), "OK" if ok else "FAIL"))
    # rank-partitioned clique searches (heuristic, incremental, exact) with the NCCL incumbent exchange
    pkg = sys.modules[PcmGpu.__module__.rsplit(".", 1)[0]]
    rng = np.random.default_rng(77)
    single = PcmGpu(3, 0, device=local)
    shard = PcmGpu(3, 0, device=local, rank=rank, world=world)
    assert shard._exchange_cb is not None
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
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL")
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
