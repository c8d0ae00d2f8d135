"""Sampled-pair parity against the CPU oracle for matrices too large to build on one core.

The oracle's per-pair entry point (oracle/pcm_oracle.cpp::orc_check_pairs = Pcm::areLoopsConsistent, Pcm.h:670-718)
is run on a few host processes; each worker folds the odometry itself (the strict left fold of Pcm.h:516-557) and then
evaluates its slice of the sampled pairs.  TEST INFRASTRUCTURE: imports the oracle, never imported by the product."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(job):
    d, mode, params, arr, pi, pj = job
    sys.path.insert(0, HERE)
    import orc
    o = orc.OraclePcm(d, mode, **params)
    o.update_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["v_keys"], arr["v_pose"])
    return o.check_pairs(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"], pi, pj)


def oracle_pairs(d, mode, params, arr, pi, pj, procs=None):
    """(ok, dist, in_band) of the oracle for closure pairs (pi[t] older, pj[t] newer) of the table in `arr`."""
    pi = np.ascontiguousarray(pi, dtype=np.int32)
    pj = np.ascontiguousarray(pj, dtype=np.int32)
    assert (pi < pj).all()
    m = len(pi)
    procs = procs or max(1, min(16, (os.cpu_count() or 2) - 1, (m + 49999) // 50000))
    small = {k: arr[k] for k in ("o_prev", "o_new", "o_pose", "o_cov", "v_keys", "v_pose", "l_from", "l_to", "l_pose", "l_cov")}
    if procs == 1:
        return _worker((d, mode, params, small, pi, pj))
    cuts = np.linspace(0, m, procs + 1).astype(int)
    jobs = [(d, mode, params, small, pi[a:b], pj[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_worker, jobs)
    return tuple(np.concatenate([r[k] for r in res]) for k in range(3))


def sample_pairs(rng, n, m):
    """m distinct-ish random pairs i < j < n (uniform over the upper triangle)"""
    i = rng.integers(0, n, size=m)
    j = rng.integers(0, n, size=m)
    keep = i != j
    i, j = i[keep], j[keep]
    return np.minimum(i, j).astype(np.int32), np.maximum(i, j).astype(np.int32)


def bits_at(rows, pi, pj):
    """bit (pi, pj) of packed little-endian uint64 adjacency rows"""
    w = rows[pi, pj >> 6]
    return ((w >> (pj & 63).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)


def ref_heu_worker(adj):
    """(size, ids) of the reference's own FMC heuristic (oracle/_ref, else the oracle's restatement) on a dense 0/1 adjacency;
    a module-level function of this light module so that spawned worker processes import nothing else"""
    sys.path.insert(0, HERE)
    import orc
    f = orc.ref_clique_heu if orc.ref_fmc() is not None else orc.clique_heu
    k, ids = f(adj)
    return int(k), [int(x) for x in ids]
