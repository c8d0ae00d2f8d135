"""Round-2 GPU tests: epoch-based clique search against the reference's own FMC on graphs that force many epochs,
the tiled mirror pass at awkward sizes, handles on several devices in one process, rpgo_reset, and the parity checks at
the sizes BASELINE.json states (configs 2-5) against the CPU oracle — full matrices where the oracle finishes in seconds,
>= 1e6 sampled pairs (tests/parity_tools.py) elsewhere."""
import ctypes as C
import importlib
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import orc
import parity_tools as pt
from gpu_common import PcmGpu, pkg, synth

pytestmark = pytest.mark.gpu


def _graph_with_degree_ladder(rng, n, clique, p_noise):
    """planted clique + noise whose degrees are spread below AND between the clique sizes the greedy passes through: every
    improvement of the bound crosses degrees that are present, i.e. the search needs many epochs"""
    a = (rng.random((n, n)) < p_noise).astype(np.uint8)
    S = rng.choice(n, size=clique, replace=False)
    a[np.ix_(S, S)] = 1
    # a ladder of partial cliques of growing size: each one improves the bound a little
    for k in range(3, clique, max(1, clique // 12)):
        T = rng.choice(n, size=k, replace=False)
        a[np.ix_(T, T)] = 1
    a = np.triu(a, 1)
    return a + a.T


def test_clique_epochs_match_reference_fmc():
    rng = np.random.default_rng(42)
    g = PcmGpu(3, 0)
    ref = orc.ref_clique_heu if orc.ref_fmc() is not None else orc.clique_heu
    refi = orc.ref_clique_heu_incremental if orc.ref_fmc() is not None else orc.clique_heu_incremental
    seen_epochs = []
    cases = []
    for t in range(24):
        n = int(rng.integers(50, 900))
        cases.append(_graph_with_degree_ladder(rng, n, int(rng.integers(4, max(5, n // 6))), rng.uniform(0.005, 0.15)))
    for t in range(10):   # sparse random graphs: almost every vertex has a degree near the clique size
        n = int(rng.integers(200, 2500))
        a = (rng.random((n, n)) < rng.uniform(0.002, 0.03)).astype(np.uint8)
        a = np.triu(a, 1)
        cases.append(a + a.T)
    for t, a in enumerate(cases):
        n = a.shape[0]
        gi = g.load_adjacency(a)
        k, ids, true = g.find_inliers_raw(gi, pkg.CLIQUE_HEU)
        seen_epochs.append(g.clique_stats()["epochs"])
        kr, ir = ref(a)
        assert k == kr and ids.tolist() == ir.tolist(), (t, n, k, kr)
        s = true.tolist()
        assert len(set(s)) == k and all(a[x, y] for x in s for y in s if x != y)
        num_new, prev = int(rng.integers(1, n)), int(rng.integers(0, max(1, k)))
        k2, ids2, _ = g.find_inliers_raw(gi, pkg.CLIQUE_HEU_INCREMENTAL, num_new, prev)
        kr2, ir2 = refi(a, num_new, prev)
        assert k2 == kr2 and ids2.tolist() == ir2.tolist(), ("incremental", t, n, num_new, prev)
    assert max(seen_epochs) >= 3, seen_epochs   # the cases really exercise the epoch logic
    g.close()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 512, 513, 1000, 1537, 4100])
def test_mirror_and_degrees_at_awkward_sizes(n):
    """upper triangle from the pairwise kernel's layout -> mirror (tiled transpose) -> symmetric matrix, degrees = popcount;
    checked through rpgo_debug_load_group + rpgo_debug_pass on random upper triangles (lower part deliberately dirty)"""
    rng = np.random.default_rng(n)
    up = np.triu((rng.random((n, n)) < 0.37).astype(np.uint8), 1)
    dirty = up + np.tril((rng.random((n, n)) < 0.5).astype(np.uint8), -1)   # stale lower bits must be overwritten
    g = PcmGpu(3, 0)
    gi = g.load_adjacency(dirty)
    g.debug_pass(gi, 0)
    g.debug_pass(gi, 1)
    adj, _ = g.group_adj(gi, with_dist=False)
    assert np.array_equal(adj, up + up.T)
    assert np.array_equal(g.degrees(gi), (up + up.T).sum(1))
    g.close()


def test_incremental_mirror_keeps_old_rows():
    """growing a group closure by closure re-mirrors only the new tile columns: compare with the batch result"""
    gph = synth.config2(seed=23, P=1500, n=1100)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=-1.0, lc_threshold=5.0)
    a, b = PcmGpu(3, 0, **params), PcmGpu(3, 0, **params)
    for x in (a, b):
        x.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    a.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    cuts = [0, 500, 511, 513, 1024, 1030, 1100]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        b.lc_append_arrays(arr["l_from"][lo:hi], arr["l_to"][lo:hi], arr["l_pose"][lo:hi], arr["l_cov"][lo:hi])
    assert np.array_equal(a.group_bits(0), b.group_bits(0))
    assert np.array_equal(a.degrees(0), b.degrees(0))
    a.close(); b.close()


def test_reset_gives_a_fresh_solver():
    gph = synth.config4(seed=5, robots=3, P=500, n=600, outlier_frac=0.3)
    params = dict(odom_threshold=20.0, lc_threshold=5.0)
    fresh = PcmGpu(3, 0, **params)
    used = PcmGpu(3, 0, **params)
    other = synth.config2(seed=2, P=700, n=400)
    used.update(other["odom"], other["values"]); used.update(other["lcs"], [])
    used.reset()
    for x in (fresh, used):
        x.update(gph["odom"], gph["values"]); x.update(gph["lcs"], [])
    assert fresh.groups() == used.groups()
    for gi in range(len(fresh.groups())):
        assert np.array_equal(fresh.group_bits(gi), used.group_bits(gi))
        assert fresh.group_inlier_ids(gi).tolist() == used.group_inlier_ids(gi).tolist()
    assert fresh.output_ids().tolist() == used.output_ids().tolist()
    fresh.close(); used.close()


def test_handles_on_two_devices_in_one_process():
    """per-device function attributes, device guard in every entry point, no reliance on the caller's
    current device.  Runs config 2 on every visible device from ONE thread whose current device stays 0."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    gph = synth.config2(seed=1, P=2500, n=1000)
    params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    torch.cuda.set_device(0)
    hs = [PcmGpu(3, 0, device=d, **params) for d in range(min(ndev, 4))]
    for h in hs:
        h.update(gph["odom"], gph["values"])
    for h in reversed(hs):
        h.update(gph["lcs"], [])
    assert torch.cuda.current_device() == 0
    for h in hs[1:]:
        assert np.array_equal(hs[0].group_bits(0), h.group_bits(0))
        assert hs[0].group_inlier_ids(0).tolist() == h.group_inlier_ids(0).tolist()
    for h in hs:
        h.close()


# ---------------------------------------------------------------------------------------------------------------
# parity at the stated sizes (BASELINE.json configs), against the oracle
# ---------------------------------------------------------------------------------------------------------------
def test_config2_stated_size_full_oracle():
    """config 2: P = 2500 poses, n = 1000 closures, 50 % outliers: every pair, flagged list and inlier ids vs the oracle"""
    gph = synth.config2(seed=1, P=2500, n=1000)
    params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    g, o = PcmGpu(3, 0, **params), orc.OraclePcm(3, 0, **params)
    o.set_reference_shaped(False)
    for x in (g, o):
        x.update(gph["odom"], gph["values"]); x.update(gph["lcs"], [])
    ao, _ = o.group_adj(0)
    ag, _ = g.group_adj(0, with_dist=False)
    assert ao.shape == (1000, 1000) and np.array_equal(ao, ag)
    assert o.group_inlier_ids(0).tolist() == g.group_inlier_ids(0).tolist()
    nf, pf = g.flagged(0)
    fo = o.flagged()
    assert nf == len(fo)
    g.close()


def _sampled_parity(d, mode, params, arr, g, gi, n, m, seed):
    rows = g.group_bits(gi)
    nfl, fl = g.flagged(gi, cap=1 << 18)
    rng = np.random.default_rng(seed)
    pi, pj = pt.sample_pairs(rng, n, m)
    if nfl:
        pi = np.concatenate([pi, np.minimum(fl[:, 0], fl[:, 1])])
        pj = np.concatenate([pj, np.maximum(fl[:, 0], fl[:, 1])])
    want, _, band = pt.oracle_pairs(d, mode, params, arr, pi, pj)
    got = pt.bits_at(rows, pi, pj)
    assert int((want != got).sum()) == 0
    assert np.array_equal(got, pt.bits_at(rows, pj, pi))                    # symmetric
    flagged = set(map(tuple, np.sort(fl, axis=1).tolist()))
    assert all((int(a), int(b)) in flagged for a, b, z in zip(pi, pj, band) if z)   # oracle band pairs are flagged
    if nfl:
        assert band[-nfl:].all()                                             # every flagged pair is inside the oracle's band
    deg = g.degrees(gi)
    some = rng.integers(0, n, size=64)
    pc = np.array([int(np.unpackbits(rows[i].view(np.uint8)).sum()) for i in some])
    assert np.array_equal(pc, deg[some])
    return len(pi)


def test_config3_stated_size_sampled_oracle():
    """config 3: 2D Manhattan, P = n = 10 000, 30 % outliers, setPcm2DParams: >= 1e6 sampled pairs + the full flagged list"""
    n = 10000
    arr = synth.as_arrays(synth.config3(seed=2, P=n, n=n))
    params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    g = PcmGpu(2, 0, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    assert g.group_info(0)[2] == n
    assert _sampled_parity(2, 0, params, arr, g, 0, n, 1_050_000, 3) >= 1_000_000
    g.close()


def test_config5_50k_sampled_oracle():
    """the headline workload (one group of 50 000 closures): >= 1e6 sampled pairs, near-threshold list vs the oracle's band"""
    n = 50000
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
    params = dict(odom_threshold=-1.0, lc_threshold=5.0)
    g = PcmGpu(3, 0, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    assert _sampled_parity(3, 0, params, arr, g, 0, n, 1_050_000, 4) >= 1_000_000
    k, ids, true = g.find_inliers_raw(0)
    rows = g.group_bits(0)
    t = np.asarray(true)
    sub = pt.bits_at(rows, np.repeat(t, len(t)), np.tile(t, len(t))).reshape(len(t), len(t))
    assert k == len(t) and (sub + np.eye(len(t), dtype=np.uint8)).all()   # the greedy chain is a clique of the GPU's own adjacency
    g.close()


def test_config4_stated_size_groups_and_fmc():
    """config 4: 8 robots x 20 000 poses, 50 000 closures, 36 ObservationId groups (Pcm.h:472-486): three whole groups against
    the oracle, sampled pairs of the others, and the inlier ids of EVERY group against the reference's own FMC run on the
    GPU's adjacency (Pcm.h:857-876 hands each group to findMaxCliqueHeu)."""
    gph = synth.config4(seed=3, robots=8, P=20000, n=50000, outlier_frac=0.3)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=50.0, lc_threshold=5.0)
    g = PcmGpu(3, 0, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    num_new, acc = g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    grp, idx = g.last_group.copy(), g.last_index.copy()
    groups = g.groups()
    assert len(groups) == 36 and sum(x[2] for x in groups) == int(acc.sum())
    res = g.find_inliers_batch(list(range(36)), pkg.CLIQUE_HEU)
    # (1) odometry check decisions and three whole groups vs the oracle, fed with those groups' closures only
    whole = [0, 17, 35]
    sel = np.isin(grp, whole) | (acc == 0)
    o = orc.OraclePcm(3, 0, **params)
    o.set_reference_shaped(False)
    o.update_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["v_keys"], arr["v_pose"])
    o.update_arrays(arr["l_from"][sel], arr["l_to"][sel], arr["l_pose"][sel], arr["l_cov"][sel])
    og = {(a, b): k for k, (a, b, _, _) in enumerate(o.groups())}
    for gi in whole:
        a_, b_, n_, _ = groups[gi]
        ao, _ = o.group_adj(og[(a_, b_)])
        ag, _ = g.group_adj(gi, with_dist=False)
        assert ao.shape[0] == n_ and np.array_equal(ao, ag), gi   # same closures accepted (K2) and same matrix (K3)
    # (2) sampled pairs of all the other groups: build per-group closure tables in the GPU's order
    rng = np.random.default_rng(9)
    total = 0
    adjs = []
    o2 = orc.OraclePcm(3, 0, **params)  # one oracle, one fold of the 160 000 odometry steps, for all sampled groups
    o2.update_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["v_keys"], arr["v_pose"])
    for gi in range(36):
        members = np.flatnonzero(grp == gi)
        order = members[np.argsort(idx[members])]
        n_ = len(order)
        rows = g.group_bits(gi)
        if gi not in whole:
            pi, pj = pt.sample_pairs(rng, n_, 12000)
            want, _, _ = o2.check_pairs(arr["l_from"][order], arr["l_to"][order], arr["l_pose"][order], arr["l_cov"][order], pi, pj)
            assert np.array_equal(want, pt.bits_at(rows, pi, pj)), gi
            total += len(pi)
        adjs.append(np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :n_])
    assert total >= 33 * 11000
    # (3) inlier ids of every group == the reference's FMC on this adjacency (its nested linear scans need ~4 s per
    # 1400-closure group: the 36 searches run on a few host processes)
    with mp.get_context("spawn").Pool(min(12, max(1, (os.cpu_count() or 2) - 1))) as pool:
        refs = pool.map(pt.ref_heu_worker, adjs)
    for gi, (kr, ir) in enumerate(refs):
        assert res[gi][0] == kr and res[gi][1].tolist() == ir, gi
    g.close()


# ---------------------------------------------------------------------------------------------------------------
# tiled kernels of the other pair functions (PcmSimple3D / PcmSimple2D / Pcm2D)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [3, 2])
def test_simple_tiled_kernel_equals_direct_kernel(d):
    """PcmSimple through the TMA-tiled kernel (compact records, straight-line pair function + exact fallback) and through the
    plain direct kernel: same bitset, flagged pairs and inliers, incl. incremental growth and the online column kernel."""
    if d == 3:
        gph = synth.config4(seed=19, robots=2, P=700, n=1500, outlier_frac=0.4)
        params = dict(odom_trans=2.0, odom_rot=1.0, dist_trans=0.4, dist_rot=0.08)
    else:
        gph = synth.config3(seed=16, P=2000, n=1200)
        params = dict(odom_trans=-1, odom_rot=-1, dist_trans=0.3, dist_rot=0.05)
    res = []
    for kern in (pkg.KERNEL_DIRECT, pkg.KERNEL_TILED):
        g = PcmGpu(d, pkg.MODE_SIMPLE, kernel=kern, **params)
        g.update(gph["odom"], gph["values"])
        half = len(gph["lcs"]) // 2
        g.update(gph["lcs"][:half], [])
        g.update(gph["lcs"][half:half + 7], [])          # column-mode kernel on the larger groups
        g.update(gph["lcs"][half + 7:], [])
        res.append(g)
    a, b = res
    assert a.groups() == b.groups() and sum(x[2] for x in a.groups()) > 500
    dens = []
    for gi in range(len(a.groups())):
        ba, bb = a.group_bits(gi), b.group_bits(gi)
        assert np.array_equal(ba, bb), gi
        na, pa = a.flagged(gi)
        nb, pb = b.flagged(gi)
        assert na == nb and sorted(map(tuple, pa.tolist())) == sorted(map(tuple, pb.tolist()))
        assert a.group_inlier_ids(gi).tolist() == b.group_inlier_ids(gi).tolist()
        n_ = a.groups()[gi][2]
        dens.append(a.degrees(gi).sum() / max(1, n_ * (n_ - 1)))
    assert 0.01 < max(dens) < 0.99, dens   # the thresholds really separate consistent from inconsistent pairs
    a.close(); b.close()


@pytest.mark.parametrize("d,mode", [(3, 1), (2, 1), (2, 0)])
def test_other_modes_sampled_oracle_at_10k(d, mode):
    """PcmSimple3D / PcmSimple2D / Pcm2D at n = 10 000 closures: 3e5 sampled pairs against the CPU oracle"""
    n = 10000
    if d == 3:
        arr = synth.as_arrays(synth.config2(seed=6, P=n, n=n))
    else:
        arr = synth.as_arrays(synth.config3(seed=7, P=n, n=n))
    if mode == 1:
        params = dict(odom_trans=-1, odom_rot=-1, dist_trans=0.5 if d == 3 else 0.3, dist_rot=0.1 if d == 3 else 0.05)
    else:
        params = dict(odom_threshold=-1.0, lc_threshold=3.0)
    g = PcmGpu(d, mode, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    rows = g.group_bits(0)
    rng = np.random.default_rng(17)
    pi, pj = pt.sample_pairs(rng, n, 300_000)
    want, _, _ = pt.oracle_pairs(d, mode, params, arr, pi, pj)
    got = pt.bits_at(rows, pi, pj)
    assert int((want != got).sum()) == 0
    assert 0.001 < got.mean() < 0.999
    g.close()


@pytest.mark.parametrize("first", [1, 3])
def test_disabled_loop_check_incremental_and_remove_last(first):
    """lc_threshold < 0 with incremental mode / removeLastLoopClosure must not raise; behaviour = the reference's
    (FMC on the never-grown 1x1 matrix) where it is defined, findInliers' all-inliers rule where the reference is UB."""
    gph = synth.config2(seed=3, P=300, n=12)
    for inc in (True, False):
        params = dict(odom_threshold=-1, lc_threshold=-1, incremental=inc)
        g, o = PcmGpu(3, 0, **params), orc.OraclePcm(3, 0, **params)
        for x in (g, o):
            x.update(gph["odom"], gph["values"])
            x.update(gph["lcs"][:first], [])
            x.update(gph["lcs"][first:first + 1], [])
            x.update(gph["lcs"][first + 1:first + 4], [])
        assert g.num_inliers() == o.num_inliers() and g.group_inlier_ids(0).tolist() == o.group_inlier_ids(0).tolist()
        assert g.output_ids().tolist() == o.output_ids().tolist()
        assert g.remove_last() == o.remove_last()
        assert g.group_inlier_ids(0).tolist() == o.group_inlier_ids(0).tolist()
        assert g.output_ids().tolist() == o.output_ids().tolist()
        g.close()


def test_long_odometry_batch_is_sliced_and_still_bit_exact():
    """rpgo_odom_append feeds batches longer than 12 288 steps to the exact fold in slices of 8192 (the fold of one slice
    overlaps the host staging of the next): the trajectory table must equal the oracle's strict left fold bit for bit, in
    3D and 2D, including across the slice boundaries, and equal what many small appends produce."""
    for d, gph in ((3, synth.config2(seed=8, P=20000, n=10)), (2, synth.config3(seed=8, P=20000, n=10))):
        arr = synth.as_arrays(gph)
        o = orc.OraclePcm(d, 0)
        o.update_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["v_keys"], arr["v_pose"])
        g = PcmGpu(d, 0)
        g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
        g2 = PcmGpu(d, 0)
        for a in range(0, len(arr["o_prev"]), 3001):
            sl = slice(a, a + 3001)
            g2.odom_append_arrays(arr["o_prev"][sl], arr["o_new"][sl], arr["o_pose"][sl], arr["o_cov"][sl], arr["o_init"][sl])
        keys = [int(k) for k in arr["o_new"]]
        probe = keys[::397] + keys[8185:8200] + keys[16377:16392] + keys[-3:]
        for key in probe:
            po, co, no, ro = o.traj_get(key)
            for h in (g, g2):
                pg, cg, ng, rg = h.traj_get(key)
                assert np.array_equal(po, pg) and np.array_equal(co, cg) and (no, ro) == (ng, rg), (d, key)
        g.close()
        g2.close()
