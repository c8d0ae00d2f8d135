"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact for the adjacency bitset, the inlier ids and the trajectory cache; distances compared with ==."""
import os

import numpy as np
import pytest

import orc
import scenarios
from gpu_common import ROOT, PcmGpu, pkg, synth

pytestmark = pytest.mark.gpu


def run_both(d, mode, params, calls, expect=None, **gpu_kw):
    o = orc.OraclePcm(d, mode, **params)
    g = PcmGpu(d, mode, **params, **gpu_kw)
    for i, (factors, values) in enumerate(calls):
        ro = o.update(factors, values)
        rg = g.update(factors, values)
        assert ro == rg, ("do_optimize", i)
        assert o.nfg_size() == g.nfg_size(), ("nfg size", i, o.nfg_size(), g.nfg_size())
        assert o.num_values() == g.num_values()
        if expect and i in expect:
            assert (g.nfg_size(), g.num_values()) == expect[i]
        assert sorted(o.output_ids().tolist()) == sorted(g.output_ids().tolist()), ("output factor set", i)
    return o, g


def compare_landmarks(o, g):
    """landmark groups: same landmarks, adjacency, distances and inlier ids"""
    ol, gl = o.landmarks(), g.landmarks()
    assert ol == gl, (ol, gl)
    for l, (key, n, ni) in enumerate(ol):
        ao, do = o.landmark_adj(l)
        gi = g.landmark_group[key]
        ag, dg = g.group_adj(gi, with_dist=True)
        assert ao.shape[0] == n and np.array_equal(ao, ag), ("landmark adjacency", l)
        iu = np.triu_indices(n, 1)
        assert np.array_equal(do[iu], dg[iu], equal_nan=True), ("landmark distances", l)
        fo, io = o.landmark_ids(l)
        assert fo.tolist() == g.group_factor_ids(gi).tolist() and io.tolist() == g.group_inlier_ids(gi).tolist()


def compare_groups(o, g, check_dist=True):
    compare_landmarks(o, g)
    og, gg = o.groups(), g.groups()
    assert [(a, b, n) for a, b, n, _ in og] == [(a, b, n) for a, b, n, _ in gg]
    for gi in range(len(og)):
        ao, do = o.group_adj(gi)
        if ao.shape[0] != og[gi][2]:
            continue  # loop check disabled: matrix stays 1x1 in the reference
        ag, dg = g.group_adj(gi, with_dist=check_dist)
        assert np.array_equal(ao, ag), ("adjacency bitset", gi)
        if check_dist:
            iu = np.triu_indices(ao.shape[0], 1)
            assert np.array_equal(do[iu], dg[iu], equal_nan=True), ("distances", gi, np.abs(do[iu] - dg[iu]).max())
        assert o.group_inlier_ids(gi).tolist() == g.group_inlier_ids(gi).tolist(), ("inlier ids", gi)
    assert o.num_lc() == g.num_lc() and o.num_inliers() == g.num_inliers()


@pytest.mark.parametrize("name", sorted(scenarios.ALL))
def test_reference_scenarios_on_gpu(name):
    d, mode, params, calls, expect = scenarios.ALL[name]()
    o, g = run_both(d, mode, params, calls, expect)
    compare_groups(o, g)


def test_trajectory_cache_bit_exact():
    gph = synth.config2(seed=5, P=600, n=10)
    for mode in (0, 1):
        o = orc.OraclePcm(3, mode)
        g = PcmGpu(3, mode)
        o.update(gph["odom"], gph["values"])
        g.update(gph["odom"], gph["values"])
        for key, _ in gph["values"][::37] + gph["values"][-1:]:
            po, co, no, ro = o.traj_get(key)
            pg, cg, ng, rg = g.traj_get(key)
            assert np.array_equal(po, pg), key
            if mode == 0:
                assert np.array_equal(co, cg), key
            assert (no, ro) == (ng, rg)


def _flagged_factor_pairs(g, gi):
    n, pairs = g.flagged(gi)
    ids = g.group_factor_ids(gi)
    return n, sorted((int(ids[a]), int(ids[b])) for a, b in pairs)


@pytest.mark.parametrize("mode,params", [
    (0, dict(odom_threshold=-1, lc_threshold=3.0)),
    (0, dict(odom_threshold=12.0, lc_threshold=5.0)),
    (1, dict(odom_trans=0.5, odom_rot=0.05, dist_trans=0.05, dist_rot=0.01)),
])
def test_config2_synthetic_3d(mode, params):
    gph = synth.config2(seed=1, P=2500, n=400)
    calls = [(gph["odom"], gph["values"]), (gph["lcs"], [])]
    o, g = run_both(3, mode, params, calls)
    compare_groups(o, g)
    if mode == 0:
        n, pairs = _flagged_factor_pairs(g, 0)
        assert sorted(map(tuple, o.flagged().tolist())) == pairs


def test_incremental_append_equals_batch():
    """adjacency grown one closure / small batches at a time == one batch (Pcm.h:725-768 incremental growth)."""
    gph = synth.config2(seed=2, P=800, n=150)
    params = dict(odom_threshold=-1, lc_threshold=4.0)
    a = PcmGpu(3, 0, **params)
    a.update(gph["odom"], gph["values"])
    a.update(gph["lcs"], [])
    b = PcmGpu(3, 0, **params, incremental=True)
    o = orc.OraclePcm(3, 0, **params, incremental=True)
    b.update(gph["odom"], gph["values"])
    o.update(gph["odom"], gph["values"])
    pos = 0
    for step in [1, 1, 2, 5, 31, 33, 64, 13]:
        chunk = gph["lcs"][pos:pos + step]
        pos += step
        b.update(chunk, [])
        o.update(chunk, [])
        assert o.group_inlier_ids(0).tolist() == b.group_inlier_ids(0).tolist()
    assert pos == 150
    assert np.array_equal(a.group_bits(0), b.group_bits(0))
    assert np.array_equal(o.group_adj(0)[0], b.group_adj(0, with_dist=False)[0])


def test_remove_last_and_ignore_prefix():
    d, mode, params, calls, _ = scenarios.multi_robot(False)
    o, g = run_both(d, mode, params, calls)
    for c1, c2 in [('a', 'b'), ('a', 'a'), ('a', 'b'), (None, None)]:
        ro = o.remove_last(c1, c2)
        rg = g.remove_last(c1, c2)
        assert ro == rg
        assert sorted(o.output_ids().tolist()) == sorted(g.output_ids().tolist())
    o.ignore_prefix('b'); g.ignore_prefix('b')
    assert sorted(o.output_ids().tolist()) == sorted(g.output_ids().tolist())
    o.revive_prefix('b'); g.revive_prefix('b')
    assert sorted(o.output_ids().tolist()) == sorted(g.output_ids().tolist())


def test_multi_robot_mixed_direction_groups():
    gph = synth.config4(seed=3, robots=3, P=300, n=240, outlier_frac=0.3)
    params = dict(odom_threshold=20.0, lc_threshold=5.0)
    calls = [(gph["odom"], gph["values"]), (gph["lcs"][:100], []), (gph["lcs"][100:], [])]
    o, g = run_both(3, 0, params, calls)
    compare_groups(o, g)


def test_config3_synthetic_2d():
    gph = synth.config3(seed=2, P=1500, n=300)
    for mode, params in [(0, dict(odom_threshold=-1, lc_threshold=3.0)),
                         (1, dict(odom_trans=-1, odom_rot=-1, dist_trans=0.05, dist_rot=0.05))]:
        calls = [(gph["odom"], gph["values"]), (gph["lcs"], [])]
        o, g = run_both(2, mode, params, calls)
        compare_groups(o, g)


def test_nan_rotation_covariance_path():
    """rotation_info = false (GeometryUtils.h:98-113, :175-183)."""
    gph = synth.config2(seed=7, P=300, n=40, outlier_frac=0.3)
    lcs = []
    for q, f in enumerate(gph["lcs"]):
        cov = np.array(f[4], dtype=np.float64).copy()
        if q % 3 == 0:
            cov[:3, :3] = np.nan
        lcs.append((f[0], f[1], f[2], f[3], cov))
    calls = [(gph["odom"], gph["values"]), (lcs, [])]
    o, g = run_both(3, 0, dict(odom_threshold=-1, lc_threshold=4.0), calls)
    compare_groups(o, g)


def test_scan_mode_flips_no_bit_outside_band():
    """The re-associated prefix scan (throughput path of K1) against the exact fold: trajectory agrees to
    ~1e-12 and the adjacency is identical except possibly for flagged near-threshold pairs."""
    gph = synth.config2(seed=11, P=3000, n=300)
    params = dict(odom_threshold=-1, lc_threshold=5.0)
    a = PcmGpu(3, 0, **params)
    b = PcmGpu(3, 0, **params, traj_mode=pkg.TRAJ_SCAN, scan_chunk=64)
    for x in (a, b):
        x.update(gph["odom"], gph["values"])
        x.update(gph["lcs"], [])
    key = gph["values"][-1][0]
    pa, ca, _, _ = a.traj_get(key)
    pb, cb, _, _ = b.traj_get(key)
    assert np.abs(pa - pb).max() < 1e-9 and np.abs(ca - cb).max() < 1e-9 * max(1.0, np.abs(ca).max())
    A, B = a.group_adj(0, with_dist=False)[0], b.group_adj(0, with_dist=False)[0]
    diff = np.argwhere(np.triu(A != B, 1))
    _, fl = a.flagged(0)
    flagged = set(map(tuple, fl.tolist()))
    assert all((int(i), int(j)) in flagged for i, j in diff)


def rand_graph(rng, n, p):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a + a.T


def test_clique_heuristic_matches_reference_fmc():
    """K4 against the reference's own FMC (oracle/_ref) when present, else the pinned restatement."""
    rng = np.random.default_rng(21)
    g = PcmGpu(3, 0)
    ref = orc.ref_clique_heu if orc.ref_fmc() is not None else orc.clique_heu
    for t in range(60):
        n = int(rng.integers(1, 200))
        a = rand_graph(rng, n, rng.uniform(0.05, 0.97))
        gi = g.load_adjacency(a)
        k, ids, true = g.find_inliers_raw(gi, pkg.CLIQUE_HEU)
        kr, ir = ref(a)
        assert k == kr and ids.tolist() == ir.tolist(), (t, n)
        s = true.tolist()
        assert len(set(s)) == k and all(a[x, y] for x in s for y in s if x != y)  # the greedy chain is a clique


def test_clique_incremental_matches_reference_fmc():
    rng = np.random.default_rng(22)
    g = PcmGpu(3, 0)
    ref = orc.ref_clique_heu_incremental if orc.ref_fmc() is not None else orc.clique_heu_incremental
    for t in range(60):
        n = int(rng.integers(3, 150))
        a = rand_graph(rng, n, rng.uniform(0.1, 0.95))
        num_new = int(rng.integers(1, n))
        prev = int(rng.integers(0, 8))
        gi = g.load_adjacency(a)
        k, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_HEU_INCREMENTAL, num_new, prev)
        kr, ir = ref(a, num_new, prev)
        assert k == kr and ids.tolist() == ir.tolist(), (t, n, num_new, prev)


def test_clique_large_planted():
    """large planted clique + noise: size-independent properties (the reference FMC is too slow here)."""
    rng = np.random.default_rng(23)
    n, k = 6000, 2500
    a = rand_graph(rng, n, 0.02)
    members = rng.choice(n, size=k, replace=False)
    a[np.ix_(members, members)] = 1
    np.fill_diagonal(a, 0)
    g = PcmGpu(3, 0)
    gi = g.load_adjacency(a)
    size, ids, true = g.find_inliers_raw(gi, pkg.CLIQUE_HEU)
    assert size >= k
    s = true.tolist()
    assert len(set(s)) == size and a[np.ix_(s, s)].sum() == size * (size - 1)
    deg = g.degrees(gi)
    assert np.array_equal(deg, a.sum(1))


@pytest.mark.parametrize("d", [3, 2])
def test_tiled_kernel_equals_direct_kernel(d):
    """the TMA-tiled K3 and the direct K3 must produce the same bitset, flagged pairs and inliers."""
    if d == 3:
        gph = synth.config4(seed=9, robots=2, P=700, n=900, outlier_frac=0.4)
        params = dict(odom_threshold=25.0, lc_threshold=5.0)
    else:
        gph = synth.config3(seed=6, P=2000, n=700)
        params = dict(odom_threshold=-1, lc_threshold=3.0)
    res = []
    for kern in (pkg.KERNEL_DIRECT, pkg.KERNEL_TILED):
        g = PcmGpu(d, 0, kernel=kern, **params)
        g.update(gph["odom"], gph["values"])
        half = len(gph["lcs"]) // 2
        g.update(gph["lcs"][:half], [])
        g.update(gph["lcs"][half:half + 7], [])
        g.update(gph["lcs"][half + 7:], [])
        res.append(g)
    a, b = res
    assert a.groups() == b.groups()
    for gi in range(len(a.groups())):
        assert np.array_equal(a.group_bits(gi), b.group_bits(gi)), gi
        na, pa = a.flagged(gi)
        nb, pb = b.flagged(gi)
        assert na == nb and sorted(map(tuple, pa.tolist())) == sorted(map(tuple, pb.tolist()))
        assert a.group_inlier_ids(gi).tolist() == b.group_inlier_ids(gi).tolist()


def test_closure_before_odometry_sees_later_trajectory():
    """A closure whose key has no trajectory entry yet uses the default entry (std::map::operator[],
    GraphUtils.h:42); pairs formed after the odometry arrives must use the real entry."""
    gph = synth.config2(seed=13, P=200, n=30, outlier_frac=0.2)
    params = dict(odom_threshold=-1, lc_threshold=6.0)
    o = orc.OraclePcm(3, 0, **params)
    g = PcmGpu(3, 0, **params)
    first, rest = gph["odom"][:100], gph["odom"][100:]
    for x in (o, g):
        x.update(first, gph["values"][:101])
        # values for all keys exist, odometry for the second half does not yet
        x.update([], gph["values"][101:])
        x.update(gph["lcs"][:15], [])
        x.update(rest, [])           # arrives late: classified as loop closures by Pcm.h:189-194 (no new value)
        x.update(gph["lcs"][15:], [])
    compare_groups(o, g, check_dist=False)


def test_tiled_equals_direct_at_scale():
    """2*10^8 pairs: the optimised kernel (shared-reciprocal divisions, convergent LLT probe, triangular
    forward solves) against the plain kernel that uses none of them — bitset equality."""
    n = 20000
    arr = synth.as_arrays(synth.config2(seed=4, P=n, n=n))
    bits = []
    for kern in (pkg.KERNEL_DIRECT, pkg.KERNEL_TILED):
        g = PcmGpu(3, 0, kernel=kern, odom_threshold=-1, lc_threshold=5.0)
        g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
        g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
        bits.append(g.group_bits(0))
        nfl = g.flagged(0)[0]
        g.close()
    assert np.array_equal(bits[0], bits[1])
    # symmetric, zero diagonal
    a = np.unpackbits(bits[1][:512].view(np.uint8), axis=1, bitorder="little")[:, :512]
    assert np.array_equal(a, a.T) and not a.diagonal().any()


def test_clique_exact_matches_reference_fmc():
    """K5 against the reference's own FMC::maxClique (size AND ids, i.e. the same tie-break)."""
    rng = np.random.default_rng(24)
    g = PcmGpu(3, 0)
    ref = orc.ref_clique_exact if orc.ref_fmc() is not None else orc.clique_exact
    for t in range(50):
        n = int(rng.integers(1, 90))
        a = rand_graph(rng, n, rng.uniform(0.05, 0.9))
        gi = g.load_adjacency(a)
        k, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_EXACT)
        kr, ir = ref(a)
        assert k == kr and ids.tolist() == ir.tolist(), (t, n, ids.tolist(), ir.tolist())


def test_clique_exact_planted():
    rng = np.random.default_rng(25)
    n, k = 1500, 40
    a = rand_graph(rng, n, 0.1)
    members = np.sort(rng.choice(n, size=k, replace=False))
    a[np.ix_(members, members)] = 1
    np.fill_diagonal(a, 0)
    g = PcmGpu(3, 0)
    gi = g.load_adjacency(a)
    size, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_EXACT)
    assert size == k and ids.tolist() == members.tolist()


@pytest.mark.parametrize("name", ["ordered", "unordered", "robot_a", "robot_b"])
def test_g2o_fixtures_on_gpu(name):
    """the reference's four g2o fixtures (tests/data/*.g2o) through both implementations, three configurations."""
    values, edges = scenarios.g2o_fixture(name)
    for mode, params in [(0, dict(odom_threshold=1.0, lc_threshold=1.0)),
                         (0, dict(odom_threshold=100.0, lc_threshold=100.0)),
                         (1, dict(odom_trans=0.05, odom_rot=0.01, dist_trans=0.05, dist_rot=0.01))]:
        o, g = run_both(3, mode, params, [(edges, values)])
        compare_groups(o, g)


def test_abi_error_behaviour():
    """status codes instead of exceptions / crashes (the reference logs a warning and continues)."""
    import ctypes as C
    g = PcmGpu(3, 0)
    lib, h = g.lib, g.h
    assert lib.rpgo_odom_append(h, 0, None, None, None, None, None) == 0          # empty batch is fine
    assert lib.rpgo_odom_append(h, 3, None, None, None, None, None) == 1          # RPGO_ERR_INVALID
    assert lib.rpgo_lc_append(h, 2, None, None, None, None, None, None, None, None) == 1
    size = C.c_int64()
    ids = (C.c_int32 * 4)()
    assert lib.rpgo_find_inliers(h, 7, 0, 0, 0, ids, C.byref(size), None) == 4    # RPGO_ERR_NOT_FOUND
    assert lib.rpgo_lc_remove_last(h, 0, None, None) == 4
    assert lib.rpgo_find_group(h, ord('a'), ord('b')) == -1
    assert g.remove_last() is None and g.remove_last('a', 'b') is None           # nothing to remove -> null edge
    assert g.update([], []) is False
    cfg = pkg._capi.RpgoCfg()
    lib.rpgo_default_cfg(C.byref(cfg))
    cfg.dim = 5
    hh = C.c_void_p()
    assert lib.rpgo_create(C.byref(cfg), C.byref(hh)) == 1


def test_single_closure_and_disabled_checks():
    """n = 1 groups (1x1 zero adjacency => clique size 1, id 0) and the 'threshold < 0 disables' rules (Pcm.h:74-82)."""
    gph = synth.config2(seed=17, P=120, n=12, outlier_frac=0.5)
    for params in [dict(odom_threshold=-1, lc_threshold=-1), dict(odom_threshold=5.0, lc_threshold=-1),
                   dict(odom_threshold=-1, lc_threshold=1e-6)]:
        calls = [(gph["odom"], gph["values"])] + [([lc], []) for lc in gph["lcs"]]
        o, g = run_both(3, 0, params, calls)
        assert o.num_inliers() == g.num_inliers()


def test_straight_line_operations_equal_ieee():
    """10^9 random operand pairs: the branch-free rcp / div / sqrt sequences == the built-in IEEE operations
    whenever they do not flag the operand for the exact path."""
    import ctypes as C
    lib = pkg._capi.load()
    m, c = C.c_uint64(), C.c_uint64()
    assert lib.rpgo_debug_check_fastmath(1_000_000_000, 12345, C.byref(m), C.byref(c)) == 0
    assert c.value > 2_000_000_000 and m.value == 0, (m.value, c.value)


def test_straight_line_kernel_equals_direct_kernel():
    """kernel variant 24 (pair_check_v2: straight-line + exact fallback) against the plain direct kernel."""
    n = 12000
    arr = synth.as_arrays(synth.config2(seed=8, P=n, n=n))
    bits = []
    for kern in (pkg.KERNEL_DIRECT, 24, 22):
        g = PcmGpu(3, 0, kernel=kern, odom_threshold=-1, lc_threshold=5.0)
        g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
        g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
        bits.append((g.group_bits(0), g.flagged(0)[0]))
        g.close()
    assert np.array_equal(bits[0][0], bits[1][0]) and bits[0][1] == bits[1][1]
    assert np.array_equal(bits[0][0], bits[2][0]) and bits[0][1] == bits[2][1]


def test_landmarks_two_robots_and_many_observations():
    """landmark observations from two robots (getBetween's different-prefix path, GraphUtils.h:43-57, as the
    reference runs it) and a landmark with > 32 observations (several adjacency words), Pcm3D and PcmSimple3D."""
    rng = np.random.default_rng(31)
    a = lambda i: orc.sym('a', i)
    b = lambda i: orc.sym('b', i)
    l = lambda i: orc.sym('l', i)
    I6 = np.eye(6)
    lcov = np.zeros((6, 6)); lcov[:3, :3] = np.nan; lcov[3:, 3:] = 0.04 * np.eye(3)
    full = 0.05 * I6
    calls = [([], [(a(0), orc.pose3()), (b(0), orc.pose3(None, (0, -2, 0)))])]
    P = 45
    for i in range(P):
        pa = orc.pose3(orc.Rz(0.05), (1, 0, 0)); pb = orc.pose3(orc.Rz(-0.03), (1, 0.1, 0))
        calls.append(([(orc.BETWEEN, a(i), a(i + 1), pa, 0.01 * I6), (orc.BETWEEN, b(i), b(i + 1), pb, 0.01 * I6)],
                      [(a(i + 1), pa), (b(i + 1), pb)]))
    # landmark 0: first seen by a, re-observed by a and b; landmark 1: 40 observations, half of them with full covariance
    calls.append(([(orc.BETWEEN, a(2), l(0), orc.pose3(None, (0, 1, 0)), lcov)], [(l(0), orc.pose3())]))
    calls.append(([(orc.BETWEEN, a(9), l(0), orc.pose3(None, (0.3, 1, 0)), lcov),
                   (orc.BETWEEN, b(4), l(0), orc.pose3(None, (0, 2, 0)), lcov),
                   (orc.BETWEEN, b(7), l(0), orc.pose3(None, (1, 2, 0)), full)], []))
    calls.append(([(orc.BETWEEN, a(1), l(1), orc.pose3(None, (5, 5, 0)), lcov)], [(l(1), orc.pose3())]))
    obs = []
    for k in range(40):
        t = rng.normal(size=3) * (0.2 if k % 3 else 4.0) + np.array([5.0 - k * 0.9, 5, 0])
        obs.append((orc.BETWEEN, a(2 + k), l(1), orc.pose3(orc.Rz(rng.normal() * 0.1), t), lcov if k % 2 else full))
    calls.append((obs[:7], []))
    calls.append((obs[7:], []))
    for mode, params in [(0, dict(odom_threshold=-1, lc_threshold=3.0)),
                         (1, dict(odom_trans=-1, odom_rot=-1, dist_trans=0.2, dist_rot=0.05))]:
        o, g = run_both(3, mode, dict(params, special_symbols=('l',)), calls)
        compare_groups(o, g)
        assert len(o.landmarks()) == 2 and o.landmarks()[1][1] == 41


def test_frame_alignment_front_half():
    """N4: T_w0_wi measurements of the inlier inter-robot closures and getRobotOdomValues, bit-exact vs oracle."""
    gph = synth.config4(seed=5, robots=3, P=400, n=500, outlier_frac=0.3)
    params = dict(odom_threshold=30.0, lc_threshold=5.0)
    o, g = run_both(3, 0, params, [(gph["odom"], gph["values"]), (gph["lcs"], [])])
    for ri in "bc":
        mo, mg = o.frame_align_measurements('a', ri), g.frame_align_measurements('a', ri)
        assert mo is not None and len(mo) > 0 and np.array_equal(mo, mg), ri
    T = orc.pose3(orc.Rz(0.3), (1.0, -2.0, 0.5))
    ko, po = o.robot_odom_values('b', T)
    kg, pg = g.robot_odom_values('b', T)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg) and len(ko) == 400
    assert g.frame_align_measurements('a', 'z') is None


def test_log_output_formats(tmp_path):
    """N2: the status files carry the reference's headers and one row per spin (OutlierRemoval.h:54-62,
    Pcm.h:1136-1164, RobustSolver.cpp:93-102)."""
    gph = synth.config2(seed=4, P=300, n=60)
    g = PcmGpu(d=3, mode=0, odom_threshold=30.0, lc_threshold=5.0)
    g.log_output(str(tmp_path))
    g.update(gph["odom"], gph["values"])
    g.update(gph["lcs"], [])
    st = open(tmp_path / "outlier_rejection_status.txt").read().splitlines()
    assert st[0] == "total inliers spin-time mc-time" and len(st) == 3
    tot, good = [int(v) for v in st[2].split()[:2]]
    assert tot == g.total_lc and good == g.total_good_lc and 0 < good <= tot
    cs = open(tmp_path / "rpgo_status.csv").read().splitlines()
    assert cs[0] == "graph-size,spin-time(mu-s),num-lc,num-inliers" and len(cs) == 3
    adj = np.loadtxt(tmp_path / "a-a_adj_matrix.txt")
    n = g.group_info(0)[2]
    assert adj.shape == (n, n) and np.array_equal(adj, adj.T)
    assert np.array_equal(adj, g.group_adj(0, with_dist=False)[0])
    g.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_clique_searches_match_single(world):
    """Candidates / roots partitioned over `world` ranks with the incumbent exchanged through rpgo_set_exchange:
    every rank must return exactly what one GPU returns (size, scratch-buffer ids, true clique; exact: same ids)."""
    from gpu_common import run_sharded
    rng = np.random.default_rng(31)
    cases = []
    for t in range(24):
        n = int(rng.integers(2, 160))
        p = rng.uniform(0.1, 0.95)
        cases.append((rand_graph(rng, n, p), int(rng.integers(1, n)), int(rng.integers(0, 6)), n <= 90 and p <= 0.8))
    single = PcmGpu(3, 0)
    want = []
    for a, num_new, prev, exact in cases:
        gi = single.load_adjacency(a)
        r = [single.find_inliers_raw(gi, pkg.CLIQUE_HEU), single.find_inliers_raw(gi, pkg.CLIQUE_HEU_INCREMENTAL, num_new, prev)]
        if exact:
            r.append(single.find_inliers_raw(gi, pkg.CLIQUE_EXACT))
        want.append(r)
    single.close()

    def work(h):
        res = []
        for a, num_new, prev, exact in cases:
            gi = h.load_adjacency(a)
            r = [h.find_inliers_raw(gi, pkg.CLIQUE_HEU), h.find_inliers_raw(gi, pkg.CLIQUE_HEU_INCREMENTAL, num_new, prev)]
            if exact:
                r.append(h.find_inliers_raw(gi, pkg.CLIQUE_EXACT))
            res.append(r)
        return res

    got, calls = run_sharded(world, lambda r: PcmGpu(3, 0, rank=r, world=world), work)
    assert calls > 0
    for r in range(world):
        for c, (w, g_) in enumerate(zip(want, got[r])):
            for m, (x, y) in enumerate(zip(w, g_)):
                assert x[0] == y[0] and x[1].tolist() == y[1].tolist(), (r, c, m)
                if m == 0:
                    assert x[2].tolist() == y[2].tolist(), (r, c)


def test_sharded_heuristic_on_pcm_graph():
    """the same on a real PCM adjacency (2000 closures): sharded == single, and fewer chains per rank"""
    from gpu_common import run_sharded
    gph = synth.config2(seed=9, P=2500, n=2000)
    one = PcmGpu(3, 0, odom_threshold=-1, lc_threshold=5.0)
    one.update(gph["odom"], gph["values"])
    one.update(gph["lcs"], [])
    adj, _ = one.group_adj(0, with_dist=False)
    k, ids, true = one.find_inliers_raw(0, pkg.CLIQUE_HEU)
    one.close()

    def work(h):
        gi = h.load_adjacency(adj)
        return h.find_inliers_raw(gi, pkg.CLIQUE_HEU)

    got, calls = run_sharded(4, lambda r: PcmGpu(3, 0, rank=r, world=4), work)
    for kk, ii, tt in got:
        assert kk == k and ii.tolist() == ids.tolist() and tt.tolist() == true.tolist()


def test_clique_exact_warp_kernel_midsize():
    """K5 warp kernel (n <= 1024: colouring bound + edge-task frontier) against the reference FMC on graphs where the
    frontier really splits roots over many warps; the block kernel (n > 1024) is covered by the planted case above."""
    rng = np.random.default_rng(26)
    ref = orc.ref_clique_exact if orc.ref_fmc() is not None else orc.clique_exact
    g = PcmGpu(3, 0)
    for n, p in [(400, 0.3), (1000, 0.12), (64, 0.85), (33, 1.0), (32, 0.0), (1024, 0.05)]:
        a = rand_graph(rng, n, p)
        gi = g.load_adjacency(a)
        k, ids, _ = g.find_inliers_raw(gi, pkg.CLIQUE_EXACT)
        kr, ir = ref(a)
        assert k == kr and ids.tolist() == ir.tolist(), (n, p, ids.tolist(), ir.tolist())


def _multi_group_handle(**kw):
    gph = synth.config4(seed=11, robots=4, P=600, n=3000, outlier_frac=0.3)
    g = PcmGpu(3, 0, odom_threshold=30.0, lc_threshold=5.0, **kw)
    arr = synth.as_arrays(gph)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    return g, arr


def test_find_inliers_batch_equals_per_group():
    """rpgo_find_inliers_batch (concurrent searches on worker streams) == rpgo_find_inliers group by group, for the
    heuristic, the incremental heuristic and the exact mode."""
    g, arr = _multi_group_handle()
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    groups = sorted(g.group_factors)
    assert len(groups) == 10
    single = [g.find_inliers_raw(gi, pkg.CLIQUE_HEU) for gi in groups]
    batch = g.find_inliers_batch(groups, pkg.CLIQUE_HEU)
    for (k, ids, _), (kb, ib) in zip(single, batch):
        assert k == kb and ids.tolist() == ib.tolist()
    nn = [max(1, len(g.group_factors[gi]) // 3) for gi in groups]
    pv = [3 + (gi % 4) for gi in groups]
    single = [g.find_inliers_raw(gi, pkg.CLIQUE_HEU_INCREMENTAL, a, b) for gi, a, b in zip(groups, nn, pv)]
    batch = g.find_inliers_batch(groups, pkg.CLIQUE_HEU_INCREMENTAL, nn, pv)
    for (k, ids, _), (kb, ib) in zip(single, batch):
        assert k == kb and ids.tolist() == ib.tolist()
    # exact mode on sparser copies of the adjacencies (keeps the search short)
    rng = np.random.default_rng(3)
    small = []
    for q in range(6):
        a = rand_graph(rng, int(rng.integers(20, 80)), rng.uniform(0.2, 0.7))
        small.append(g.load_adjacency(a, c1=chr(ord('m') + q), c2='z'))
    single = [g.find_inliers_raw(gi, pkg.CLIQUE_EXACT) for gi in small]
    batch = g.find_inliers_batch(small, pkg.CLIQUE_EXACT)
    for (k, ids, _), (kb, ib) in zip(single, batch):
        assert k == kb and ids.tolist() == ib.tolist()
    g.close()


def test_find_inliers_batch_spread_over_ranks():
    """with an exchange registered the batch assigns whole groups to ranks and combines with one all-reduce"""
    from gpu_common import run_sharded
    ref, arr = _multi_group_handle()
    ref.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    groups = sorted(ref.group_factors)
    want = ref.find_inliers_batch(groups, pkg.CLIQUE_HEU)
    adjs = [ref.group_adj(gi, with_dist=False)[0] for gi in groups]
    ref.close()

    def work(h):
        gl = [h.load_adjacency(a, c1=chr(ord('a') + q), c2='z') for q, a in enumerate(adjs)]
        return h.find_inliers_batch(gl, pkg.CLIQUE_HEU)

    got, calls = run_sharded(3, lambda r: PcmGpu(3, 0, rank=r, world=3), work)
    assert calls == 1
    for res in got:
        for (k, ids), (kb, ib) in zip(want, res):
            assert k == kb and ids.tolist() == ib.tolist()


def test_online_column_kernel_equals_batch():
    """a few closures appended one at a time to a large group (column-mode kernel: lanes over the older closures)
    give the same adjacency, degrees, flagged pairs and inliers as one batch append (tiled kernel)"""
    n0, extra = 5000, 5
    gph = synth.config2(seed=13, P=5200, n=n0 + extra)
    arr = synth.as_arrays(gph)
    params = dict(odom_threshold=-1.0, lc_threshold=5.0)
    a = PcmGpu(3, 0, **params)
    b = PcmGpu(3, 0, **params)
    for x in (a, b):
        x.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    a.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])
    b.lc_append_arrays(arr["l_from"][:n0], arr["l_to"][:n0], arr["l_pose"][:n0], arr["l_cov"][:n0])
    for k in range(n0, n0 + extra - 2):
        b.lc_append_arrays(arr["l_from"][k:k + 1], arr["l_to"][k:k + 1], arr["l_pose"][k:k + 1], arr["l_cov"][k:k + 1])
    k = n0 + extra - 2
    b.lc_append_arrays(arr["l_from"][k:], arr["l_to"][k:], arr["l_pose"][k:], arr["l_cov"][k:])   # two at once
    assert np.array_equal(a.group_bits(0), b.group_bits(0))
    assert np.array_equal(a.degrees(0), b.degrees(0))
    fa, fb = a.flagged(0), b.flagged(0)
    assert fa[0] == fb[0] and sorted(map(tuple, fa[1].tolist())) == sorted(map(tuple, fb[1].tolist()))
    ka, ia, _ = a.find_inliers_raw(0, pkg.CLIQUE_HEU)
    kb, ib, _ = b.find_inliers_raw(0, pkg.CLIQUE_HEU)
    assert ka == kb and ia.tolist() == ib.tolist()
    # recomputing the last columns in place must not change anything (set-or-clear stores)
    b.recompute(0, n0 + extra - 3)
    assert np.array_equal(a.group_bits(0), b.group_bits(0))
    a.close(); b.close()


def test_rpgo_read_g2o_cli(tmp_path):
    """config 1 end to end through files: the `ordered` fixture written as g2o, run through tools/rpgo_read_g2o.py
    (PcmSimple3D like examples/RpgoReadG2o.cpp), result.g2o + status logs read back.  Expected counts: SURVEY 8(d) config 1
    (17 closures all kept at 1.0/1.0 -> 153 factors; 0.05/0.01 -> 16 inliers, 152 factors)."""
    import importlib
    import importlib.util
    g2o = importlib.import_module("kimera-rpgo_b200.g2o")
    values, edges = scenarios.g2o_fixture("ordered")
    src = str(tmp_path / "ordered.g2o")
    g2o.write_g2o(src, values, [(a, b, p, c) for _, a, b, p, c in edges], d=3)
    spec = importlib.util.spec_from_file_location("rpgo_read_g2o", os.path.join(ROOT, "tools", "rpgo_read_g2o.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    for thr, n_fac, n_inl in [(("1.0", "1.0"), 153, 17), (("0.05", "0.01"), 152, 16)]:
        out = str(tmp_path / ("out_%s" % thr[0]))
        assert cli.main(["rpgo_read_g2o.py", "3d", src, thr[0], thr[1], out]) == 0
        v2, e2 = g2o.load3d(os.path.join(out, "result.g2o"))
        assert len(v2) == len(values) and len(e2) == n_fac
        st = open(os.path.join(out, "outlier_rejection_status.txt")).read().splitlines()
        assert st[0] == "total inliers spin-time mc-time" and [int(x) for x in st[1].split()[:2]] == [17, n_inl]
        assert os.path.exists(os.path.join(out, "rpgo_status.csv"))
