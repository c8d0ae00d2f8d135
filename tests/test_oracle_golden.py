"""Pins the CPU oracle against every golden vector / known-answer the reference's own tests hold
for the PCM hot path.  CPU only."""
import numpy as np
import pytest

import orc
import scenarios
from orc import pose3, quat_R, Rz, sym

TOL = 1e-9  # gtsam::assert_equal default tolerance used by the reference tests


def test_pwc_inverse():
    """tests/testPoseWithCovariance.cpp:32-48"""
    a = (pose3(), np.eye(6), 1)
    p, c, r = orc.pwc_inverse(3, a)
    np.testing.assert_allclose(p, orc.pose_inverse(3, pose3()), atol=TOL)
    np.testing.assert_allclose(c, np.eye(6), atol=TOL)


B_COVAR = 0.1 * np.array([[2, 0, 0, 0, -1, 1], [0, 2, 0, 1, 0, -1], [0, 0, 2, -1, 1, 0],
                          [0, 1, -1, 4, -1, -1], [-1, 0, 1, -1, 4, -1], [1, -1, 0, -1, -1, 4.0]])
C_COVAR = np.array([[0.3, 0, 0, 0, -0.1, -0.1], [0, 0.3, 0, 0.1, 0, 0.1], [0, 0, 0.3, 0.1, -0.1, 0],
                    [0, 0.1, 0.1, 0.41, -0.1, 0.1], [-0.1, 0, -0.1, -0.1, 0.41, 0.1], [-0.1, 0.1, 0, 0.1, 0.1, 0.41]])
D_COVAR = np.array([[0.4, 0, 0, 0, 0.1, -0.1], [0, 0.4, 0, -0.1, 0, 0.2], [0, 0, 0.4, 0.1, -0.2, 0],
                    [0, -0.1, 0.1, 0.51, 0, 0], [0.1, 0, -0.2, 0, 0.61, -0.1], [-0.1, 0.2, 0, 0, -0.1, 0.61]])


def test_pwc_compose_golden():
    """tests/testPoseWithCovariance.cpp:51-167 — B_covar :68-77, C_covar :107-115, D_covar :141-149"""
    A = (pose3(None, (1, 1, 1)), 0.1 * np.eye(6), 1)
    AB = (pose3(None, (1, 1, 1)), 0.1 * np.eye(6), 1)
    B = orc.pwc_compose(3, A, AB)
    np.testing.assert_allclose(B[0], orc.pose_compose(3, A[0], AB[0]), atol=TOL)
    np.testing.assert_allclose(B[1], B_COVAR, atol=TOL)
    bc_cov = 0.1 * np.eye(6)
    bc_cov[3:, 3:] = 0.01 * np.eye(3)
    BC = (pose3(quat_R(0, 0, 0, 1)), bc_cov, 1)
    Cc = orc.pwc_compose(3, B, BC)
    np.testing.assert_allclose(Cc[1], C_COVAR, atol=TOL)
    CD = (pose3(quat_R(0, 0, 1, 0), (1, 0, 0)), 0.1 * np.eye(6), 1)
    D = orc.pwc_compose(3, Cc, CD)
    np.testing.assert_allclose(D[1], D_COVAR, atol=TOL)


def test_pwc_between_golden():
    """tests/testPoseWithCovariance.cpp:170-217 — between covariance 0.1*I :194-200"""
    A = (pose3(), C_COVAR, 1)
    Cc = (pose3(quat_R(0, 0, 1, 0), (1, 0, 0)), D_COVAR, 1)
    B = orc.pwc_between(3, A, Cc)
    np.testing.assert_allclose(B[0], pose3(quat_R(0, 0, 1, 0), (1, 0, 0)), atol=TOL)
    np.testing.assert_allclose(B[1], 0.1 * np.eye(6), atol=TOL)


def test_pwn_norms():
    """tests/testPoseWithNode.cpp:82-99"""
    assert orc.pwn_norms(3, pose3(), 1) == (0.0, 0.0)
    t, _ = orc.pwn_norms(3, pose3(quat_R(1, 0, 0, 0), (1, 0, 0)), 5)
    assert t == 0.2
    _, r = orc.pwn_norms(3, pose3(quat_R(0, 0, 1, 0)), 5)
    assert abs(r - 3.1415927 / 5) < 1e-6


def test_trajectory_between_equals_restitched_fold():
    """tests/testTrajectory.cpp:35-109: getBetween(2,99) on a 100-step trajectory == re-stitched fold."""
    step = pose3(quat_R(1, 0, 0, 0), (1, 1, 0))
    dposes = np.tile(step, (100, 1))
    dcovs = np.tile(1e-4 * np.eye(6), (100, 1, 1))
    cp, cc, _, _ = orc.traj_fold(3, 0, pose3(), dposes, dcovs)
    # reference stores poses[i] = fold after i+1 steps
    at = lambda i: (cp[i + 1], cc[i + 1], 1)
    btw = orc.pwc_between(3, at(2), at(99))
    rp, rc, _, _ = orc.traj_fold(3, 0, pose3(), dposes[:97], dcovs[:97])
    np.testing.assert_allclose(btw[0], rp[97], atol=TOL)
    np.testing.assert_allclose(btw[1], rc[97], atol=TOL)


def _run(name):
    d, mode, params, calls, expect = scenarios.ALL[name]()
    pcm = orc.OraclePcm(d, mode, **params)
    for i, (factors, values) in enumerate(calls):
        do_opt = pcm.update(factors, values)
        if i in expect:
            assert (pcm.nfg_size(), pcm.num_values()) == expect[i], (name, i)
    return pcm, do_opt


@pytest.mark.parametrize("name", sorted(scenarios.ALL))
def test_reference_scenarios(name):
    """Factor / value counts asserted by the reference tests (file:line in tests/scenarios.py)."""
    _run(name)


def test_known_answer_distances():
    """Known-answer values: Mahalanobis / average distances of the reference tests."""
    # testPcm.cpp:61-86 odometry check distances: recompute through the oracle primitives
    R90 = scenarios.R90
    dposes = np.tile(pose3(R90, (1, 0, 0)), (3, 1))
    dcovs = np.tile(0.1 * np.eye(6), (3, 1, 1))
    cp, cc, _, _ = orc.traj_fold(3, 0, pose3(), dposes, dcovs)
    for var, want in [(0.1, 0.293598), (0.05, 0.309098)]:
        odom = orc.pwc_between(3, (cp[3], cc[3], 1), (cp[0], cc[0], 1))
        lc_inv = orc.pwc_inverse(3, (pose3(Rz(1.51), (0.8, 0, 0)), var * np.eye(6), 1))
        res = orc.pwc_compose(3, odom, lc_inv)
        assert abs(orc.pwc_mahalanobis(3, res) - want) < 2e-6
    # testPcm.cpp:147-191 pairwise distances
    pcm, _ = _run("pcm_consistency_check")
    adj, dist = pcm.group_adj(0)
    want = {(0, 1): 0.1655, (0, 2): 0.3138, (1, 2): 0.3375, (0, 3): 0.5723, (1, 3): 0.6379, (2, 3): 0.4589}
    for (i, j), w in want.items():
        assert abs(dist[i, j] - w) < 6e-5, (i, j, dist[i, j])
    assert adj.tolist() == [[0, 1, 1, 0], [1, 0, 1, 0], [1, 1, 0, 1], [0, 0, 1, 0]]
    assert list(pcm.group_inlier_ids(0)) == list(pcm.group_factor_ids(0)[:3])
    # testPcmSimple.cpp:234-280: the reference's own near-threshold pair (0.05 + 1.1e-13)
    pcm, _ = _run("simple_consistency_trans_check")
    adj, dist = pcm.group_adj(0)
    assert 0.05 < dist[0, 3] < 0.05 + 1e-9
    assert adj[0, 3] == 0
    # testMultiRobot.cpp (Pcm3D 3.0/0.05) group {a,b}
    pcm, _ = _run("multi_robot_pcm")
    groups = pcm.groups()
    gi = [i for i, g in enumerate(groups) if (g[0], g[1]) == ('a', 'b')][0]
    adj, dist = pcm.group_adj(gi)
    want = {(0, 1): 0.0, (0, 2): 2.9059, (1, 2): 1.8423, (0, 3): 0.1953, (1, 3): 0.2827, (2, 3): 5.6552}
    for (i, j), w in want.items():
        assert abs(dist[i, j] - w) < 6e-5, (i, j, dist[i, j])


def test_config1_ordered_g2o():
    """BASELINE config 1 on the `ordered` fixture: 17 LCs, all consistent under
    Pcm3D(1,1) and PcmSimple3D(1,1) => 153 factors; PcmSimple3D(0.05, 0.01): only pair (2,4)
    inconsistent => heuristic size 16 with the scratch-buffer ids [0,1,3,3,4,...,15]."""
    values, edges = scenarios.g2o_fixture("ordered")
    for mode, params in [(0, dict(odom_threshold=1.0, lc_threshold=1.0)),
                         (1, dict(odom_trans=1.0, odom_rot=1.0, dist_trans=1.0, dist_rot=1.0))]:
        pcm = orc.OraclePcm(3, mode, **params)
        pcm.update(edges, values)
        assert pcm.nfg_size() == 153
        assert pcm.num_lc() == 17 and pcm.num_inliers() == 17
    pcm = orc.OraclePcm(3, 1, odom_trans=0.05, odom_rot=0.01, dist_trans=0.05, dist_rot=0.01)
    pcm.update(edges, values)
    adj, _ = pcm.group_adj(0)
    off = [(i, j) for i in range(17) for j in range(i + 1, 17) if not adj[i, j]]
    assert off == [(2, 4)]
    fids = list(pcm.group_factor_ids(0))
    inl = [fids.index(x) for x in pcm.group_inlier_ids(0)]
    assert inl == [0, 1, 3, 3] + list(range(4, 16))


def test_unordered_g2o_classification():
    """`unordered` fixture: same graph, shuffled lines.  Odometry arriving out of order is classified per
    Pcm.h:189-194 and folded in arrival order (the reference's documented assumption is incremental odometry)."""
    values, edges = scenarios.g2o_fixture("unordered")
    pcm = orc.OraclePcm(3, 0, odom_threshold=1.0, lc_threshold=1.0)
    pcm.update(edges, values)
    assert pcm.num_values() == 140
    assert lib_odom(pcm) == 136


def lib_odom(pcm):
    return orc.lib().orc_num_odom(pcm.h)
