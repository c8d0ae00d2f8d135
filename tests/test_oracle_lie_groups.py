"""Pins the oracle's Pose2 / Pose3 formulas to the Lie-group definitions themselves (CPU only).

The reference's own tests hold golden vectors for Pose3 only (tests/testPoseWithCovariance.cpp); its Pose2 path
(Pcm2D / PcmSimple2D, reference include/KimeraRPGO/utils/GeometryUtils.h:117-186 over gtsam::Pose2) has none, and GTSAM is
not installable here.  What CAN be checked independently is that every closed-form expression the oracle restates
(compose, inverse, between, Logmap, AdjointMap and the covariance propagation built on them) equals its definition:
matrix products / inverses of the homogeneous matrices, scipy's matrix logarithm, and the adjoint defined by
hat(Ad_T xi) = T hat(xi) T^-1, in GTSAM's tangent orderings (Pose2: x, y, theta; Pose3: omega, v)."""
import numpy as np
import pytest
from scipy.linalg import logm

import orc


def mat(d, p):
    T = np.eye(d + 1)
    if d == 2:
        c, s, x, y = p
        T[:2, :2] = [[c, -s], [s, c]]
        T[:2, 2] = [x, y]
    else:
        T[:3, :3] = np.asarray(p[:9]).reshape(3, 3)
        T[:3, 3] = p[9:12]
    return T


def hat(d, xi):
    if d == 2:
        vx, vy, w = xi
        return np.array([[0, -w, vx], [w, 0, vy], [0, 0, 0.0]])
    wx, wy, wz, vx, vy, vz = xi
    return np.array([[0, -wz, wy, vx], [wz, 0, -wx, vy], [-wy, wx, 0, vz], [0, 0, 0, 0.0]])


def vee(d, X):
    if d == 2:
        return np.array([X[0, 2], X[1, 2], X[1, 0]])
    return np.array([X[2, 1], X[0, 2], X[1, 0], X[0, 3], X[1, 3], X[2, 3]])


def adjoint(d, T):
    n = orc.ndim(d)
    Ti = np.linalg.inv(T)
    return np.stack([vee(d, T @ hat(d, e) @ Ti) for e in np.eye(n)], axis=1)


def rand_pose(d, rng, rot=2.0, trans=5.0):
    if d == 2:
        return orc.pose2(rng.uniform(-rot, rot), rng.uniform(-trans, trans, size=2))
    w = rng.normal(size=3)
    w *= rng.uniform(0.05, rot) / np.linalg.norm(w)
    th = np.linalg.norm(w)
    K = hat(3, np.concatenate([w / th, np.zeros(3)]))[:3, :3]
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    return orc.pose3(R, rng.uniform(-trans, trans, size=3))


def rand_cov(d, rng, scale=1.0):
    n = orc.ndim(d)
    A = rng.normal(size=(n, n))
    return scale * (A @ A.T + n * np.eye(n))


@pytest.mark.parametrize("d", [2, 3])
def test_pose_algebra_equals_matrix_algebra(d):
    rng = np.random.default_rng(100 + d)
    for _ in range(200):
        a, b = rand_pose(d, rng), rand_pose(d, rng)
        assert np.allclose(mat(d, orc.pose_compose(d, a, b)), mat(d, a) @ mat(d, b), rtol=0, atol=1e-12)
        assert np.allclose(mat(d, orc.pose_inverse(d, a)), np.linalg.inv(mat(d, a)), rtol=0, atol=1e-12)


@pytest.mark.parametrize("d", [2, 3])
def test_logmap_is_the_matrix_logarithm(d):
    rng = np.random.default_rng(200 + d)
    for _ in range(200):
        p = rand_pose(d, rng, rot=2.8)
        want = vee(d, np.real(logm(mat(d, p))))
        assert np.allclose(orc.logmap(d, p), want, rtol=1e-9, atol=1e-10), (p, orc.logmap(d, p), want)


@pytest.mark.parametrize("d", [2, 3])
def test_compose_covariance_uses_the_adjoint_of_the_definition(d):
    """GeometryUtils.h:120-131: cov = Ha cov_a Ha^T + Hb cov_b Hb^T with gtsam's compose Jacobians Ha = Ad(b^-1), Hb = I"""
    rng = np.random.default_rng(300 + d)
    for _ in range(100):
        a, b = rand_pose(d, rng), rand_pose(d, rng)
        ca, cb = rand_cov(d, rng), rand_cov(d, rng)
        po, co, _ = orc.pwc_compose(d, (a, ca, 1), (b, cb, 1))
        Ha = adjoint(d, np.linalg.inv(mat(d, b)))
        assert np.allclose(mat(d, po), mat(d, a) @ mat(d, b), atol=1e-12)
        assert np.allclose(co, Ha @ ca @ Ha.T + cb, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("d", [2, 3])
def test_between_covariance_both_branches(d):
    """GeometryUtils.h:143-170: pose = a^-1 b, cov = cov_b - Ha cov_a Ha^T with Ha = -Ad((a^-1 b)^-1); when that is not
    positive definite (LLT), cov = cov_a - Ha' cov_b Ha'^T with Ha' = -Ad((b^-1 a)^-1) and the pose kept"""
    rng = np.random.default_rng(400 + d)
    seen = [0, 0]
    for k in range(200):
        a, b = rand_pose(d, rng), rand_pose(d, rng)
        big, small = (50.0, 0.01) if k % 2 == 0 else (0.01, 50.0)
        ca, cb = rand_cov(d, rng, small), rand_cov(d, rng, big)
        po, co, _ = orc.pwc_between(d, (a, ca, 1), (b, cb, 1))
        P = np.linalg.inv(mat(d, a)) @ mat(d, b)
        assert np.allclose(mat(d, po), P, atol=1e-11)
        Ha = -adjoint(d, np.linalg.inv(P))
        fwd = cb - Ha @ ca @ Ha.T
        if np.all(np.linalg.eigvalsh((fwd + fwd.T) / 2) > 1e-6):
            assert np.allclose(co, fwd, rtol=1e-10, atol=1e-9)
            seen[0] += 1
        else:
            P2 = np.linalg.inv(mat(d, b)) @ mat(d, a)
            Hb = -adjoint(d, np.linalg.inv(P2))
            assert np.allclose(co, ca - Hb @ cb @ Hb.T, rtol=1e-10, atol=1e-9)
            seen[1] += 1
    assert min(seen) >= 50  # both branches exercised


@pytest.mark.parametrize("d", [2, 3])
def test_inverse_keeps_the_covariance_and_mahalanobis_is_the_quadratic_form(d):
    rng = np.random.default_rng(500 + d)
    for _ in range(100):
        a, ca = rand_pose(d, rng), rand_cov(d, rng)
        pi, ci, _ = orc.pwc_inverse(d, (a, ca, 1))          # GeometryUtils.h:135-141: covariance copied unchanged
        assert np.array_equal(ci, ca) and np.allclose(mat(d, pi), np.linalg.inv(mat(d, a)), atol=1e-12)
        lg = vee(d, np.real(logm(mat(d, a))))
        want = np.sqrt(lg @ np.linalg.inv(ca) @ lg)          # GeometryUtils.h:172-186
        assert np.isclose(orc.pwc_mahalanobis(d, (a, ca, 1)), want, rtol=1e-9)


@pytest.mark.parametrize("d", [2, 3])
def test_pose_with_node_norms_take_head_and_tail_of_the_log_vector(d):
    """GeometryUtils.h:275-288: avg_trans_norm = |log.tail(t_dim)| / node, avg_rot_norm = |log.head(r_dim)| / node, literally:
    for Pose2 (Logmap order x, y, theta; r_dim = 1, t_dim = 2) the 'translation' norm is |(y, theta)| and the 'rotation' norm
    is |x| -- the reference's own quirk, which the oracle and the kernels reproduce"""
    rng = np.random.default_rng(600 + d)
    r_dim, t_dim = (1, 2) if d == 2 else (3, 3)
    for _ in range(100):
        p = rand_pose(d, rng)
        node = int(rng.integers(1, 40))
        lg = vee(d, np.real(logm(mat(d, p))))
        tr, ro = orc.pwn_norms(d, p, node)
        assert np.isclose(tr, np.linalg.norm(lg[-t_dim:]) / node, rtol=1e-9, atol=1e-12)
        assert np.isclose(ro, np.linalg.norm(lg[:r_dim]) / node, rtol=1e-9, atol=1e-12)
    assert orc.pwn_norms(d, rand_pose(d, rng), 3, rot=0)[1] == 0.0   # rotation_info = false: avg_rot_norm returns 0
