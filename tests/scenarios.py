"""Scenario builders shared by the oracle tests (CPU) and the GPU parity tests.

Each scenario is a transcription of a reference test (file:line cited per function) into a
list of update() calls: [(factors, values), ...] with factors = (type, key1, key2, pose, cov)
and values = (key, pose).  `expect` holds what the reference test asserts after given calls.
"""
import os

import numpy as np

import orc
from orc import BETWEEN, PRIOR, pose3, Rz, sym

I6 = np.eye(6)
R90 = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
RM90 = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1.0]])
R180 = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, 1.0]])


def _init_prior():
    return ([(PRIOR, 0, 0, pose3(), 0.01 * I6)], [(0, pose3())])


def _odom(i, R=None, t=(1, 0, 0), var=0.1):
    p = pose3(R, t)
    return ([(BETWEEN, i, i + 1, p, var * I6)], [(i + 1, p)])


def _lc(a, b, R, t, var=0.1):
    return ([(BETWEEN, a, b, pose3(R, t), var * I6)], [])


def pcm_odometry_check():
    """tests/testPcm.cpp:17-86.  Pcm3D, lc_threshold=-1, odom_threshold=0.3."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(3)]
    calls.append(_lc(3, 0, Rz(1.51), (0.8, 0, 0), 0.1))
    calls.append(_lc(3, 0, Rz(1.51), (0.8, 0, 0), 0.05))
    params = dict(lc_threshold=-1, odom_threshold=0.3)
    expect = {3: (4, 4), 4: (5, 4), 5: (5, 4)}  # after call index -> (nfg.size, est.size)
    return 3, 0, params, calls, expect


def pcm_consistency_check():
    """tests/testPcm.cpp:89-192.  Pcm3D, lc_threshold=0.5, odom_threshold=-1."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(2)] + [_odom(i, None) for i in range(2, 6)]
    calls.append(_lc(3, 0, Rz(3.1416), (0, 0.9, 0)))
    calls.append(_lc(4, 0, Rz(3.1416), (-1, 0.8, 0)))
    calls.append(_lc(5, 0, Rz(0.99 * 3.1416), (-1.8, 0.8, 0)))
    calls.append(_lc(6, 0, Rz(0.98 * 3.1416), (-2.6, 0.6, 0)))
    params = dict(lc_threshold=0.5, odom_threshold=-1)
    expect = {6: (7, 7), 8: (9, 7), 9: (10, 7), 10: (10, 7)}
    return 3, 0, params, calls, expect


def simple_odom_trans_check():
    """tests/testPcmSimple.cpp:18-96."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(3)]
    calls.append(_lc(3, 0, R90, (0.8, 0, 0), 0.1))
    calls.append(_lc(3, 0, R90, (0.792, 0, 0), 0.05))
    params = dict(dist_trans=-1, dist_rot=-1, odom_trans=0.051, odom_rot=100.0)
    expect = {3: (4, 4), 4: (5, 4), 5: (5, 4)}
    return 3, 1, params, calls, expect


def simple_odom_rot_check():
    """tests/testPcmSimple.cpp:99-171."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(3)]
    calls.append(_lc(3, 0, Rz(1.551), (1.0, 0, 0)))
    calls.append(_lc(3, 0, Rz(1.55), (1.0, 0, 0)))
    params = dict(dist_trans=-1, dist_rot=-1, odom_trans=100.0, odom_rot=0.005)
    expect = {3: (4, 4), 4: (5, 4), 5: (5, 4)}
    return 3, 1, params, calls, expect


def simple_consistency_trans_check():
    """tests/testPcmSimple.cpp:174-280 (lc4 excluded only because 0.05+1.1e-13 is not < 0.05)."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(2)] + [_odom(i, None) for i in range(2, 6)]
    calls.append(_lc(3, 0, R180, (0, 0.9, 0)))
    calls.append(_lc(4, 0, Rz(3.1416), (-0.9, 0.9, 0)))
    calls.append(_lc(5, 0, Rz(3.1416), (-1.9, 0.8, 0)))
    calls.append(_lc(6, 0, Rz(3.1416), (-2.8, 0.75, 0)))
    params = dict(odom_trans=-1, odom_rot=-1, dist_trans=0.05, dist_rot=100.0)
    expect = {6: (7, 7), 8: (9, 7), 9: (10, 7), 10: (10, 7)}
    return 3, 1, params, calls, expect


def simple_consistency_rot_check():
    """tests/testPcmSimple.cpp:283-389."""
    calls = [_init_prior()] + [_odom(i, R90) for i in range(2)] + [_odom(i, None) for i in range(2, 6)]
    calls.append(_lc(3, 0, R180, (0, 1.0, 0)))
    calls.append(_lc(4, 0, Rz(3.141), (-1.0, 1.0, 0)))
    calls.append(_lc(5, 0, Rz(3.13), (-2.0, 1.0, 0)))
    calls.append(_lc(6, 0, Rz(3.12), (-3.0, 1.0, 0)))
    params = dict(odom_trans=-1, odom_rot=-1, dist_trans=100.0, dist_rot=0.005)
    expect = {6: (7, 7), 8: (9, 7), 9: (10, 7), 10: (10, 7)}
    return 3, 1, params, calls, expect


def multi_robot(simple=False):
    """tests/testMultiRobot.cpp:22-188 (Pcm3D 3.0/0.05) and :191-357 (PcmSimple3D 0.04/0.01)."""
    a = lambda i: sym('a', i)
    b = lambda i: sym('b', i)
    calls = [([], [(a(0), pose3()), (b(0), pose3(None, (0, -1, 0)))])]
    od = pose3(None, (1, 0, 0))
    for i in range(3):
        calls.append(([(BETWEEN, a(i), a(i + 1), od, 0.1 * I6), (BETWEEN, b(i), b(i + 1), od, 0.1 * I6)],
                      [(a(i + 1), od), (b(i + 1), od)]))
    for i in range(3, 5):
        p = pose3(R90, (1, 0, 0))
        calls.append(([(BETWEEN, a(i), a(i + 1), p, 0.1 * I6)], [(a(i + 1), p)]))
    for i in range(3, 5):
        p = pose3(RM90, (1, 0, 0))
        calls.append(([(BETWEEN, b(i), b(i + 1), p, 0.1 * I6)], [(b(i + 1), p)]))
    n0 = len(calls)
    calls.append(([(BETWEEN, b(1), a(1), pose3(None, (0, 1, 0)), 0.1 * I6),
                   (BETWEEN, b(4), a(4), pose3(R180, (-1, 0, 0)), 0.1 * I6)], []))
    calls.append(([(BETWEEN, a(2), b(2), pose3(None, (0, -1, 0)), 0.1 * I6),
                   (BETWEEN, b(5), a(5), pose3(None, (0, -3.3, 0)), 0.1 * I6)], []))
    calls.append(([(BETWEEN, a(4), a(1), pose3(RM90, (0, 3, 0)), 0.1 * I6),
                   (BETWEEN, a(4), a(2), pose3(Rz(-1.54), (0, 2, 0)), 0.1 * I6)], []))
    if simple:
        params = dict(odom_trans=0.04, odom_rot=0.01, dist_trans=0.04, dist_rot=0.01)
    else:
        params = dict(odom_threshold=3.0, lc_threshold=0.05)
    expect = {n0: (12, 12), n0 + 1: (12, 12), n0 + 2: (13, 12)}
    return 3, (1 if simple else 0), params, calls, expect


def landmark_pcm(simple=False):
    """tests/testLandmark.cpp:23-196 (Pcm3D 5.0/2.5) — special symbol 'l', NaN rotation covariance (rotation_info=false)."""
    a = lambda i: sym('a', i)
    l = lambda i: sym('l', i)
    calls = [([], [(a(0), pose3())])]
    for i in range(2):
        p = pose3(None, (1, 0, 0))
        calls.append(([(BETWEEN, a(i), a(i + 1), p, 0.1 * I6)], [(a(i + 1), p)]))
    for i in range(2, 5):
        p = pose3(R90, (1, 0, 0))
        calls.append(([(BETWEEN, a(i), a(i + 1), p, 0.1 * I6)], [(a(i + 1), p)]))
    calls.append(([(BETWEEN, a(3), a(2), pose3(Rz(-1.57), (0, 0.9, 0)), 0.1 * I6),
                   (BETWEEN, a(4), a(1), pose3(Rz(3.14), (2.1, 1.1, 2.5)), 0.1 * I6)], []))
    n0 = len(calls) - 1
    lcov = np.full((6, 6), 0.0)
    lcov[:3, :3] = np.nan       # Diagonal::Precisions((0,0,0,25,25,25)).covariance(): no rotation information
    lcov[3:, 3:] = np.eye(3) / 25.0
    calls.append(([(BETWEEN, a(1), l(0), pose3(None, (0, 1, 0)), lcov)], [(l(0), pose3(None, (1, 1, 0)))]))
    calls.append(([(BETWEEN, a(5), l(0), pose3(None, (0, -1, 0)), lcov)], []))
    calls.append(([(BETWEEN, a(4), l(0), pose3(None, (1, 0, 0)), lcov)], []))
    calls.append(([(BETWEEN, a(2), l(1), pose3(None, (0, -1, 0)), lcov)], [(l(1), pose3(None, (1, 1, 0)))]))
    calls.append(([(BETWEEN, a(5), l(1), pose3(None, (2, 0, 0)), lcov)], []))
    if simple:  # tests/testLandmark.cpp:186-348: same graph, setPcmSimple3DParams(0.3, 0.05)
        params = dict(odom_trans=0.3, odom_rot=0.05, dist_trans=0.3, dist_rot=0.05, special_symbols=('l',))
    else:
        params = dict(odom_threshold=5.0, lc_threshold=2.5, special_symbols=('l',))
    expect = {n0: (6, 6), n0 + 1: (7, 7), n0 + 2: (8, 7), n0 + 3: (8, 7), n0 + 4: (9, 8), n0 + 5: (10, 8)}
    return 3, (1 if simple else 0), params, calls, expect


_G2O = None


def g2o_fixture(name):
    """Arrays parsed from the reference's tests/data/<name>.g2o (tools/make_golden_from_reference.py)."""
    global _G2O
    if _G2O is None:
        _G2O = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g2o_fixtures.npz"))
    g = _G2O
    values = [(int(k), p) for k, p in zip(g[name + "_vkeys"], g[name + "_vposes"])]
    edges = [(BETWEEN, int(a), int(b), p, c) for a, b, p, c in
             zip(g[name + "_k1"], g[name + "_k2"], g[name + "_eposes"], g[name + "_ecovs"])]
    return values, edges


def load_graph_calls(which):
    """tests/testLoadGraph.cpp: Load1 (:25-58) 'load1', Add1 (:61-124) 'add1', Load2 (:127-161) 'load2',
    Add2 (:164-221) 'add2'.  Returns (d, mode, params, calls, expect)."""
    va, ea = g2o_fixture("robot_a")
    vb, eb = g2o_fixture("robot_b")
    prior = (PRIOR, sym('a', 0), sym('a', 0), dict(va)[sym('a', 0)], 0.01 * I6)
    inter = [(BETWEEN, sym('a', 1), sym('b', 1), pose3(), 0.01 * I6),
             (BETWEEN, sym('a', 2), sym('b', 2), pose3(), 0.01 * I6)]
    if which == "load1":
        return 3, 0, dict(odom_threshold=0.0, lc_threshold=10.0), [(ea + [prior], va)], {0: (50, 50)}
    if which == "add1":
        return 3, 0, dict(odom_threshold=0.0, lc_threshold=0.0), [(ea + [prior], va), (eb, vb), (inter, [])], \
            {1: (91, 92), 2: (92, 92)}
    if which == "load2":
        return 3, 0, dict(odom_threshold=100.0, lc_threshold=100.0), [(ea + [prior], va)], {0: (53, 50)}
    if which == "add2":
        return 3, 0, dict(odom_threshold=100.0, lc_threshold=100.0), [(ea + [prior], va), (eb, vb), (inter[:1], [])], \
            {1: (96, 92), 2: (97, 92)}
    raise KeyError(which)


ALL = {
    "pcm_odometry_check": pcm_odometry_check,
    "pcm_consistency_check": pcm_consistency_check,
    "simple_odom_trans_check": simple_odom_trans_check,
    "simple_odom_rot_check": simple_odom_rot_check,
    "simple_consistency_trans_check": simple_consistency_trans_check,
    "simple_consistency_rot_check": simple_consistency_rot_check,
    "multi_robot_pcm": lambda: multi_robot(False),
    "multi_robot_simple": lambda: multi_robot(True),
    "landmark_pcm": landmark_pcm,
    "landmark_simple": lambda: landmark_pcm(True),
    "load1": lambda: load_graph_calls("load1"),
    "add1": lambda: load_graph_calls("add1"),
    "load2": lambda: load_graph_calls("load2"),
    "add2": lambda: load_graph_calls("add2"),
}
