"""CPU-only: the product's block-structured math header, compiled for the host, must agree BIT FOR BIT
with the dense oracle on random inputs (this pins the summation order the CUDA kernels use)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
from orc import dp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_cpu_math_shim.so")


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(ROOT, "tests", "cpu_math_shim.cpp")
    hdr = os.path.join(ROOT, "kimera-rpgo_b200", "csrc", "rpgo_math.cuh")
    hdr2 = os.path.join(ROOT, "kimera-rpgo_b200", "csrc", "rpgo_pair_v2.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read() else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++"] + fma +
                              ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "kimera-rpgo_b200", "csrc"),
                               src, "-o", SO, "-lm"])
    L = C.CDLL(SO)
    L.shim_entry_size.restype = C.c_int
    L.shim_pair_check.restype = C.c_int
    L.shim_pair_check_v1.restype = C.c_int
    L.shim_pair_check_v2.restype = C.c_int
    L.shim_pair_check_simple_v2.restype = C.c_int
    return L


def rand_pose(rng, d, scale=5.0):
    if d == 3:
        q = rng.normal(size=4)
        R = orc.quat_R(*q)
        return orc.pose3(R, rng.uniform(-scale, scale, size=3))
    th = rng.uniform(-np.pi, np.pi)
    return orc.pose2(th, rng.uniform(-scale, scale, size=2))


def rand_cov(rng, n, scale):
    a = rng.normal(size=(n, n))
    return scale * (a @ a.T / n + 0.05 * np.eye(n))


def entry(L, d, pose, cov, rot=1, node=0):
    E = L.shim_entry_size(d)
    e = np.zeros(E)
    ps, n = orc.psize(d), orc.ndim(d)
    e[:ps] = pose
    oc = 12 if d == 3 else 4
    e[oc:oc + n * n] = np.asarray(cov).reshape(-1)
    e[48 if d == 3 else 13] = rot
    e[49 if d == 3 else 14] = node
    return e


@pytest.mark.parametrize("d", [3, 2])
def test_compose_between_bitwise(shim, d):
    rng = np.random.default_rng(100 + d)
    n = orc.ndim(d)
    E = shim.shim_entry_size(d)
    oc = 12 if d == 3 else 4
    for t in range(300):
        pa, pb = rand_pose(rng, d), rand_pose(rng, d)
        ca, cb = rand_cov(rng, n, 0.01), rand_cov(rng, n, 0.01)
        if t % 2:
            cb = ca + rand_cov(rng, n, 0.001)  # exercise both LLT outcomes
        out = np.zeros(E)
        shim.shim_compose(d, 0, dp(entry(shim, d, pa, ca)), dp(entry(shim, d, pb, cb)), dp(out))
        po, co, _ = orc.pwc_compose(d, (pa, ca, 1), (pb, cb, 1))
        assert np.array_equal(out[:orc.psize(d)], po) and np.array_equal(out[oc:oc + n * n].reshape(n, n), co), t
        shim.shim_between(d, 0, dp(entry(shim, d, pa, ca)), dp(entry(shim, d, pb, cb)), dp(out))
        po, co, _ = orc.pwc_between(d, (pa, ca, 1), (pb, cb, 1))
        assert np.array_equal(out[:orc.psize(d)], po) and np.array_equal(out[oc:oc + n * n].reshape(n, n), co), t


@pytest.mark.parametrize("d", [3, 2])
def test_pair_check_pcm_bitwise(shim, d):
    """areLoopsConsistent chain (Pcm.h:703-717) + mahalanobis: product header == oracle, bit for bit."""
    rng = np.random.default_rng(200 + d)
    n = orc.ndim(d)
    thr = np.array([1.0, 3.0, 1, 1, 1, 1.0])
    n_bad = [0]
    for t in range(400):
        P = [rand_pose(rng, d, 3.0) for _ in range(6)]
        Ccum = [rand_cov(rng, n, 0.02) for _ in range(2)]
        # trajectory covariances: later = earlier + something (PSD difference) or the other way round
        Ta = (P[0], Ccum[0], 1); Tc = (P[1], Ccum[0] + (1 if t % 2 else -0.5) * rand_cov(rng, n, 0.004), 1)
        Tb = (P[2], Ccum[1], 1); Td = (P[3], Ccum[1] + (1 if t % 3 else -0.5) * rand_cov(rng, n, 0.004), 1)
        rot_i = 0 if t % 11 == 0 else 1
        lci = (P[4], rand_cov(rng, n, 0.01), rot_i); lcj = (P[5], rand_cov(rng, n, 0.01), 1)
        # oracle chain
        a_odom_c = orc.pwc_between(d, Ta, Tc)
        b_odom_d = orc.pwc_between(d, Tb, Td)
        a_path_d = orc.pwc_compose(d, a_odom_c, lcj)
        d_path_b = orc.pwc_compose(d, orc.pwc_inverse(d, a_path_d), lci)
        loop = orc.pwc_compose(d, d_path_b, b_odom_d)
        want = orc.pwc_mahalanobis(d, loop)
        dist = C.c_double(); near = C.c_int()
        E = [entry(shim, d, *x) for x in (Ta, Tb, lci, Tc, Td, lcj)]
        ok = shim.shim_pair_check(d, 0, *[dp(e) for e in E], dp(thr), C.byref(dist), C.byref(near))
        assert (dist.value == want) or (np.isnan(dist.value) and np.isnan(want)), (t, dist.value, want)
        assert bool(ok) == bool(want < thr[1])
        # the restructured (tiled-kernel) pair function must give the same bits
        d1 = C.c_double()
        ok1 = shim.shim_pair_check_v1(d, *[dp(e) for e in E], dp(thr), C.byref(d1), C.byref(near))
        assert (d1.value == want) or (np.isnan(d1.value) and np.isnan(want)), ("v1", t, d1.value, want)
        assert bool(ok1) == bool(ok)
        # straight-line form: either it flags the lane for the exact path, or it agrees bit for bit
        d2 = C.c_double(); bad = C.c_int()
        ok2 = shim.shim_pair_check_v2(d, *[dp(e) for e in E], dp(thr), C.byref(d2), C.byref(bad))
        if bad.value:
            n_bad[0] += 1
        else:
            assert d2.value == want, ("v2", t, d2.value, want)
            assert bool(ok2) == bool(ok)
    assert n_bad[0] < 400  # random (non-trajectory) covariances are mostly indefinite differences: flagged, never wrong


@pytest.mark.parametrize("d", [3, 2])
def test_pair_check_simple_bitwise(shim, d):
    rng = np.random.default_rng(300 + d)
    n = orc.ndim(d)
    L = orc.lib()
    thr = np.array([1.0, 3.0, 1, 1, 0.05, 0.01])
    n_flag = [0]
    for t in range(300):
        P = [rand_pose(rng, d, 1.0) for _ in range(6)]
        nodes = rng.integers(0, 50, size=4)
        # oracle via its pose primitives
        btw = lambda a, b: orc.pose_compose(d, orc.pose_inverse(d, a), b)
        a_odom_c, nac = btw(P[0], P[1]), abs(int(nodes[1]) - int(nodes[0]))
        b_odom_d, nbd = btw(P[2], P[3]), abs(int(nodes[3]) - int(nodes[2]))
        a_path_d = orc.pose_compose(d, a_odom_c, P[5])
        d_path_b = orc.pose_compose(d, orc.pose_inverse(d, a_path_d), P[4])
        loop = orc.pose_compose(d, d_path_b, b_odom_d)
        node = nac + 1 + 1 + nbd
        wt, wr = orc.pwn_norms(d, loop, node)
        dist = C.c_double(); near = C.c_int()
        z = np.zeros((n, n))
        E = [entry(shim, d, P[0], z, 1, nodes[0]), entry(shim, d, P[2], z, 1, nodes[2]), entry(shim, d, P[4], z, 1, 1),
             entry(shim, d, P[1], z, 1, nodes[1]), entry(shim, d, P[3], z, 1, nodes[3]), entry(shim, d, P[5], z, 1, 1)]
        ok = shim.shim_pair_check(d, 1, *[dp(e) for e in E], dp(thr), C.byref(dist), C.byref(near))
        assert dist.value == wt, (t, dist.value, wt)
        assert bool(ok) == bool(wt < thr[4] and wr < thr[5])
        # the tiled kernel's forms on compact entries: exact fallback == general code; straight-line form either flags
        # the lane or agrees bit for bit (distance, decision, near flag)
        for exact in (1, 0):
            d2 = C.c_double(); near2 = C.c_int(); bad = C.c_int()
            ok2 = shim.shim_pair_check_simple_v2(d, *[dp(e) for e in E], dp(thr), C.byref(d2), C.byref(near2), C.byref(bad), exact)
            if exact or not bad.value:
                assert (d2.value == wt) or (np.isnan(d2.value) and np.isnan(wt)), (t, exact, d2.value, wt)
                assert bool(ok2) == bool(ok) and near2.value == near.value
            n_flag[0] += bad.value
    assert n_flag[0] < 60, n_flag   # zero hop counts (nodes drawn equal) and the like: flagged, never wrong


def test_div_by_equals_ieee_division(shim):
    """div_by (shared correctly-rounded reciprocal + FMA correction) == IEEE a / x on 10^7 operands,
    including significands near 1 and 2 (all-ones / all-zeros mantissas) and wide exponent ranges."""
    rng = np.random.default_rng(77)
    shim.shim_div_by_mismatches.restype = C.c_longlong
    shim.shim_div_by_mismatches.argtypes = [C.c_longlong, orc.c_dp, orc.c_dp]
    n = 2_500_000
    total = 0
    for kind in range(4):
        if kind == 0:
            a = rng.normal(size=n) * 10.0 ** rng.uniform(-8, 8, size=n)
            x = rng.normal(size=n) * 10.0 ** rng.uniform(-8, 8, size=n)
        elif kind == 1:  # mantissas hugging powers of two
            ma = rng.integers(0, 64, size=n).astype(np.uint64)
            mx = rng.integers(0, 64, size=n).astype(np.uint64)
            top = np.uint64(0x000FFFFFFFFFFFFF)
            abits = (np.uint64(1023) << np.uint64(52)) | np.where(rng.random(n) < 0.5, ma, top - ma)
            xbits = (np.uint64(1023) << np.uint64(52)) | np.where(rng.random(n) < 0.5, mx, top - mx)
            a = abits.view(np.float64) * 2.0 ** rng.integers(-40, 40, size=n)
            x = xbits.view(np.float64) * 2.0 ** rng.integers(-40, 40, size=n)
        elif kind == 2:  # exact and nearly exact quotients
            x = rng.integers(1, 1 << 20, size=n).astype(np.float64)
            qq = rng.integers(1, 1 << 30, size=n).astype(np.float64)
            a = x * qq + rng.integers(-1, 2, size=n)
        else:  # random bit patterns of moderate exponent
            a = (rng.integers(0, 1 << 52, size=n, dtype=np.uint64) | (np.uint64(1000 + 46) << np.uint64(52))).view(np.float64)
            x = (rng.integers(0, 1 << 52, size=n, dtype=np.uint64) | (np.uint64(1000 + 11) << np.uint64(52))).view(np.float64)
        a = np.ascontiguousarray(a); x = np.ascontiguousarray(x)
        total += shim.shim_div_by_mismatches(n, dp(a), dp(x))
    assert total == 0


def test_pair_check_v2_on_trajectory_data(shim):
    """On realistic inputs (a synthetic helix graph) the straight-line form must agree bit for bit with the oracle's
    pairwise distances and flag (almost) nothing for the exact path."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    synth = importlib.import_module("kimera-rpgo_b200.synth")
    gph = synth.config2(seed=3, P=500, n=60, outlier_frac=0.4)
    o = orc.OraclePcm(3, 0, odom_threshold=-1, lc_threshold=5.0)
    o.update(gph["odom"], gph["values"])
    o.update(gph["lcs"], [])
    _, dist = o.group_adj(0)
    thr = np.array([1.0, 5.0, 1, 1, 1, 1.0])
    ent = []
    for f in gph["lcs"]:
        pf, cf, _, rf = o.traj_get(f[1])
        pb, cb, _, rb = o.traj_get(f[2])
        ent.append((entry(shim, 3, pf, cf, rf), entry(shim, 3, pb, cb, rb), entry(shim, 3, f[3], f[4], 1)))
    n_bad = 0
    n = len(ent)
    for i in range(n):
        for j in range(i + 1, n):
            d2 = C.c_double(); bad = C.c_int()
            shim.shim_pair_check_v2(3, dp(ent[i][0]), dp(ent[i][1]), dp(ent[i][2]), dp(ent[j][0]), dp(ent[j][1]), dp(ent[j][2]),
                                    dp(thr), C.byref(d2), C.byref(bad))
            if bad.value:
                n_bad += 1
            else:
                assert d2.value == dist[i, j], (i, j, d2.value, dist[i, j])
    assert n_bad <= 0.02 * n * (n - 1) / 2, n_bad
