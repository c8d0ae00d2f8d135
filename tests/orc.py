"""ctypes bindings to the CPU oracle (oracle/liboracle.so) and to the reference's own FMC
clique finder (oracle/_ref/libref_fmc.so).  TEST INFRASTRUCTURE ONLY — imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product package."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None
_ref = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_u8p = C.POINTER(C.c_ubyte)
c_u64p = C.POINTER(C.c_uint64)
c_llp = C.POINTER(C.c_longlong)


def dp(a):
    return a.ctypes.data_as(c_dp)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L = _lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, c_dp, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_reference_shaped.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_special_symbols.argtypes = [C.c_void_p, C.c_int, c_u8p]
        L.orc_frame_align.restype = C.c_int
        L.orc_frame_align.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
        L.orc_robot_odom_values.restype = C.c_int
        L.orc_robot_odom_values.argtypes = [C.c_void_p, C.c_int, c_dp, c_u64p, c_dp]
        L.orc_num_landmarks.restype = C.c_int
        L.orc_num_landmarks.argtypes = [C.c_void_p]
        L.orc_landmark_info.argtypes = [C.c_void_p, C.c_int, c_u64p, c_ip, c_ip]
        L.orc_landmark_adj.restype = C.c_int
        L.orc_landmark_adj.argtypes = [C.c_void_p, C.c_int, c_u8p, c_dp]
        L.orc_landmark_ids.argtypes = [C.c_void_p, C.c_int, c_llp, c_llp]
        L.orc_check_pairs.restype = None
        L.orc_check_pairs.argtypes = [C.c_void_p, C.c_int, c_u64p, c_u64p, c_dp, c_dp, C.c_longlong, c_ip, c_ip, c_u8p, c_dp, c_u8p]
        L.orc_update.restype = C.c_int
        L.orc_update.argtypes = [C.c_void_p, C.c_int, c_ip, c_u64p, c_u64p, c_dp, c_dp, C.c_int, c_u64p, c_dp]
        for n in ["orc_num_lc", "orc_num_inliers", "orc_num_odom", "orc_num_special", "orc_num_values",
                  "orc_pair_checks", "orc_output_size", "orc_num_flagged"]:
            getattr(L, n).restype = C.c_longlong
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_output_ids.argtypes = [C.c_void_p, c_llp]
        L.orc_num_groups.restype = C.c_int
        L.orc_num_groups.argtypes = [C.c_void_p]
        L.orc_group_info.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip, c_ip, c_ip]
        L.orc_group_adj.restype = C.c_int
        L.orc_group_adj.argtypes = [C.c_void_p, C.c_int, c_u8p, c_dp]
        L.orc_group_factor_ids.argtypes = [C.c_void_p, C.c_int, c_llp]
        L.orc_group_inlier_ids.argtypes = [C.c_void_p, C.c_int, c_llp]
        L.orc_flagged.argtypes = [C.c_void_p, c_llp]
        L.orc_remove_last.restype = C.c_int
        L.orc_remove_last.argtypes = [C.c_void_p, C.c_int, C.c_int, c_u64p, c_u64p]
        L.orc_remove_last_any.restype = C.c_int
        L.orc_remove_last_any.argtypes = [C.c_void_p, c_u64p, c_u64p]
        L.orc_ignore_prefix.argtypes = [C.c_void_p, C.c_int]
        L.orc_revive_prefix.argtypes = [C.c_void_p, C.c_int]
        L.orc_traj_get.restype = C.c_int
        L.orc_traj_get.argtypes = [C.c_void_p, C.c_uint64, c_dp, c_dp, c_ip, c_ip]
        L.orc_pose_compose.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.orc_pose_inverse.argtypes = [C.c_int, c_dp, c_dp]
        L.orc_logmap.argtypes = [C.c_int, c_dp, c_dp]
        L.orc_pwc_compose.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_ip]
        L.orc_pwc_between.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_ip]
        L.orc_pwc_inverse.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_ip]
        L.orc_pwc_mahalanobis.restype = C.c_double
        L.orc_pwc_mahalanobis.argtypes = [C.c_int, c_dp, c_dp, C.c_int]
        L.orc_pwn_norms.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_llt_ok.restype = C.c_int
        L.orc_llt_ok.argtypes = [C.c_int, c_dp]
        L.orc_lu_inverse.argtypes = [C.c_int, c_dp, c_dp]
        L.orc_traj_fold.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]
        L.orc_clique_heu.restype = C.c_int
        L.orc_clique_heu.argtypes = [C.c_int, c_u8p, c_ip]
        L.orc_clique_heu_incremental.restype = C.c_int
        L.orc_clique_heu_incremental.argtypes = [C.c_int, c_u8p, C.c_int, C.c_int, c_ip]
        L.orc_clique_exact.restype = C.c_int
        L.orc_clique_exact.argtypes = [C.c_int, c_u8p, c_ip]
    return _lib


def ref_fmc():
    """The reference's own FMC, compiled from /root/reference by oracle/Makefile (None if absent)."""
    global _ref
    p = os.path.join(ROOT, "oracle", "_ref", "libref_fmc.so")
    if _ref is None and os.path.exists(p):
        _ref = C.CDLL(p)
        _ref.ref_find_max_clique_heu.restype = C.c_int
        _ref.ref_find_max_clique_heu.argtypes = [C.c_int, c_u8p, c_ip, C.c_int, c_ip]
        _ref.ref_find_max_clique_heu_incremental.restype = C.c_int
        _ref.ref_find_max_clique_heu_incremental.argtypes = [C.c_int, c_u8p, C.c_int, C.c_int, c_ip, C.c_int, c_ip]
        _ref.ref_find_max_clique.restype = C.c_int
        _ref.ref_find_max_clique.argtypes = [C.c_int, c_u8p, c_ip, C.c_int, c_ip]
    return _ref


# ---- pose helpers (storage: 3D = R row-major(9) + t(3); 2D = c, s, x, y) -------------------
def psize(d):
    return 12 if d == 3 else 4


def ndim(d):
    return 6 if d == 3 else 3


def pose3(R=None, t=(0, 0, 0)):
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    return np.concatenate([R.reshape(9), np.asarray(t, dtype=np.float64)])


def pose2(theta=0.0, t=(0, 0)):
    return np.array([np.cos(theta), np.sin(theta), t[0], t[1]], dtype=np.float64)


def quat_R(w, x, y, z):
    """gtsam::Rot3(w,x,y,z) -> Eigen quaternion toRotationMatrix (quaternion normalised)."""
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def Rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def Ry(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def Rx(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def sym(ch, idx):
    """gtsam::Symbol(chr, index) -> Key."""
    return (ord(ch) << 56) | int(idx)


# ---- primitive wrappers -------------------------------------------------------------------
def pwc_compose(d, a, b):
    n = ndim(d)
    po = np.zeros(psize(d)); co = np.zeros((n, n)); ro = C.c_int(1)
    pa, ca, ra = a; pb, cb, rb = b
    lib().orc_pwc_compose(d, dp(np.ascontiguousarray(pa)), dp(np.ascontiguousarray(ca)), ra,
                          dp(np.ascontiguousarray(pb)), dp(np.ascontiguousarray(cb)), rb, dp(po), dp(co), C.byref(ro))
    return po, co, ro.value


def pwc_between(d, a, b):
    n = ndim(d)
    po = np.zeros(psize(d)); co = np.zeros((n, n)); ro = C.c_int(1)
    pa, ca, ra = a; pb, cb, rb = b
    lib().orc_pwc_between(d, dp(np.ascontiguousarray(pa)), dp(np.ascontiguousarray(ca)), ra,
                          dp(np.ascontiguousarray(pb)), dp(np.ascontiguousarray(cb)), rb, dp(po), dp(co), C.byref(ro))
    return po, co, ro.value


def pwc_inverse(d, a):
    n = ndim(d)
    po = np.zeros(psize(d)); co = np.zeros((n, n)); ro = C.c_int(1)
    pa, ca, ra = a
    lib().orc_pwc_inverse(d, dp(np.ascontiguousarray(pa)), dp(np.ascontiguousarray(ca)), ra, dp(po), dp(co), C.byref(ro))
    return po, co, ro.value


def pwc_mahalanobis(d, a):
    pa, ca, ra = a
    return lib().orc_pwc_mahalanobis(d, dp(np.ascontiguousarray(pa)), dp(np.ascontiguousarray(ca)), ra)


def pose_compose(d, a, b):
    out = np.zeros(psize(d))
    lib().orc_pose_compose(d, dp(np.ascontiguousarray(a)), dp(np.ascontiguousarray(b)), dp(out))
    return out


def pose_inverse(d, a):
    out = np.zeros(psize(d))
    lib().orc_pose_inverse(d, dp(np.ascontiguousarray(a)), dp(out))
    return out


def logmap(d, a):
    out = np.zeros(ndim(d))
    lib().orc_logmap(d, dp(np.ascontiguousarray(a)), dp(out))
    return out


def pwn_norms(d, pose, node, rot=1):
    t = C.c_double(); r = C.c_double()
    lib().orc_pwn_norms(d, dp(np.ascontiguousarray(pose)), node, rot, C.byref(t), C.byref(r))
    return t.value, r.value


def traj_fold(d, mode, init_pose, dpose, dcov):
    P = len(dpose) + 1
    n = ndim(d)
    cp = np.zeros((P, psize(d))); cc = np.zeros((P, n, n)); cn = np.zeros(P, dtype=np.int32); cr = np.zeros(P, dtype=np.int32)
    dpose = np.ascontiguousarray(dpose, dtype=np.float64); dcov = np.ascontiguousarray(dcov, dtype=np.float64)
    lib().orc_traj_fold(d, mode, P, dp(np.ascontiguousarray(init_pose, dtype=np.float64)), dp(dpose), dp(dcov), dp(cp), dp(cc),
                        cn.ctypes.data_as(c_ip), cr.ctypes.data_as(c_ip))
    return cp, cc, cn, cr


def clique_heu(adj):
    adj = np.ascontiguousarray(adj, dtype=np.uint8); n = adj.shape[0]
    ids = np.zeros(max(n, 1), dtype=np.int32)
    k = lib().orc_clique_heu(n, adj.ctypes.data_as(c_u8p), ids.ctypes.data_as(c_ip))
    return k, ids[:max(k, 0)].copy()


def clique_heu_incremental(adj, num_new, prev):
    adj = np.ascontiguousarray(adj, dtype=np.uint8); n = adj.shape[0]
    ids = np.zeros(max(n, 1), dtype=np.int32)
    k = lib().orc_clique_heu_incremental(n, adj.ctypes.data_as(c_u8p), num_new, prev, ids.ctypes.data_as(c_ip))
    return k, ids[:max(k, 0)].copy()


def clique_exact(adj):
    adj = np.ascontiguousarray(adj, dtype=np.uint8); n = adj.shape[0]
    ids = np.zeros(max(n, 1), dtype=np.int32)
    k = lib().orc_clique_exact(n, adj.ctypes.data_as(c_u8p), ids.ctypes.data_as(c_ip))
    return k, ids[:max(k, 0)].copy()


def _ref_call(fn, adj, *extra):
    adj = np.ascontiguousarray(adj, dtype=np.uint8); n = adj.shape[0]
    buf = np.zeros(n + 2, dtype=np.int32); blen = C.c_int(0)
    k = fn(n, adj.ctypes.data_as(c_u8p), *extra, buf.ctypes.data_as(c_ip), n + 2, C.byref(blen))
    return k, buf[:max(min(k, blen.value), 0)].copy()


def ref_clique_heu(adj):
    return _ref_call(ref_fmc().ref_find_max_clique_heu, adj)


def ref_clique_heu_incremental(adj, num_new, prev):
    return _ref_call(ref_fmc().ref_find_max_clique_heu_incremental, adj, num_new, prev)


def ref_clique_exact(adj):
    return _ref_call(ref_fmc().ref_find_max_clique, adj)


# ---- pipeline wrapper: mirrors the OutlierRemoval interface at factor level ------------------
BETWEEN, PRIOR, OTHER = 0, 1, 2


class OraclePcm:
    """Pcm<poseT, T> restated (reference: include/KimeraRPGO/outlier/Pcm.h).  d: 2|3, mode: 0 PCM, 1 Simple."""

    def __init__(self, d, mode, odom_threshold=10.0, lc_threshold=5.0, odom_trans=0.05, odom_rot=0.005,
                 dist_trans=0.01, dist_rot=0.001, incremental=False, special_symbols=()):
        self.d, self.mode = d, mode
        thr = np.array([odom_threshold, lc_threshold, odom_trans, odom_rot, dist_trans, dist_rot], dtype=np.float64)
        self.h = lib().orc_create(d, mode, dp(thr), int(incremental))
        if special_symbols:
            sy = np.array([ord(c) for c in special_symbols], dtype=np.uint8)
            lib().orc_set_special_symbols(self.h, len(sy), sy.ctypes.data_as(c_u8p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def set_reference_shaped(self, on):
        lib().orc_set_reference_shaped(self.h, int(on))

    def update(self, factors, values):
        """factors: list of (type, key1, key2, pose, cov); values: list of (key, pose).  Returns do_optimize."""
        d, n, ps = self.d, ndim(self.d), psize(self.d)
        nf = len(factors)
        types = np.array([f[0] for f in factors], dtype=np.int32)
        k1 = np.array([f[1] for f in factors], dtype=np.uint64)
        k2 = np.array([f[2] for f in factors], dtype=np.uint64)
        poses = np.zeros((max(nf, 1), ps)); covs = np.zeros((max(nf, 1), n, n))
        for i, f in enumerate(factors):
            poses[i] = f[3]
            covs[i] = f[4]
        nv = len(values)
        vk = np.array([v[0] for v in values], dtype=np.uint64)
        vp = np.zeros((max(nv, 1), ps))
        for i, v in enumerate(values):
            vp[i] = v[1]
        if nf == 0:
            types = np.zeros(1, dtype=np.int32); k1 = np.zeros(1, dtype=np.uint64); k2 = np.zeros(1, dtype=np.uint64)
        if nv == 0:
            vk = np.zeros(1, dtype=np.uint64)
        return bool(lib().orc_update(self.h, nf, types.ctypes.data_as(c_ip), k1.ctypes.data_as(c_u64p),
                                     k2.ctypes.data_as(c_u64p), dp(poses), dp(covs), nv, vk.ctypes.data_as(c_u64p), dp(vp)))

    def update_arrays(self, k1, k2, poses, covs, vkeys=None, vposes=None):
        """update() for BetweenFactors given as arrays (what synth.as_arrays produces): no per-factor Python work."""
        nf = len(k1)
        ps, nn = psize(self.d), ndim(self.d) ** 2
        types = np.zeros(max(nf, 1), dtype=np.int32)
        k1 = np.ascontiguousarray(k1, dtype=np.uint64) if nf else np.zeros(1, dtype=np.uint64)
        k2 = np.ascontiguousarray(k2, dtype=np.uint64) if nf else np.zeros(1, dtype=np.uint64)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(max(nf, 1), ps) if nf else np.zeros((1, ps))
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(max(nf, 1), nn) if nf else np.zeros((1, nn))
        nv = 0 if vkeys is None else len(vkeys)
        vk = np.ascontiguousarray(vkeys, dtype=np.uint64) if nv else np.zeros(1, dtype=np.uint64)
        vp = np.ascontiguousarray(vposes, dtype=np.float64).reshape(nv, ps) if nv else np.zeros((1, ps))
        return bool(lib().orc_update(self.h, nf, types.ctypes.data_as(c_ip), k1.ctypes.data_as(c_u64p),
                                     k2.ctypes.data_as(c_u64p), dp(poses), dp(covs), nv, vk.ctypes.data_as(c_u64p), dp(vp)))

    def check_pairs(self, k1, k2, poses, covs, pi, pj):
        """areLoopsConsistent (Pcm.h:670-718) for the pairs (pi[t] older, pj[t] newer) of a closure table given as arrays
        (keys, n x ps poses, n x N*N covariances), against the trajectories folded so far.
        Returns (ok uint8[m], dist float64[m], in_band uint8[m])."""
        n = len(k1)
        k1 = np.ascontiguousarray(k1, dtype=np.uint64)
        k2 = np.ascontiguousarray(k2, dtype=np.uint64)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(n, psize(self.d))
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(n, ndim(self.d) ** 2)
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
        m = len(pi)
        ok = np.zeros(max(m, 1), dtype=np.uint8)
        band = np.zeros(max(m, 1), dtype=np.uint8)
        dist = np.zeros(max(m, 1))
        lib().orc_check_pairs(self.h, n, k1.ctypes.data_as(c_u64p), k2.ctypes.data_as(c_u64p), dp(poses), dp(covs), m,
                              pi.ctypes.data_as(c_ip), pj.ctypes.data_as(c_ip), ok.ctypes.data_as(c_u8p), dp(dist),
                              band.ctypes.data_as(c_u8p))
        return ok[:m], dist[:m], band[:m]

    def nfg_size(self):
        return lib().orc_output_size(self.h)

    def output_ids(self):
        n = self.nfg_size()
        ids = np.zeros(max(n, 1), dtype=np.int64)
        lib().orc_output_ids(self.h, ids.ctypes.data_as(c_llp))
        return ids[:n]

    def num_values(self):
        return lib().orc_num_values(self.h)

    def num_lc(self):
        return lib().orc_num_lc(self.h)

    def num_inliers(self):
        return lib().orc_num_inliers(self.h)

    def pair_checks(self):
        return lib().orc_pair_checks(self.h)

    def groups(self):
        out = []
        for g in range(lib().orc_num_groups(self.h)):
            c1, c2, n, ni = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            lib().orc_group_info(self.h, g, C.byref(c1), C.byref(c2), C.byref(n), C.byref(ni))
            out.append((chr(c1.value), chr(c2.value), n.value, ni.value))
        return out

    def group_adj(self, g):
        n = lib().orc_group_adj(self.h, g, None, None)
        adj = np.zeros((n, n), dtype=np.uint8); dist = np.zeros((n, n))
        lib().orc_group_adj(self.h, g, adj.ctypes.data_as(c_u8p), dp(dist))
        return adj, dist

    def group_factor_ids(self, g):
        n = self.groups()[g][2]
        ids = np.zeros(max(n, 1), dtype=np.int64)
        lib().orc_group_factor_ids(self.h, g, ids.ctypes.data_as(c_llp))
        return ids[:n]

    def group_inlier_ids(self, g):
        n = self.groups()[g][3]
        ids = np.zeros(max(n, 1), dtype=np.int64)
        lib().orc_group_inlier_ids(self.h, g, ids.ctypes.data_as(c_llp))
        return ids[:n]

    def frame_align_measurements(self, r0, ri):
        cap = max(self.num_lc(), 1)
        out = np.zeros((cap, psize(self.d)))
        n = lib().orc_frame_align(self.h, ord(r0), ord(ri), dp(out))
        return None if n < 0 else out[:n]

    def robot_odom_values(self, prefix, transform=None):
        n = lib().orc_robot_odom_values(self.h, ord(prefix), None, None, None)
        if n < 0:
            return None
        keys = np.zeros(max(n, 1), dtype=np.uint64); poses = np.zeros((max(n, 1), psize(self.d)))
        tp = dp(np.ascontiguousarray(transform, dtype=np.float64)) if transform is not None else None
        lib().orc_robot_odom_values(self.h, ord(prefix), tp, keys.ctypes.data_as(c_u64p), dp(poses))
        return keys[:n], poses[:n]

    def landmarks(self):
        """[(key, n_observations, n_inliers)] in first-seen order"""
        out = []
        for l in range(lib().orc_num_landmarks(self.h)):
            k, n, ni = C.c_uint64(), C.c_int(), C.c_int()
            lib().orc_landmark_info(self.h, l, C.byref(k), C.byref(n), C.byref(ni))
            out.append((k.value, n.value, ni.value))
        return out

    def landmark_adj(self, l):
        n = lib().orc_landmark_adj(self.h, l, None, None)
        adj = np.zeros((n, n), dtype=np.uint8); dist = np.zeros((n, n))
        lib().orc_landmark_adj(self.h, l, adj.ctypes.data_as(c_u8p), dp(dist))
        return adj, dist

    def landmark_ids(self, l):
        _, n, ni = self.landmarks()[l]
        f = np.zeros(max(n, 1), dtype=np.int64); i = np.zeros(max(ni, 1), dtype=np.int64)
        lib().orc_landmark_ids(self.h, l, f.ctypes.data_as(c_llp), i.ctypes.data_as(c_llp))
        return f[:n], i[:ni]

    def flagged(self):
        n = lib().orc_num_flagged(self.h)
        p = np.zeros((max(n, 1), 2), dtype=np.int64)
        lib().orc_flagged(self.h, p.ctypes.data_as(c_llp))
        return p[:n]

    def remove_last(self, c1=None, c2=None):
        k1, k2 = C.c_uint64(), C.c_uint64()
        if c1 is None:
            ok = lib().orc_remove_last_any(self.h, C.byref(k1), C.byref(k2))
        else:
            ok = lib().orc_remove_last(self.h, ord(c1), ord(c2), C.byref(k1), C.byref(k2))
        return (k1.value, k2.value) if ok else None

    def ignore_prefix(self, c):
        lib().orc_ignore_prefix(self.h, ord(c))

    def revive_prefix(self, c):
        lib().orc_revive_prefix(self.h, ord(c))

    def traj_get(self, key):
        n = ndim(self.d)
        pose = np.zeros(psize(self.d)); cov = np.zeros((n, n)); node = C.c_int(); rot = C.c_int()
        ok = lib().orc_traj_get(self.h, key, dp(pose), dp(cov), C.byref(node), C.byref(rot))
        return (pose, cov, node.value, rot.value) if ok else None
