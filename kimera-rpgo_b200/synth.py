"""Synthetic pose graphs of the shapes BASELINE.json names.

Everything is generated on the host with numpy from a seed; outputs are the factor tuples the
OutlierRemoval interface consumes: (type, key1, key2, pose, cov) and values (key, pose)."""
import numpy as np

BETWEEN = 0


def sym(ch, idx):
    return (ord(ch) << 56) | int(idx)


def _rodrigues(w):
    """batched exp map so(3) -> SO(3); w: (..., 3)"""
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    small = th < 1e-12
    th_s = np.where(small, 1.0, th)
    a = np.where(small, 1.0, np.sin(th_s) / th_s)
    b = np.where(small, 0.5, (1 - np.cos(th_s)) / (th_s * th_s))
    return np.eye(3) + a * K + b * (K @ K)


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _pack3(R, t):
    return np.concatenate([R.reshape(R.shape[:-2] + (9,)), t], axis=-1)


class Graph3D:
    """One robot's 3D trajectory: ground truth, noisy odometry factors and values."""

    def __init__(self, rng, prefix, P, sig_r2=1e-4, sig_t2=1e-3, origin=(0.0, 0.0, 0.0)):
        self.prefix, self.P = prefix, P
        tilt = _rodrigues(np.array([0.01, 0.0, 0.0]))
        c, s = np.cos(2 * np.pi / 50), np.sin(2 * np.pi / 50)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        dR_true = Rz @ tilt
        dt_true = np.array([1.0, 0.0, 0.02])
        # ground truth
        self.R = np.empty((P, 3, 3)); self.t = np.empty((P, 3))
        self.R[0] = np.eye(3); self.t[0] = np.asarray(origin)
        for k in range(1, P):
            self.R[k] = self.R[k - 1] @ dR_true
            self.t[k] = self.t[k - 1] + self.R[k - 1] @ dt_true
        # measured odometry = true delta (+) noise
        nr = rng.normal(scale=np.sqrt(sig_r2), size=(P - 1, 3))
        nt = rng.normal(scale=np.sqrt(sig_t2), size=(P - 1, 3))
        self.odo_R = dR_true @ _rodrigues(nr)
        self.odo_t = dt_true + nt
        self.odo_cov = np.diag([sig_r2] * 3 + [sig_t2] * 3)
        self.keys = np.array([sym(prefix, k) for k in range(P)], dtype=np.uint64)

    def odom_factors(self):
        poses = _pack3(self.odo_R, self.odo_t)
        return [(BETWEEN, int(self.keys[k]), int(self.keys[k + 1]), poses[k], self.odo_cov) for k in range(self.P - 1)]

    def values(self):
        # initial guess = dead-reckoned odometry (what a front end would provide)
        R = np.eye(3); t = self.t[0].copy()
        out = [(int(self.keys[0]), _pack3(R, t))]
        for k in range(self.P - 1):
            t = t + R @ self.odo_t[k]
            R = R @ self.odo_R[k]
            out.append((int(self.keys[k + 1]), _pack3(R, t)))
        return out


def loop_closures_3d(rng, ga, gb, n, outlier_frac, sig_r2=1e-3, sig_t2=1e-2, min_sep=20, mixed_direction=False):
    """n closures between trajectories ga and gb (ga is gb for intra-robot).  Returns factor tuples and the
    boolean outlier mask."""
    same = ga is gb
    i = rng.integers(0, ga.P, size=n)
    j = rng.integers(0, gb.P, size=n)
    if same:
        if ga.P <= 2 * min_sep + 1:
            raise ValueError("need P > 2*min_sep+1 poses to draw separated closures")
        bad = np.abs(i - j) <= min_sep
        while bad.any():
            j[bad] = rng.integers(0, gb.P, size=int(bad.sum()))
            bad = np.abs(i - j) <= min_sep
    Ri, ti, Rj, tj = ga.R[i], ga.t[i], gb.R[j], gb.t[j]
    Rij = np.transpose(Ri, (0, 2, 1)) @ Rj
    tij = np.einsum("nji,nj->ni", Ri, tj - ti)
    nr = rng.normal(scale=np.sqrt(sig_r2), size=(n, 3))
    nt = rng.normal(scale=np.sqrt(sig_t2), size=(n, 3))
    Rm = Rij @ _rodrigues(nr)
    tm = tij + nt
    out = rng.random(n) < outlier_frac
    k = int(out.sum())
    Rm[out] = _random_rotations(rng, k)
    tm[out] = rng.uniform(-10, 10, size=(k, 3))
    cov = np.diag([sig_r2] * 3 + [sig_t2] * 3)
    poses = _pack3(Rm, tm)
    kf, kt = ga.keys[i].copy(), gb.keys[j].copy()
    if mixed_direction and not same:
        # half of the closures are stated b -> a (exercises the Pcm.h:691-698 key swap)
        flip = rng.random(n) < 0.5
        Rf = np.transpose(Rm, (0, 2, 1))
        tf = -np.einsum("nij,nj->ni", Rf, tm)
        pf = _pack3(Rf, tf)
        poses = np.where(flip[:, None], pf, poses)
        kf2 = np.where(flip, kt, kf); kt2 = np.where(flip, kf, kt)
        kf, kt = kf2, kt2
    facs = [(BETWEEN, int(kf[q]), int(kt[q]), poses[q], cov) for q in range(n)]
    return facs, out


def config2(seed=1, P=2500, n=1000, outlier_frac=0.5):
    """3D single robot: P poses, n closures, 50 % outliers (BASELINE config 2; also config 5 shapes)."""
    rng = np.random.default_rng(seed)
    g = Graph3D(rng, 'a', P)
    lcs, out = loop_closures_3d(rng, g, g, n, outlier_frac)
    return dict(d=3, values=g.values(), odom=g.odom_factors(), lcs=lcs, outlier=out)


def config4(seed=3, robots=8, P=20000, n=50000, outlier_frac=0.3, skew=False):
    """8 robots x P poses, n closures spread over the <= 36 unordered prefix pairs (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    gs = [Graph3D(rng, chr(ord('a') + r), P, origin=(0.0, 5.0 * r, 0.0)) for r in range(robots)]
    pairs = [(a, b) for a in range(robots) for b in range(a, robots)]
    w = np.ones(len(pairs))
    if skew:
        w[1] = len(pairs)  # group (a,b) gets about half
    counts = rng.multinomial(n, w / w.sum())
    lcs, outl = [], []
    for (a, b), c in zip(pairs, counts):
        f, o = loop_closures_3d(rng, gs[a], gs[b], int(c), outlier_frac, mixed_direction=True)
        lcs += f
        outl.append(o)
    order = rng.permutation(len(lcs))
    lcs = [lcs[q] for q in order]
    outlier = np.concatenate(outl)[order]
    values, odom = [], []
    for g in gs:
        values += g.values()
        odom += g.odom_factors()
    return dict(d=3, values=values, odom=odom, lcs=lcs, outlier=outlier)


def config3(seed=2, P=10000, n=10000, outlier_frac=0.3):
    """2D Manhattan-style: unit steps, 90-degree turns w.p. 0.1 (BASELINE config 3).  pose = (c, s, x, y)."""
    rng = np.random.default_rng(seed)
    sig_xy2, sig_th2 = 1e-3, 1e-4
    turn = rng.random(P - 1) < 0.1
    dth_true = np.where(turn, np.where(rng.random(P - 1) < 0.5, np.pi / 2, -np.pi / 2), 0.0)
    th = np.zeros(P); xy = np.zeros((P, 2))
    for k in range(1, P):
        xy[k] = xy[k - 1] + np.array([np.cos(th[k - 1]), np.sin(th[k - 1])])
        th[k] = th[k - 1] + dth_true[k - 1]
    o_th = dth_true + rng.normal(scale=np.sqrt(sig_th2), size=P - 1)
    o_xy = np.array([1.0, 0.0]) + rng.normal(scale=np.sqrt(sig_xy2), size=(P - 1, 2))
    ocov = np.diag([sig_xy2, sig_xy2, sig_th2])
    keys = [sym('a', k) for k in range(P)]
    odom = [(BETWEEN, keys[k], keys[k + 1], np.array([np.cos(o_th[k]), np.sin(o_th[k]), o_xy[k, 0], o_xy[k, 1]]), ocov)
            for k in range(P - 1)]
    values = []
    cth, cxy = 0.0, np.zeros(2)
    values.append((keys[0], np.array([1.0, 0.0, 0.0, 0.0])))
    for k in range(P - 1):
        c, s = np.cos(cth), np.sin(cth)
        cxy = cxy + np.array([c * o_xy[k, 0] - s * o_xy[k, 1], s * o_xy[k, 0] + c * o_xy[k, 1]])
        cth = cth + o_th[k]
        values.append((keys[k + 1], np.array([np.cos(cth), np.sin(cth), cxy[0], cxy[1]])))
    if P <= 41:
        raise ValueError("need P > 41 poses to draw separated closures")
    i = rng.integers(0, P, size=n); j = rng.integers(0, P, size=n)
    bad = np.abs(i - j) <= 20
    while bad.any():
        j[bad] = rng.integers(0, P, size=int(bad.sum()))
        bad = np.abs(i - j) <= 20
    lsig_xy2, lsig_th2 = 1e-2, 1e-3
    dth = th[j] - th[i] + rng.normal(scale=np.sqrt(lsig_th2), size=n)
    d = xy[j] - xy[i]
    ci, si = np.cos(th[i]), np.sin(th[i])
    dx = ci * d[:, 0] + si * d[:, 1] + rng.normal(scale=np.sqrt(lsig_xy2), size=n)
    dy = -si * d[:, 0] + ci * d[:, 1] + rng.normal(scale=np.sqrt(lsig_xy2), size=n)
    out = rng.random(n) < outlier_frac
    k = int(out.sum())
    dth[out] = rng.uniform(-np.pi, np.pi, size=k)
    dx[out] = rng.uniform(-10, 10, size=k); dy[out] = rng.uniform(-10, 10, size=k)
    lcov = np.diag([lsig_xy2, lsig_xy2, lsig_th2])
    lcs = [(BETWEEN, keys[i[q]], keys[j[q]], np.array([np.cos(dth[q]), np.sin(dth[q]), dx[q], dy[q]]), lcov)
           for q in range(n)]
    return dict(d=2, values=values, odom=odom, lcs=lcs, outlier=out)


def as_arrays(gph):
    """Array form of a generated graph (what PcmGpu.odom_append_arrays / lc_append_arrays take)."""
    d = gph["d"]
    ps, nn = (12, 36) if d == 3 else (4, 9)
    vals = dict((int(k), p) for k, p in gph["values"])
    od, lc = gph["odom"], gph["lcs"]
    out = dict(d=d)
    out["o_prev"] = np.array([f[1] for f in od], dtype=np.uint64)
    out["o_new"] = np.array([f[2] for f in od], dtype=np.uint64)
    out["o_pose"] = np.ascontiguousarray(np.stack([f[3] for f in od])).reshape(len(od), ps)
    out["o_cov"] = np.ascontiguousarray(np.stack([np.asarray(f[4]).reshape(nn) for f in od]))
    out["o_init"] = np.ascontiguousarray(np.stack([vals[int(f[1])] for f in od]))
    out["l_from"] = np.array([f[1] for f in lc], dtype=np.uint64)
    out["l_to"] = np.array([f[2] for f in lc], dtype=np.uint64)
    out["l_pose"] = np.ascontiguousarray(np.stack([f[3] for f in lc])).reshape(len(lc), ps)
    out["l_cov"] = np.ascontiguousarray(np.stack([np.asarray(f[4]).reshape(nn) for f in lc]))
    out["v_keys"] = np.array([k for k, _ in gph["values"]], dtype=np.uint64)
    out["v_pose"] = np.ascontiguousarray(np.stack([p for _, p in gph["values"]])).reshape(len(gph["values"]), ps)
    return out
