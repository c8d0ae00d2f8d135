"""Host-side mirror of the reference's OutlierRemoval plug-in for the PCM path, on top of the C ABI.

`PcmGpu` keeps exactly what Pcm<poseT,T> keeps on the host — factor classification, the odometry /
special / per-group factor lists, ignored prefixes, output-graph assembly (reference
include/KimeraRPGO/outlier/Pcm.h:148-281, :299-408, :977-1005) — and sends the three arithmetic stages
to the GPU through include/rpgo_b200.h (odometry cache, pairwise consistency, max clique).
Factors are plain tuples (type, key1, key2, pose, cov) and values (key, pose), i.e. what a
gtsam::BetweenFactor / PriorFactor / Values entry carries; the C++ adapter in INTEGRATION.md maps
GTSAM objects onto the same calls.
"""
import ctypes as C

import numpy as np

from . import _capi

BETWEEN, PRIOR, OTHER = 0, 1, 2
MODE_PCM, MODE_SIMPLE = 0, 1
CLIQUE_HEU, CLIQUE_HEU_INCREMENTAL, CLIQUE_EXACT = 0, 1, 2
TRAJ_FOLD, TRAJ_SCAN = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TILED = 0, 1, 2


def key_chr(k):
    return (int(k) >> 56) & 0xFF


def _dp(a):
    return a.ctypes.data_as(_capi.c_dp)


class RpgoError(RuntimeError):
    pass


class IdList:
    """A list of factor ids / group ordinals whose bulk appends from the array entry points (numpy arrays, ranges) stay
    as chunks until a list operation needs the elements: a 50 000-closure batch then costs microseconds of host
    bookkeeping instead of building (and later freeing) 50 000 Python integers per list."""
    __slots__ = ("_l", "_chunks", "_n")

    def __init__(self, it=()):
        self._l = list(it)
        self._chunks = []
        self._n = len(self._l)

    def _flush(self):
        if self._chunks:
            for c in self._chunks:
                self._l.extend(c.tolist() if isinstance(c, np.ndarray) else c)
            self._chunks = []
        return self._l

    def extend(self, it):
        if isinstance(it, (np.ndarray, range)):
            if len(it):
                self._chunks.append(it)
                self._n += len(it)
        else:
            self._flush().extend(it)
            self._n = len(self._l)

    def append(self, x):
        self._flush().append(x)
        self._n += 1

    def pop(self):
        v = self._flush().pop()
        self._n -= 1
        return v

    def index(self, x):
        return self._flush().index(x)

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(self._flush())

    def __getitem__(self, i):
        if self._chunks and isinstance(i, (int, np.integer)):
            j = int(i) + self._n if i < 0 else int(i)
            if j < len(self._l):
                return self._l[j]
            j -= len(self._l)
            for c in self._chunks:
                if j < len(c):
                    return int(c[j])
                j -= len(c)
            raise IndexError(i)
        return self._flush()[i]

    def __eq__(self, other):
        return list(self) == list(other)

    def __array__(self, dtype=None, copy=None):
        parts = [np.asarray(self._l, dtype=np.int64)] + [np.asarray(c, dtype=np.int64) for c in self._chunks]
        a = np.concatenate(parts) if len(parts) > 1 else parts[0]
        return a if dtype is None else a.astype(dtype, copy=False)

    def __repr__(self):
        return "IdList(%r)" % (self._flush(),)


class PcmGpu:
    """OutlierRemoval interface (OutlierRemoval.h:19-102) for Pcm2D/Pcm3D/PcmSimple2D/PcmSimple3D."""

    def __init__(self, d=3, mode=MODE_PCM, odom_threshold=10.0, lc_threshold=5.0, odom_trans=0.05, odom_rot=0.005,
                 dist_trans=0.01, dist_rot=0.001, incremental=False, device=-1, traj_mode=TRAJ_FOLD,
                 kernel=KERNEL_AUTO, rank=0, world=1, special_symbols=(), scan_chunk=64, comm_id=None):
        self.lib = _capi.load()
        cfg = _capi.RpgoCfg()
        self.lib.rpgo_default_cfg(C.byref(cfg))
        cfg.dim, cfg.mode = d, mode
        cfg.odom_threshold, cfg.lc_threshold = odom_threshold, lc_threshold
        cfg.odom_trans_threshold, cfg.odom_rot_threshold = odom_trans, odom_rot
        cfg.dist_trans_threshold, cfg.dist_rot_threshold = dist_trans, dist_rot
        cfg.incremental = int(incremental)
        cfg.device, cfg.traj_mode, cfg.kernel, cfg.rank, cfg.world = device, traj_mode, kernel, rank, world
        cfg.scan_chunk = scan_chunk
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.rpgo_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RpgoError("rpgo_create failed with status %d (no usable CUDA device? there is no CPU fallback)" % rc)
        self.d, self.mode = d, mode
        self._exchange_cb = None
        self.has_comm = False
        if world > 1 and comm_id is not None:
            self.comm_init(comm_id)
        self.ps = 12 if d == 3 else 4
        self.n = 6 if d == 3 else 3
        self.incremental = bool(incremental)
        # Pcm.h:74-82
        self.odom_check = not (odom_threshold < 0 or odom_rot < 0 or odom_trans < 0)
        self.loop_check = not (lc_threshold < 0 or dist_rot < 0 or dist_trans < 0)
        self.special_symbols = set(ord(c) if isinstance(c, str) else int(c) for c in special_symbols)
        self._clear_host_state()

    def reset(self):
        """A fresh Pcm object on the same handle (rpgo_reset keeps stream, arena, pinned staging, communicator)."""
        self._check(self.lib.rpgo_reset(self.h), "rpgo_reset")
        self._clear_host_state()

    def _clear_host_state(self):
        self.values = {}
        self.nfg_odom, self.nfg_special = IdList(), []
        self.special_is_prior = {}     # factor id -> prior key (for removePriorFactorsWithPrefix)
        self.group_factors = {}        # group ordinal -> [factor ids]
        self.group_consistent = {}     # group ordinal -> [factor ids]
        self.group_order = []
        self.landmark_group = {}       # landmark key -> group ordinal
        self.landmark_order = []
        self.lc_in_order = IdList()
        self.ignored = []
        self.total_lc = 0
        self.total_good_lc = 0
        self.next_id = 0
        self.output = []

    # ---- logging in the reference's on-disk formats (OutlierRemoval.h:54-62, Pcm.h:1136-1164,
    #      RobustSolver.cpp:93-102, :359-367) -------------------------------------------------------
    def log_output(self, folder):
        import os
        self.log_folder = folder
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "outlier_rejection_status.txt"), "w") as f:
            f.write("total inliers spin-time mc-time\n")
        with open(os.path.join(folder, "rpgo_status.csv"), "w") as f:
            f.write("graph-size,spin-time(mu-s),num-lc,num-inliers\n")

    def _log_spin(self, spin_s, clique_s):
        import os
        folder = getattr(self, "log_folder", None)
        if not folder:
            return
        for g in self.group_order:  # saveAdjacencyMatrix: "<id1>-<id2>_adj_matrix.txt", dense rows
            a, b, n = self.group_info(g)
            if 0 < n <= 4096 and self.loop_check:
                adj, _ = self.group_adj(g, with_dist=False)
                # keys without a symbol character have prefix '\0' (e.g. plain g2o ids): the reference's file name would
                # then contain a NUL and the file is never created; the ordinal is written instead
                name = lambda c: c if c.isprintable() and c not in "/\\" else str(ord(c))
                np.savetxt(os.path.join(folder, "%s-%s_adj_matrix.txt" % (name(a), name(b))), adj, fmt="%d")
        with open(os.path.join(folder, "outlier_rejection_status.txt"), "a") as f:  # logSpinStatus
            f.write("%d %d %d %d\n" % (self.total_lc, self.total_good_lc, int(spin_s * 1e3), int(clique_s * 1e3)))
        with open(os.path.join(folder, "rpgo_status.csv"), "a") as f:  # RobustSolver::update
            f.write("%d,%d,%d,%d\n" % (len(self.output), int(spin_s * 1e6), self.total_lc, self.total_good_lc))

    # ---- multi-GPU data plane (rpgo_comm_*: NCCL behind the C ABI) ------------------------------------
    @staticmethod
    def comm_unique_id():
        """128 opaque bytes from rank 0 (ncclGetUniqueId); hand them to every rank, then call comm_init everywhere."""
        buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
        rc = _capi.load().rpgo_comm_unique_id(buf)
        if rc != 0:
            raise RpgoError("rpgo_comm_unique_id failed (%d): NCCL not available?" % rc)
        return buf.raw

    def comm_init(self, comm_id):
        """Collective: every rank of the world calls this with rank 0's id."""
        assert len(comm_id) == _capi.COMM_ID_BYTES
        buf = C.create_string_buffer(bytes(comm_id), _capi.COMM_ID_BYTES)
        self._check(self.lib.rpgo_comm_init(self.h, C.cast(buf, C.c_void_p), self.cfg.rank, self.cfg.world), "rpgo_comm_init")
        self.has_comm = True

    def set_exchange(self, cb):
        """Register (or clear, cb=None) the collective the sharded clique searches call; see rpgo_set_exchange."""
        self._exchange_cb = cb  # keep the ctypes thunk alive
        self._check(self.lib.rpgo_set_exchange(self.h, C.cast(cb, C.c_void_p) if cb is not None else None, None),
                    "rpgo_set_exchange")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.rpgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RpgoError("%s failed (%d): %s" % (what, rc, self.lib.rpgo_last_error(self.h).decode()))

    # ---- OutlierRemoval::removeOutliers (Pcm.h:148-281) ---------------------------------------
    def update(self, factors, values):
        """One removeOutliers() call.  Returns do_optimize."""
        import time
        t_start = time.perf_counter()
        t_clique = 0.0
        new_keys = set()
        for k, p in values:
            self.values[int(k)] = np.asarray(p, dtype=np.float64)
            new_keys.add(int(k))
        if len(factors) == 0:
            return False
        do_optimize = False
        odom, lcs = [], []
        for f in factors:
            fid = self.next_id
            self.next_id += 1
            ftype, k1, k2 = f[0], int(f[1]), int(f[2])
            if ftype == BETWEEN:
                if key_chr(k1) in self.special_symbols or key_chr(k2) in self.special_symbols:
                    # landmark observation, Pcm.h:180-188
                    if k1 in new_keys or k2 in new_keys:
                        self._landmark_first(fid, k1, k2, f[3], f[4])            # FIRST_LANDMARK_OBSERVATION :207-220
                    elif k1 != k2:
                        lcs.append((fid, k1, k2, f[3], f[4]))                    # re-observation: a loop closure
                elif k1 + 1 == k2 and k2 in new_keys:
                    odom.append((fid, k1, k2, f[3], f[4]))
                elif k1 != k2:
                    lcs.append((fid, k1, k2, f[3], f[4]))
            else:
                self.nfg_special.append(fid)
                if ftype == PRIOR:
                    self.special_is_prior[fid] = k1
                do_optimize = True
        # the reference folds each odometry factor as it meets it and processes all loop closures
        # after the loop (Pcm.h:203-206 vs :242-246), so batching odometry first is order-preserving
        if odom:
            self._odom_append(odom)
        if lcs:
            is_lm = [key_chr(l[1]) in self.special_symbols or key_chr(l[2]) in self.special_symbols for l in lcs]
            lm = [l for l, m in zip(lcs, is_lm) if m]
            if lm:
                lcs = [l for l, m in zip(lcs, is_lm) if not m]
                self._landmark_reobserve(lm)
            num_new = self._lc_append(lcs) if lcs else {}
            t0 = time.perf_counter()
            if self.incremental:
                self._find_inliers_incremental(num_new)
            else:
                self._find_inliers()
            t_clique = time.perf_counter() - t0
            do_optimize = True
        self._build_graph()
        if getattr(self, "log_folder", None):
            self._log_spin(time.perf_counter() - t_start, t_clique)
        return do_optimize

    def _odom_append(self, odom):
        prev = np.array([o[1] for o in odom], dtype=np.uint64)
        new = np.array([o[2] for o in odom], dtype=np.uint64)
        dpose = np.stack([np.asarray(o[3], dtype=np.float64) for o in odom])
        dcov = np.stack([np.asarray(o[4], dtype=np.float64).reshape(self.n * self.n) for o in odom])
        ident = np.zeros(self.ps)
        if self.d == 3:
            ident[[0, 4, 8]] = 1.0
        else:
            ident[0] = 1.0
        init = np.stack([self.values.get(int(o[1]), ident) for o in odom])
        self.odom_append_arrays(prev, new, dpose, dcov, init, ids=[o[0] for o in odom])

    def odom_append_arrays(self, prev, new, dpose, dcov, init=None, ids=None):
        """Array form of the odometry stage (K1): n factors prev[i] -> new[i] with measured pose dpose[i]
        (n x 12|4) and covariance dcov[i] (n x 36|9); init[i] = values.at(prev[i]) (used for new prefixes)."""
        n = len(prev)
        prev = np.ascontiguousarray(prev, dtype=np.uint64)
        new = np.ascontiguousarray(new, dtype=np.uint64)
        dpose = np.ascontiguousarray(dpose, dtype=np.float64).reshape(n, self.ps)
        dcov = np.ascontiguousarray(dcov, dtype=np.float64).reshape(n, self.n * self.n)
        ip = None
        if init is not None:
            init = np.ascontiguousarray(init, dtype=np.float64).reshape(n, self.ps)
            ip = _dp(init)
        if ids is None:
            ids = range(self.next_id, self.next_id + n)
            self.next_id += n
        self.nfg_odom.extend(ids)
        self._check(self.lib.rpgo_odom_append(self.h, n, prev.ctypes.data_as(_capi.c_u64p), new.ctypes.data_as(_capi.c_u64p),
                                              _dp(dpose), _dp(dcov), ip), "rpgo_odom_append")
        return (dpose.nbytes + dcov.nbytes + (init.nbytes if init is not None else 0))

    def _lc_append(self, lcs):
        # Pcm.h:431-435: both keys must exist in the values
        lcs = [l for l in lcs if l[1] in self.values and l[2] in self.values]
        if not lcs:
            return {}
        kf = np.array([l[1] for l in lcs], dtype=np.uint64)
        kt = np.array([l[2] for l in lcs], dtype=np.uint64)
        pose = np.stack([np.asarray(l[3], dtype=np.float64) for l in lcs])
        cov = np.stack([np.asarray(l[4], dtype=np.float64).reshape(self.n * self.n) for l in lcs])
        num_new, _ = self.lc_append_arrays(kf, kt, pose, cov, ids=[l[0] for l in lcs])
        return num_new

    def lc_append_arrays(self, kf, kt, pose, cov, ids=None):
        """Array form of the loop-closure stage (K2 + K3).  Returns ({group: number of new closures}, accepted)."""
        n = len(kf)
        kf = np.ascontiguousarray(kf, dtype=np.uint64)
        kt = np.ascontiguousarray(kt, dtype=np.uint64)
        pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(n, self.ps)
        cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(n, self.n * self.n)
        acc = np.zeros(n, dtype=np.uint8)
        grp = np.zeros(n, dtype=np.int32)
        idx = np.zeros(n, dtype=np.int32)
        self._check(self.lib.rpgo_lc_append(self.h, n, kf.ctypes.data_as(_capi.c_u64p), kt.ctypes.data_as(_capi.c_u64p),
                                            _dp(pose), _dp(cov), acc.ctypes.data_as(_capi.c_u8p),
                                            grp.ctypes.data_as(_capi.c_i32p), idx.ctypes.data_as(_capi.c_i32p), None),
                    "rpgo_lc_append")
        if ids is None:
            ids = range(self.next_id, self.next_id + n)
            self.next_id += n
        num_new = {}
        if n > 0 and acc.all() and (grp == grp[0]).all():
            # the common batch: everything accepted into one group -- no per-element work on the host
            g = int(grp[0])
            if g not in self.group_factors:
                self.group_factors[g] = IdList()
                self.group_consistent[g] = []
                self.group_order.append(g)
            assert len(self.group_factors[g]) == int(idx[0])
            self.group_factors[g].extend(ids if isinstance(ids, (range, np.ndarray)) else list(ids))
            num_new[g] = n
            self.lc_in_order.extend(grp)
            self.total_lc += n
        else:
            ids = np.asarray(ids)
            ok = acc.astype(bool)
            touched = np.flatnonzero(np.bincount(grp[ok])) if ok.any() else []  # (np.unique's first call costs ~50 ms)
            for g in touched:
                g = int(g)
                sel = ok & (grp == g)
                if g not in self.group_factors:
                    self.group_factors[g] = IdList()
                    self.group_consistent[g] = []
                    self.group_order.append(g)
                assert len(self.group_factors[g]) == int(idx[sel][0])
                self.group_factors[g].extend(ids[sel])
                num_new[g] = int(sel.sum())
            self.lc_in_order.extend(grp[ok])
            self.total_lc += int(ok.sum())
        self.last_h2d_bytes = pose.nbytes + cov.nbytes + 9 * n
        self.last_d2h_bytes = n
        self.last_group, self.last_index = grp, idx   # per input closure: group ordinal / position in it (-1: rejected)
        # multi-GPU: with a communicator (comm_init) the library has all-gathered the row chunks and rebuilt mirror +
        # degrees on every rank before rpgo_lc_append returned
        return num_new, acc

    # ---- landmarks (Pcm.h:207-220, :437-455, :775-844) -----------------------------------------
    def _landmark_key(self, k1, k2):
        return k1 if key_chr(k1) in self.special_symbols else k2

    def _landmark_call(self, lkey, obs, reset):
        n = len(obs)
        pk = np.array([o[1] for o in obs], dtype=np.uint64)
        pose = np.ascontiguousarray(np.stack([np.asarray(o[2], dtype=np.float64) for o in obs])).reshape(n, self.ps)
        cov = np.ascontiguousarray(np.stack([np.asarray(o[3], dtype=np.float64).reshape(self.n * self.n) for o in obs]))
        g = C.c_int32(-1)
        self._check(self.lib.rpgo_landmark_append(self.h, int(lkey), n, pk.ctypes.data_as(_capi.c_u64p), _dp(pose), _dp(cov),
                                                  int(reset), C.byref(g)), "rpgo_landmark_append")
        return g.value

    def _landmark_first(self, fid, k1, k2, pose, cov):
        lkey = self._landmark_key(k1, k2)
        if k1 == lkey:
            # the reference stores it, then trips over it at the first re-observation (Pcm.h:803-808 returns
            # before saving the grown matrices; the next growth copies a wrongly sized block)
            # Not an error of update(): skipped with a warning, as the reference only warns about observations it cannot
            # use, and nothing of this call has been applied yet.
            import warnings
            warnings.warn("rpgo: landmark observation %d is stated landmark -> pose; skipped (observations must be pose -> landmark)" % fid)
            return
        g = self._landmark_call(lkey, [(fid, k1, pose, cov)], reset=True)
        if lkey not in self.landmark_group:
            self.landmark_order.append(lkey)
        self.landmark_group[lkey] = g
        self.group_factors[g] = [fid]
        self.group_consistent[g] = [fid]
        self.total_lc += 1

    def _landmark_reobserve(self, lm):
        by_key = {}
        for fid, k1, k2, pose, cov in lm:
            if k1 not in self.values or k2 not in self.values:
                continue                                           # Pcm.h:431-435
            lkey = self._landmark_key(k1, k2)
            if k1 == lkey:
                continue                                           # malformed: see _landmark_first
            by_key.setdefault(lkey, []).append((fid, k1, pose, cov))
        for lkey, obs in by_key.items():
            g = self._landmark_call(lkey, obs, reset=False)
            if lkey not in self.landmark_group:
                self.landmark_order.append(lkey)
                self.landmark_group[lkey] = g
                self.group_factors[g] = []
                self.group_consistent[g] = []
            self.group_factors[g].extend(o[0] for o in obs)
            self.total_lc += len(obs)

    def _landmark_inliers(self):  # Pcm.h:878-895
        for lkey in self.landmark_order:
            g = self.landmark_group[lkey]
            fs = self.group_factors[g]
            k, ids, _ = self.find_inliers_raw(g, CLIQUE_HEU)
            self.group_consistent[g] = [fs[i] for i in ids[:k]]
            self.total_good_lc += k

    def landmarks(self):
        """[(key, n_observations, n_inliers)] in first-seen order"""
        return [(k, len(self.group_factors[self.landmark_group[k]]), len(self.group_consistent[self.landmark_group[k]]))
                for k in self.landmark_order]

    def find_inliers_raw(self, g, clique_mode=CLIQUE_HEU, n_new=0, prev_size=0):
        """(size, ids, true_clique) straight from rpgo_find_inliers."""
        n = len(self.group_factors[g])
        ids = np.zeros(max(n, 1), dtype=np.int32)
        true = np.zeros(max(n, 1), dtype=np.int32)
        size = C.c_int64(0)
        self._check(self.lib.rpgo_find_inliers(self.h, g, clique_mode, n_new, prev_size, ids.ctypes.data_as(_capi.c_i32p),
                                               C.byref(size), true.ctypes.data_as(_capi.c_i32p)), "rpgo_find_inliers")
        k = int(size.value)
        return k, ids[:max(k, 0)].copy(), true[:max(k, 0)].copy()

    def find_inliers_batch(self, groups, clique_mode=CLIQUE_HEU, n_new=None, prev_size=None):
        """rpgo_find_inliers_batch: [(size, ids)] for the given group ordinals, searched concurrently."""
        groups = [int(g) for g in groups]
        if not groups:
            return []
        caps = np.array([max(len(self.group_factors[g]), 1) for g in groups], dtype=np.int64)
        off = np.zeros(len(groups), dtype=np.int64)
        off[1:] = np.cumsum(caps)[:-1]
        ids = np.zeros(int(caps.sum()), dtype=np.int32)
        sizes = np.zeros(len(groups), dtype=np.int64)
        garr = np.array(groups, dtype=np.int32)
        nn = None if n_new is None else np.ascontiguousarray(n_new, dtype=np.int64)
        pv = None if prev_size is None else np.ascontiguousarray(prev_size, dtype=np.int64)
        self._check(self.lib.rpgo_find_inliers_batch(
            self.h, len(groups), garr.ctypes.data_as(_capi.c_i32p), clique_mode,
            None if nn is None else nn.ctypes.data_as(_capi.c_i64p), None if pv is None else pv.ctypes.data_as(_capi.c_i64p),
            ids.ctypes.data_as(_capi.c_i32p), off.ctypes.data_as(_capi.c_i64p), sizes.ctypes.data_as(_capi.c_i64p)),
            "rpgo_find_inliers_batch")
        return [(int(sizes[k]), ids[off[k]:off[k] + max(int(sizes[k]), 0)].copy()) for k in range(len(groups))]

    def _find_inliers(self):  # Pcm.h:851-899
        self.total_good_lc = 0
        todo = []
        for g in self.group_order:
            fs = self.group_factors[g]
            if self.loop_check:
                if len(fs) == 0:
                    self.group_consistent[g] = []
                    continue
                todo.append(g)
            else:
                self.group_consistent[g] = list(fs)
        # the groups are independent: one batched call (rpgo_find_inliers_batch) searches them concurrently
        for g, (k, ids) in zip(todo, self.find_inliers_batch(todo, CLIQUE_HEU)):
            fs = self.group_factors[g]
            self.group_consistent[g] = [fs[i] for i in ids[:k]]
        self.total_good_lc = sum(len(self.group_consistent[g]) for g in self.group_order)
        self._landmark_inliers()

    def _find_inliers_incremental(self, num_new):  # Pcm.h:906-970
        if not self.loop_check:
            # the reference still calls findMaxCliqueHeuIncremental, on the 1x1 zero matrix the disabled check leaves behind
            # (Pcm.h:484-486): the only candidate is vertex 0, and only when exactly one closure is new and nothing was
            # selected before (findCliqueHeu.cpp:141-145).  Nothing to compute on the GPU.
            for g, nn in num_new.items():
                if nn == 1 and len(self.group_consistent[g]) == 0:
                    self.group_consistent[g] = [self.group_factors[g][0]]
            self.total_good_lc = sum(len(self.group_consistent[g]) for g in self.group_order)
            self._landmark_inliers()
            return
        gs = list(num_new.keys())
        prevs = [len(self.group_consistent[g]) for g in gs]
        res = self.find_inliers_batch(gs, CLIQUE_HEU_INCREMENTAL, [num_new[g] for g in gs], prevs)
        for g, (k, ids) in zip(gs, res):
            fs = self.group_factors[g]
            if k > 0:
                self.group_consistent[g] = [fs[i] for i in ids[:k]]
        self.total_good_lc = sum(len(self.group_consistent[g]) for g in self.group_order)
        self._landmark_inliers()

    def _group_ids(self, g):
        a, b, _ = self.group_info(g)
        return a, b

    def _build_graph(self):  # Pcm.h:977-1005
        out = list(self.nfg_odom) + list(self.nfg_special)
        for g in self.group_order:
            a, b = self._group_ids(g)
            if ord(a) in self.ignored or ord(b) in self.ignored:
                continue
            out.extend(self.group_consistent[g])
        for lkey in self.landmark_order:  # Pcm.h:996-1002
            out.extend(self.group_consistent[self.landmark_group[lkey]])
        self.output = out

    # ---- the rest of the OutlierRemoval interface ----------------------------------------------
    def remove_last(self, c1=None, c2=None):  # Pcm.h:299-353
        if c1 is None:
            if not self.lc_in_order:
                return None
            g = self.lc_in_order.pop()
        else:
            g = self.lib.rpgo_find_group(self.h, ord(c1), ord(c2))
            if g < 0 or g not in self.group_factors:
                return None
        fs = self.group_factors[g]
        if len(fs) == 0:
            return None
        k1, k2 = C.c_uint64(), C.c_uint64()
        self._check(self.lib.rpgo_lc_remove_last(self.h, g, C.byref(k1), C.byref(k2)), "rpgo_lc_remove_last")
        removed = fs.pop()
        if len(fs) < 2:
            self.group_consistent[g] = list(fs)
        elif not self.loop_check:
            # no adjacency is kept when the pairwise check is disabled; the reference's own path is undefined behaviour
            # there (0x0 block into findMaxCliqueHeu).  Defined as findInliers' rule for that configuration
            # (Pcm.h:870-873: all factors) / "previous set minus the removed factor" in incremental mode.
            if self.incremental:
                self.group_consistent[g] = [f for f in self.group_consistent[g] if f != removed]
            else:
                self.group_consistent[g] = list(fs)
        else:
            k, ids, _ = self.find_inliers_raw(g, CLIQUE_HEU)
            self.group_consistent[g] = [fs[i] for i in ids[:k]]
        self._build_graph()
        return (k1.value, k2.value)

    def ignore_prefix(self, c):  # Pcm.h:357-365
        if ord(c) not in self.ignored:
            self.ignored.append(ord(c))
        self._build_graph()

    def revive_prefix(self, c):  # Pcm.h:369-377
        self.ignored = [x for x in self.ignored if x != ord(c)]
        self._build_graph()

    def get_ignored_prefixes(self):
        return [chr(x) for x in self.ignored]

    def remove_prior_factors_with_prefix(self, c):  # Pcm.h:387-408
        self.nfg_special = [f for f in self.nfg_special
                            if not (f in self.special_is_prior and key_chr(self.special_is_prior[f]) == ord(c))]
        self._build_graph()

    # ---- N4: multi-robot frame alignment, front half (Pcm.h:1024-1082) ------------------------------
    def frame_align_measurements(self, r0, ri):
        """T_w0_wi for every inlier closure of group (r0, ri); None when the group / a trajectory key is missing
        (the reference logs a warning and skips the robot)."""
        g = self.lib.rpgo_find_group(self.h, ord(r0), ord(ri))
        if g < 0 or g not in self.group_factors:
            return None
        fs = self.group_factors[g]
        idx = np.array([fs.index(f) for f in self.group_consistent[g]], dtype=np.int32)
        out = np.zeros((max(len(idx), 1), self.ps))
        rc = self.lib.rpgo_frame_align_measurements(self.h, g, ord(r0), len(idx), idx.ctypes.data_as(_capi.c_i32p), _dp(out))
        if rc == 4:
            return None
        self._check(rc, "rpgo_frame_align_measurements")
        return out[:len(idx)]

    def robot_odom_values(self, prefix, transform=None):
        """getRobotOdomValues: (keys, transform . pose) for the trajectory of `prefix`, ascending keys."""
        n = C.c_int64()
        self._check(self.lib.rpgo_robot_odom_values(self.h, ord(prefix), None, 0, None, None, C.byref(n)), "rpgo_robot_odom_values")
        keys = np.zeros(max(n.value, 1), dtype=np.uint64)
        poses = np.zeros((max(n.value, 1), self.ps))
        tp = None
        if transform is not None:
            transform = np.ascontiguousarray(transform, dtype=np.float64)
            tp = _dp(transform)
        self._check(self.lib.rpgo_robot_odom_values(self.h, ord(prefix), tp, n.value, keys.ctypes.data_as(_capi.c_u64p),
                                                    _dp(poses), C.byref(n)), "rpgo_robot_odom_values")
        return keys[:n.value], poses[:n.value]

    # counters: OutlierRemoval.h:24-27
    def num_lc(self):
        return self.total_lc

    def num_inliers(self):
        return self.total_good_lc

    def num_odom(self):
        return len(self.nfg_odom)

    def num_special(self):
        return len(self.nfg_special)

    def nfg_size(self):
        return len(self.output)

    def output_ids(self):
        return np.array(self.output, dtype=np.int64)

    def num_values(self):
        return len(self.values)

    # ---- inspection (parity) ---------------------------------------------------------------------
    def group_info(self, g):
        a, b, n = C.c_uint8(), C.c_uint8(), C.c_int64()
        self._check(self.lib.rpgo_group_info(self.h, g, C.byref(a), C.byref(b), C.byref(n)), "rpgo_group_info")
        return chr(a.value), chr(b.value), n.value

    def groups(self):
        out = []
        lm = set(self.landmark_group.values())
        for g in range(self.lib.rpgo_num_groups(self.h)):
            if g in lm:
                continue  # landmark groups are listed by landmarks()
            a, b, n = self.group_info(g)
            out.append((a, b, n, len(self.group_consistent.get(g, []))))
        return out

    def group_adj(self, g, with_dist=True):
        _, _, n = self.group_info(g)
        sw = max((n + 63) // 64, 1)
        rows = np.zeros((max(n, 1), sw), dtype=np.uint64)
        self._check(self.lib.rpgo_adj_bits(self.h, g, rows.ctypes.data_as(_capi.c_u64p), sw), "rpgo_adj_bits")
        adj = np.unpackbits(rows[:n].view(np.uint8), axis=1, bitorder="little")[:, :n]
        dist = None
        if with_dist:
            dist = np.zeros((n, n))
            if n > 0:
                self._check(self.lib.rpgo_pair_distances(self.h, g, _dp(dist)), "rpgo_pair_distances")
        return adj, dist

    def group_bits(self, g):
        """raw bitset rows (n x ceil(n/64) uint64)"""
        _, _, n = self.group_info(g)
        sw = max((n + 63) // 64, 1)
        rows = np.zeros((max(n, 1), sw), dtype=np.uint64)
        self._check(self.lib.rpgo_adj_bits(self.h, g, rows.ctypes.data_as(_capi.c_u64p), sw), "rpgo_adj_bits")
        return rows[:n]

    def group_factor_ids(self, g):
        return np.array(self.group_factors[g], dtype=np.int64)

    def group_inlier_ids(self, g):
        return np.array(self.group_consistent[g], dtype=np.int64)

    def degrees(self, g):
        _, _, n = self.group_info(g)
        deg = np.zeros(max(n, 1), dtype=np.int32)
        self._check(self.lib.rpgo_degrees(self.h, g, deg.ctypes.data_as(_capi.c_i32p)), "rpgo_degrees")
        return deg[:n]

    def flagged(self, g, cap=4096):
        pairs = np.zeros((cap, 2), dtype=np.int32)
        n = C.c_int64()
        self._check(self.lib.rpgo_near_threshold(self.h, g, pairs.ctypes.data_as(_capi.c_i32p), cap, C.byref(n)),
                    "rpgo_near_threshold")
        return n.value, pairs[:min(n.value, cap)]

    def traj_get(self, key):
        pose = np.zeros(self.ps)
        cov = np.zeros((self.n, self.n))
        node, rot = C.c_int32(), C.c_int32()
        rc = self.lib.rpgo_traj_get(self.h, int(key), _dp(pose), _dp(cov), C.byref(node), C.byref(rot))
        if rc != 0:
            return None
        return pose, cov, node.value, rot.value

    def load_adjacency(self, adj, c1='y', c2='z'):
        """test/benchmark hook: install a dense 0/1 adjacency as a group; returns the group ordinal."""
        adj = np.ascontiguousarray(adj, dtype=np.uint8)
        n = adj.shape[0]
        sw = max((n + 63) // 64, 1)
        packed = np.zeros((max(n, 1), sw * 8), dtype=np.uint8)
        if n:
            pb = np.packbits(adj, axis=1, bitorder="little")
            packed[:n, :pb.shape[1]] = pb
        rows = packed.view(np.uint64)
        g = C.c_int32(-1)
        self._check(self.lib.rpgo_debug_load_group(self.h, ord(c1), ord(c2), n, rows.ctypes.data_as(_capi.c_u64p), sw,
                                                   C.byref(g)), "rpgo_debug_load_group")
        self.group_factors[g.value] = list(range(n))
        self.group_consistent.setdefault(g.value, [])
        if g.value not in self.group_order:
            self.group_order.append(g.value)
        return g.value

    def launch_count(self):
        return int(self.lib.rpgo_launch_count(self.h))

    def sync(self):
        self._check(self.lib.rpgo_sync(self.h), "rpgo_sync")

    def stream_ptr(self):
        return self.lib.rpgo_stream(self.h)

    def pairwise_only(self, g, j_begin=0):
        self._check(self.lib.rpgo_group_pairwise(self.h, g, j_begin), "rpgo_group_pairwise")

    def finalize(self, g):
        self._check(self.lib.rpgo_group_finalize(self.h, g), "rpgo_group_finalize")

    def clique_stats(self):
        """statistics of the last heuristic search (rpgo_clique_stats)"""
        a, b, e = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self.lib.rpgo_clique_stats(self.h, C.byref(a), C.byref(b), C.byref(e)), "rpgo_clique_stats")
        return dict(row_ands=a.value, chains=b.value, epochs=e.value)

    def debug_pass(self, g, which):
        """one bitset pass alone (0 = mirror, 1 = degrees): bandwidth measurements"""
        self._check(self.lib.rpgo_debug_pass(self.h, g, which), "rpgo_debug_pass")

    def allgather(self, g):
        """all-gather of group g's row chunks over the handle's communicator + mirror + degrees"""
        self._check(self.lib.rpgo_group_allgather(self.h, g), "rpgo_group_allgather")

    def adj_bits_device(self, g):
        ptr, sw, n = C.c_void_p(), C.c_int64(), C.c_int64()
        self._check(self.lib.rpgo_adj_bits_device(self.h, g, C.byref(ptr), C.byref(sw), C.byref(n)), "rpgo_adj_bits_device")
        return ptr.value, sw.value, n.value

    def group_chunking(self, g):
        c, p = C.c_int64(), C.c_int64()
        self._check(self.lib.rpgo_group_chunking(self.h, g, C.byref(c), C.byref(p)), "rpgo_group_chunking")
        return c.value, p.value

    def recompute(self, g, j_begin=0):
        self._check(self.lib.rpgo_group_recompute(self.h, g, j_begin), "rpgo_group_recompute")
