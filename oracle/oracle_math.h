/* oracle_math.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Literal, dense restatement of the arithmetic on Kimera-RPGO's PCM hot path:
 *   PoseWithCovariance / PoseWithNode      include/KimeraRPGO/utils/GeometryUtils.h:56-289
 *   (GTSAM Pose3/Pose2/Rot3/Rot2 Lie ops, Eigen LLT / PartialPivLU restated from their
 *    published algorithms — GTSAM and Eigen are not vendored in /root/reference and are
 *    not installed; pinned only by README.md:32-50 "commit 686e16aa...".)
 *
 * Deliberately *different in structure* from the product's math header
 * (kimera-rpgo_b200/csrc/rpgo_math.cuh): everything here is a runtime-sized dense
 * n x n loop with full 6x6 / 3x3 adjoint matrices, identity Jacobians multiplied out,
 * no block-structure shortcuts.  The product exploits structure; tests demand the two
 * agree bit-for-bit, which pins the summation order of the product.
 *
 * Numerical contract (shared with the product, stated in DESIGN.md):
 *   - matrix / dot products accumulate in k-order:  acc = a0*b0; acc = fma(ak,bk,acc)
 *   - every other scalar expression is evaluated as written, one rounding per operator
 *     (compile with -ffp-contract=off)
 *   - acos/sin/cos/tan/atan2 come from include/rpgo_elem.h (deterministic)
 *
 * Parity status: pinned at 1e-9 against the golden covariances of
 * tests/testPoseWithCovariance.cpp:68-77,107-115,141-149,194-200 and the decision-level
 * expectations of tests/testPcm.cpp, testPcmSimple.cpp, testMultiRobot.cpp, testLoadGraph.cpp;
 * bit-level parity with a real GTSAM/Eigen build is UNPINNED (no such build exists here).
 */
#ifndef ORACLE_MATH_H_
#define ORACLE_MATH_H_

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "rpgo_elem.h"

#define O_MAXN 6

/* pose storage: d==3: m[0..8] = R row-major, m[9..11] = t.   d==2: m[0]=c m[1]=s m[2]=x m[3]=y */
typedef struct {
  double m[12];
} opose;

static inline int o_n(int d) { return d == 3 ? 6 : 3; }
static inline int o_rdim(int d) { return d == 3 ? 3 : 1; } /* GeometryUtils.h:28-34 */
static inline int o_tdim(int d) { return d == 3 ? 3 : 2; } /* GeometryUtils.h:36-42 */

static inline void o_pose_identity(int d, opose* p) {
  memset(p, 0, sizeof(*p));
  if (d == 3) {
    p->m[0] = p->m[4] = p->m[8] = 1.0;
  } else {
    p->m[0] = 1.0;
  }
}

/* ---- dense helpers ------------------------------------------------------------- */
/* C = A(n x n) * B(n x n), Eigen coefficient-based lazy product order (k ascending) */
static inline void o_matmul(int n, const double* A, const double* B, double* C) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double acc = A[i * n + 0] * B[0 * n + j];
      for (int k = 1; k < n; ++k) acc = fma(A[i * n + k], B[k * n + j], acc);
      C[i * n + j] = acc;
    }
}
/* C = A * B^T */
static inline void o_matmul_bt(int n, const double* A, const double* B, double* C) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double acc = A[i * n + 0] * B[j * n + 0];
      for (int k = 1; k < n; ++k) acc = fma(A[i * n + k], B[j * n + k], acc);
      C[i * n + j] = acc;
    }
}
static inline void o_mat3_vec(const double* R, const double* v, double* out) {
  for (int i = 0; i < 3; ++i) {
    double acc = R[i * 3 + 0] * v[0];
    acc = fma(R[i * 3 + 1], v[1], acc);
    acc = fma(R[i * 3 + 2], v[2], acc);
    out[i] = acc;
  }
}

/* ---- Rot2 (gtsam/geometry/Rot2.{h,cpp}) ------------------------------------------ */
/* Rot2::fromCosSin -> normalize(): rescale when |c^2+s^2-1| > 1e-10 */
static inline void o_rot2_from_cos_sin(double c, double s, double* co, double* so) {
  double scale = fma(s, s, c * c);
  if (fabs(scale - 1.0) > 1e-10) {
    scale = 1.0 / sqrt(scale);
    c = c * scale;
    s = s * scale;
  }
  *co = c;
  *so = s;
}

/* ---- group operations ---------------------------------------------------------- */
/* Pose3::operator* : (R1*R2, t1 + R1*t2);  Pose2::operator* : (r1*r2, t1 + r1*t2) */
static inline void o_pose_compose(int d, const opose* a, const opose* b, opose* out) {
  opose r;
  if (d == 3) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = a->m[i * 3 + 0] * b->m[0 * 3 + j];
        acc = fma(a->m[i * 3 + 1], b->m[1 * 3 + j], acc);
        acc = fma(a->m[i * 3 + 2], b->m[2 * 3 + j], acc);
        r.m[i * 3 + j] = acc;
      }
    double rt[3];
    o_mat3_vec(a->m, b->m + 9, rt);
    for (int i = 0; i < 3; ++i) r.m[9 + i] = a->m[9 + i] + rt[i];
  } else {
    const double c1 = a->m[0], s1 = a->m[1], c2 = b->m[0], s2 = b->m[1];
    /* Rot2::operator*: fromCosSin(c1*c2 - s1*s2, s1*c2 + c1*s2) */
    const double c = fma(-s1, s2, c1 * c2);
    const double s = fma(c1, s2, s1 * c2);
    o_rot2_from_cos_sin(c, s, &r.m[0], &r.m[1]);
    /* Rot2::rotate: (c*x + -s*y, s*x + c*y) */
    const double rx = fma(-s1, b->m[3], c1 * b->m[2]);
    const double ry = fma(c1, b->m[3], s1 * b->m[2]);
    r.m[2] = a->m[2] + rx;
    r.m[3] = a->m[3] + ry;
    for (int i = 4; i < 12; ++i) r.m[i] = 0.0;
  }
  *out = r;
}

/* Pose3::inverse : Rt = R^T; (Rt, Rt * (-t)).  Pose2::inverse : (r^-1, r.unrotate(-t)) */
static inline void o_pose_inverse(int d, const opose* a, opose* out) {
  opose r;
  if (d == 3) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a->m[j * 3 + i];
    double nt[3] = {-a->m[9], -a->m[10], -a->m[11]};
    o_mat3_vec(r.m, nt, r.m + 9);
  } else {
    const double c = a->m[0], s = a->m[1];
    const double nx = -a->m[2], ny = -a->m[3];
    r.m[0] = c;
    r.m[1] = -s;
    /* Rot2::unrotate: (c*x + s*y, -s*x + c*y) */
    r.m[2] = fma(s, ny, c * nx);
    r.m[3] = fma(c, ny, (-s) * nx);
    for (int i = 4; i < 12; ++i) r.m[i] = 0.0;
  }
  *out = r;
}

/* LieGroup::between : inverse() * g */
static inline void o_pose_between(int d, const opose* a, const opose* b, opose* out) {
  opose ai;
  o_pose_inverse(d, a, &ai);
  o_pose_compose(d, &ai, b, out);
}

/* Pose3::AdjointMap : [R 0; [t]x R, R]   (tangent order [omega; v])
 * Pose2::AdjointMap : [c -s y; s c -x; 0 0 1]   (tangent order (x, y, theta)) */
static inline void o_adjoint(int d, const opose* p, double* Ad) {
  if (d == 3) {
    const double* R = p->m;
    const double tx = p->m[9], ty = p->m[10], tz = p->m[11];
    const double S[9] = {0.0, -tz, ty, tz, 0.0, -tx, -ty, tx, 0.0};
    double A[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = S[i * 3 + 0] * R[0 * 3 + j];
        acc = fma(S[i * 3 + 1], R[1 * 3 + j], acc);
        acc = fma(S[i * 3 + 2], R[2 * 3 + j], acc);
        A[i * 3 + j] = acc;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Ad[i * 6 + j] = R[i * 3 + j];
        Ad[i * 6 + 3 + j] = 0.0;
        Ad[(3 + i) * 6 + j] = A[i * 3 + j];
        Ad[(3 + i) * 6 + 3 + j] = R[i * 3 + j];
      }
  } else {
    const double c = p->m[0], s = p->m[1], x = p->m[2], y = p->m[3];
    Ad[0] = c;   Ad[1] = -s;  Ad[2] = y;
    Ad[3] = s;   Ad[4] = c;   Ad[5] = -x;
    Ad[6] = 0.0; Ad[7] = 0.0; Ad[8] = 1.0;
  }
}

/* SO3::Logmap (gtsam/geometry/SO3.cpp, GTSAM 4.0/4.1 constants) */
static inline void o_so3_logmap(const double* R, double* w) {
  const double R11 = R[0], R12 = R[1], R13 = R[2];
  const double R21 = R[3], R22 = R[4], R23 = R[5];
  const double R31 = R[6], R32 = R[7], R33 = R[8];
  const double tr = (R11 + R22) + R33;
  if (tr + 1.0 < 1e-10) {
    if (fabs(R33 + 1.0) > 1e-5) {
      const double f = RPGO_PI_1 / sqrt(2.0 + 2.0 * R33);
      w[0] = f * R13; w[1] = f * R23; w[2] = f * (1.0 + R33);
    } else if (fabs(R22 + 1.0) > 1e-5) {
      const double f = RPGO_PI_1 / sqrt(2.0 + 2.0 * R22);
      w[0] = f * R12; w[1] = f * (1.0 + R22); w[2] = f * R32;
    } else {
      const double f = RPGO_PI_1 / sqrt(2.0 + 2.0 * R11);
      w[0] = f * (1.0 + R11); w[1] = f * R21; w[2] = f * R31;
    }
  } else {
    double magnitude;
    const double tr_3 = tr - 3.0;
    if (tr_3 < -1e-7) {
      const double theta = rpgo_acos((tr - 1.0) / 2.0);
      magnitude = theta / (2.0 * rpgo_sin(theta));
    } else {
      magnitude = 0.5 - tr_3 / 12.0;
    }
    w[0] = magnitude * (R32 - R23);
    w[1] = magnitude * (R13 - R31);
    w[2] = magnitude * (R21 - R12);
  }
}

/* Pose3::Logmap / Pose2::Logmap (static, full SE(3)/SE(2) log) */
static inline void o_logmap(int d, const opose* p, double* v) {
  if (d == 3) {
    double w[3];
    o_so3_logmap(p->m, w);
    const double* T = p->m + 9;
    const double t = sqrt(fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0])));
    if (t < 1e-10) {
      v[0] = w[0]; v[1] = w[1]; v[2] = w[2];
      v[3] = T[0]; v[4] = T[1]; v[5] = T[2];
    } else {
      const double wx = w[0] / t, wy = w[1] / t, wz = w[2] / t;
      const double W[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
      const double Tan = rpgo_tan(0.5 * t);
      double WT[3], WWT[3];
      o_mat3_vec(W, T, WT);
      o_mat3_vec(W, WT, WWT);
      const double a = 0.5 * t;
      const double b = 1.0 - t / (2.0 * Tan);
      v[0] = w[0]; v[1] = w[1]; v[2] = w[2];
      for (int i = 0; i < 3; ++i) v[3 + i] = (T[i] - a * WT[i]) + b * WWT[i];
    }
  } else {
    const double c = p->m[0], s = p->m[1], x = p->m[2], y = p->m[3];
    const double w = rpgo_atan2(s, c);
    if (fabs(w) < 1e-10) {
      v[0] = x; v[1] = y; v[2] = w;
    } else {
      const double c_1 = c - 1.0;
      const double det = fma(s, s, c_1 * c_1);
      /* R.unrotate(t) - t */
      const double ux = fma(s, y, c * x) - x;
      const double uy = fma(c, y, (-s) * x) - y;
      /* R_PI_2 * (.) : Rot2(0,1).rotate = (0*x + -1*y, 1*x + 0*y) */
      const double px = fma(-1.0, uy, 0.0 * ux);
      const double py = fma(0.0, uy, 1.0 * ux);
      const double f = w / det;
      v[0] = f * px; v[1] = f * py; v[2] = w;
    }
  }
}

/* Eigen::LLT<MatrixXd> unblocked in-place (lower).  Returns 1 on Success, 0 on NumericalIssue. */
static inline int o_llt_ok(int n, const double* Min) {
  double A[O_MAXN * O_MAXN];
  memcpy(A, Min, sizeof(double) * n * n);
  for (int k = 0; k < n; ++k) {
    double x = A[k * n + k];
    if (k > 0) {
      double sn = A[k * n + 0] * A[k * n + 0];
      for (int j = 1; j < k; ++j) sn = fma(A[k * n + j], A[k * n + j], sn);
      x = x - sn;
    }
    if (x <= 0.0) return 0;
    x = sqrt(x);
    A[k * n + k] = x;
    for (int i = k + 1; i < n; ++i) {
      if (k > 0) {
        double dot = A[i * n + 0] * A[k * n + 0];
        for (int j = 1; j < k; ++j) dot = fma(A[i * n + j], A[k * n + j], dot);
        A[i * n + k] = A[i * n + k] - dot;
      }
      A[i * n + k] = A[i * n + k] / x;
    }
  }
  return 1;
}

/* Eigen dynamic MatrixXd::inverse() = PartialPivLU (unblocked, size <= 16) + solve(Identity):
 * dst = P*I; unit-lower solve; upper solve — column-major right-looking kernels, the upper
 * solve multiplying by the reciprocal of the diagonal (Eigen TriangularSolverMatrix.h). */
static inline void o_lu_inverse(int n, const double* Min, double* inv) {
  double lu[O_MAXN * O_MAXN];
  int perm[O_MAXN];
  memcpy(lu, Min, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double biggest = fabs(lu[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(lu[i * n + k]);
      if (v > biggest) { biggest = v; piv = i; }
    }
    if (biggest != 0.0) {
      if (piv != k) {
        for (int j = 0; j < n; ++j) {
          const double tmp = lu[k * n + j];
          lu[k * n + j] = lu[piv * n + j];
          lu[piv * n + j] = tmp;
        }
        const int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
      }
      for (int i = k + 1; i < n; ++i) lu[i * n + k] = lu[i * n + k] / lu[k * n + k];
    }
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j < n; ++j) lu[i * n + j] = fma(-lu[i * n + k], lu[k * n + j], lu[i * n + j]);
  }
  /* rhs = P * I : row i of rhs is e_{perm[i]} */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (perm[i] == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; ++c) {
    /* unit lower, forward */
    for (int i = 0; i < n; ++i) {
      const double b = inv[i * n + c];
      for (int r = i + 1; r < n; ++r) inv[r * n + c] = fma(-b, lu[r * n + i], inv[r * n + c]);
    }
    /* upper, backward */
    for (int i = n - 1; i >= 0; --i) {
      const double a = 1.0 / lu[i * n + i];
      const double b = inv[i * n + c] * a;
      inv[i * n + c] = b;
      for (int r = 0; r < i; ++r) inv[r * n + c] = fma(-b, lu[r * n + i], inv[r * n + c]);
    }
  }
}

/* ---- PoseWithCovariance (GeometryUtils.h:56-187) -------------------------------- */
typedef struct {
  opose pose;
  double cov[O_MAXN * O_MAXN];
  int rotation_info;
} opwc;

static inline void o_pwc_default(int d, opwc* p) { /* :65-71 */
  o_pose_identity(d, &p->pose);
  memset(p->cov, 0, sizeof(p->cov));
  p->rotation_info = 1;
}

/* ctor from BetweenFactor, GeometryUtils.h:91-115 */
static inline void o_pwc_from_factor(int d, const double* pose, const double* cov, opwc* p) {
  const int n = o_n(d), r = o_rdim(d), t = o_tdim(d);
  memset(p, 0, sizeof(*p));
  memcpy(p->pose.m, pose, sizeof(double) * (d == 3 ? 12 : 4));
  double tr = 0.0;
  for (int i = 0; i < r; ++i) tr = (i == 0) ? cov[0] : tr + cov[i * n + i];
  p->rotation_info = 1;
  if (isnan(tr)) {
    p->rotation_info = 0;
    for (int i = 0; i < t; ++i)
      for (int j = 0; j < t; ++j) p->cov[(r + i) * n + r + j] = cov[(r + i) * n + r + j];
  } else {
    memcpy(p->cov, cov, sizeof(double) * n * n);
  }
}

/* compose, GeometryUtils.h:119-129.  Ha = Ad(other^-1), Hb = I */
static inline void o_pwc_compose(int d, const opwc* a, const opwc* b, opwc* out) {
  const int n = o_n(d);
  opwc r;
  memset(&r, 0, sizeof(r));
  o_pose_compose(d, &a->pose, &b->pose, &r.pose);
  opose binv;
  o_pose_inverse(d, &b->pose, &binv);
  double Ha[36], Hb[36], T1[36], T2[36], T3[36], T4[36];
  o_adjoint(d, &binv, Ha);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Hb[i * n + j] = (i == j) ? 1.0 : 0.0;
  o_matmul(n, Ha, a->cov, T1);
  o_matmul_bt(n, T1, Ha, T2);
  o_matmul(n, Hb, b->cov, T3);
  o_matmul_bt(n, T3, Hb, T4);
  for (int i = 0; i < n * n; ++i) r.cov[i] = T2[i] + T4[i];
  r.rotation_info = (a->rotation_info && b->rotation_info) ? 1 : 0;
  *out = r;
}

/* inverse, GeometryUtils.h:133-139: pose inverted, covariance copied unchanged */
static inline void o_pwc_inverse(int d, const opwc* a, opwc* out) {
  opwc r = *a;
  o_pose_inverse(d, &a->pose, &r.pose);
  *out = r;
}

/* between, GeometryUtils.h:143-170.  Ha = -Ad((A^-1 B)^-1) */
static inline void o_pwc_between(int d, const opwc* a, const opwc* b, opwc* out) {
  const int n = o_n(d);
  opwc r;
  memset(&r, 0, sizeof(r));
  o_pose_between(d, &a->pose, &b->pose, &r.pose);
  opose rinv;
  o_pose_inverse(d, &r.pose, &rinv);
  double Ha[36], T1[36], T2[36];
  o_adjoint(d, &rinv, Ha);
  for (int i = 0; i < n * n; ++i) Ha[i] = -Ha[i];
  o_matmul(n, Ha, a->cov, T1);
  o_matmul_bt(n, T1, Ha, T2);
  for (int i = 0; i < n * n; ++i) r.cov[i] = b->cov[i] - T2[i];
  if (!o_llt_ok(n, r.cov)) {
    /* :157-161  other.pose.between(pose, Ha, Hb); cov = cov_a - Ha cov_b Ha^T (pose kept) */
    opose r2, r2inv;
    o_pose_between(d, &b->pose, &a->pose, &r2);
    o_pose_inverse(d, &r2, &r2inv);
    o_adjoint(d, &r2inv, Ha);
    for (int i = 0; i < n * n; ++i) Ha[i] = -Ha[i];
    o_matmul(n, Ha, b->cov, T1);
    o_matmul_bt(n, T1, Ha, T2);
    for (int i = 0; i < n * n; ++i) r.cov[i] = a->cov[i] - T2[i];
  }
  r.rotation_info = (a->rotation_info && b->rotation_info) ? 1 : 0;
  *out = r;
}

/* mahalanobis_norm, GeometryUtils.h:172-186 */
static inline double o_pwc_mahalanobis(int d, const opwc* a) {
  const int n = o_n(d), r = o_rdim(d), t = o_tdim(d);
  double lg[6];
  o_logmap(d, &a->pose, lg);
  double blk[36], inv[36], v[6];
  int m, off;
  if (!a->rotation_info) {
    m = t; off = r;
    for (int i = 0; i < t; ++i)
      for (int j = 0; j < t; ++j) blk[i * t + j] = a->cov[(r + i) * n + r + j];
  } else {
    m = n; off = 0;
    memcpy(blk, a->cov, sizeof(double) * n * n);
  }
  o_lu_inverse(m, blk, inv);
  /* (log^T * inv) * log */
  for (int j = 0; j < m; ++j) {
    double acc = lg[off + 0] * inv[0 * m + j];
    for (int i = 1; i < m; ++i) acc = fma(lg[off + i], inv[i * m + j], acc);
    v[j] = acc;
  }
  double q = v[0] * lg[off + 0];
  for (int j = 1; j < m; ++j) q = fma(v[j], lg[off + j], q);
  return sqrt(q);
}

/* ---- PoseWithNode (GeometryUtils.h:193-289) ------------------------------------- */
typedef struct {
  opose pose;
  int node;
  int rotation_info;
} opwn;

static inline void o_pwn_default(int d, opwn* p) {
  o_pose_identity(d, &p->pose);
  p->node = 0;
  p->rotation_info = 1;
}
static inline void o_pwn_from_factor(int d, const double* pose, const double* cov, opwn* p) {
  const int n = o_n(d), r = o_rdim(d);
  memset(p, 0, sizeof(*p));
  memcpy(p->pose.m, pose, sizeof(double) * (d == 3 ? 12 : 4));
  double tr = 0.0;
  for (int i = 0; i < r; ++i) tr = (i == 0) ? cov[0] : tr + cov[i * n + i];
  p->rotation_info = isnan(tr) ? 0 : 1;
  p->node = 1;
}
static inline void o_pwn_compose(int d, const opwn* a, const opwn* b, opwn* out) {
  opwn r;
  o_pose_compose(d, &a->pose, &b->pose, &r.pose);
  r.node = a->node + b->node;
  r.rotation_info = (a->rotation_info && b->rotation_info) ? 1 : 0;
  *out = r;
}
static inline void o_pwn_inverse(int d, const opwn* a, opwn* out) {
  opwn r = *a;
  o_pose_inverse(d, &a->pose, &r.pose);
  *out = r;
}
static inline void o_pwn_between(int d, const opwn* a, const opwn* b, opwn* out) {
  opwn r;
  o_pose_between(d, &a->pose, &b->pose, &r.pose);
  r.node = abs(b->node - a->node);
  r.rotation_info = (a->rotation_info && b->rotation_info) ? 1 : 0;
  *out = r;
}
/* avg_trans_norm :275-280 — note log.tail(t_dim): for Pose2 this is (y, theta) (reference quirk) */
static inline double o_pwn_avg_trans(int d, const opwn* a) {
  const int n = o_n(d), t = o_tdim(d);
  double lg[6];
  o_logmap(d, &a->pose, lg);
  double q = lg[n - t] * lg[n - t];
  for (int i = 1; i < t; ++i) q = fma(lg[n - t + i], lg[n - t + i], q);
  return sqrt(q) / (double)a->node;
}
/* avg_rot_norm :282-288 — log.head(r_dim): for Pose2 this is (x) (reference quirk) */
static inline double o_pwn_avg_rot(int d, const opwn* a) {
  const int r = o_rdim(d);
  if (!a->rotation_info) return 0.0;
  double lg[6];
  o_logmap(d, &a->pose, lg);
  double q = lg[0] * lg[0];
  for (int i = 1; i < r; ++i) q = fma(lg[i], lg[i], q);
  return sqrt(q) / (double)a->node;
}

#endif /* ORACLE_MATH_H_ */
