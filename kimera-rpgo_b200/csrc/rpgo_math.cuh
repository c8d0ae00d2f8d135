/* rpgo_math.cuh — fp64 SE(2)/SE(3) pose + covariance algebra for the PCM hot path (host + device).
 *
 * Replaces, on the GPU, the arithmetic of
 *   KimeraRPGO::PoseWithCovariance / PoseWithNode   (reference include/KimeraRPGO/utils/GeometryUtils.h:56-289)
 * and the GTSAM Lie-group calls it makes (Pose3/Pose2 compose, inverse, between with Jacobians,
 * Logmap; Eigen LLT and PartialPivLU inverse).
 *
 * Numerical contract (DESIGN.md "Numerics"): identical to the CPU oracle, bit for bit —
 *   - matrix/dot products accumulate in k-order: acc = a0*b0; acc = fma(ak, bk, acc)
 *   - every other expression: one rounding per written operator (nvcc -fmad=false, gcc -ffp-contract=off)
 *   - transcendental functions from include/rpgo_elem.h
 * Unlike the oracle's dense loops this header exploits structure WITHOUT changing any rounding:
 *   Ad(T) = [R 0; [t]xR R] has a zero block (skipped terms are exact zeros), compose's Hb = I is never
 *   multiplied out, and the sign of between's Ha = -Ad(.) cancels exactly in Ha S Ha^T.
 * All loops have compile-time trip counts so that every matrix lives in registers.
 */
#pragma once

#include <math.h>
#include <stdint.h>

#include "rpgo_elem.h"

#if defined(__CUDACC__)
#define RPGO_FN __host__ __device__ __forceinline__
#define RPGO_UNROLL _Pragma("unroll")
#else
#define RPGO_FN static inline
#define RPGO_UNROLL
#endif

namespace rpgo {

enum { MODE_PCM = 0, MODE_SIMPLE = 1 };

template <int D>
struct Dim;
template <>
struct Dim<3> {
  static constexpr int N = 6, PS = 12, RD = 3, TD = 3;
  /* entry layout (doubles): pose[12] cov[36] rot_info node */
  static constexpr int ENTRY = 50, OFF_COV = 12, OFF_ROT = 48, OFF_NODE = 49;
};
template <>
struct Dim<2> {
  static constexpr int N = 3, PS = 4, RD = 1, TD = 2;
  /* entry layout (doubles): pose[4] cov[9] rot_info node, padded to 16 */
  static constexpr int ENTRY = 16, OFF_COV = 4, OFF_ROT = 13, OFF_NODE = 14;
};

/* pose: D==3: m[0..8] R row-major, m[9..11] t;  D==2: m[0]=c m[1]=s m[2]=x m[3]=y */
template <int D>
struct Pose {
  double m[Dim<D>::PS];
};

template <int D>
RPGO_FN void pose_identity(Pose<D>& p) {
  RPGO_UNROLL
  for (int i = 0; i < Dim<D>::PS; ++i) p.m[i] = 0.0;
  if (D == 3) {
    p.m[0] = 1.0; p.m[4] = 1.0; p.m[8] = 1.0;
  } else {
    p.m[0] = 1.0;
  }
}

RPGO_FN double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
  double acc = a0 * b0;
  acc = fma(a1, b1, acc);
  acc = fma(a2, b2, acc);
  return acc;
}

/* Rot2::fromCosSin -> normalize() */
RPGO_FN void rot2_from_cos_sin(double c, double s, double& co, double& so) {
  double scale = fma(s, s, c * c);
  if (fabs(scale - 1.0) > 1e-10) {
    scale = 1.0 / sqrt(scale);
    c = c * scale;
    s = s * scale;
  }
  co = c;
  so = s;
}

/* Pose::operator*  (R1 R2, t1 + R1 t2) */
template <int D>
RPGO_FN Pose<D> compose(const Pose<D>& a, const Pose<D>& b) {
  Pose<D> r;
  if (D == 3) {
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i) {
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j)
        r.m[i * 3 + j] = dot3(a.m[i * 3], a.m[i * 3 + 1], a.m[i * 3 + 2], b.m[j], b.m[3 + j], b.m[6 + j]);
    }
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i)
      r.m[9 + i] = a.m[9 + i] + dot3(a.m[i * 3], a.m[i * 3 + 1], a.m[i * 3 + 2], b.m[9], b.m[10], b.m[11]);
  } else {
    const double c1 = a.m[0], s1 = a.m[1], c2 = b.m[0], s2 = b.m[1];
    const double c = fma(-s1, s2, c1 * c2);
    const double s = fma(c1, s2, s1 * c2);
    rot2_from_cos_sin(c, s, r.m[0], r.m[1]);
    const double rx = fma(-s1, b.m[3], c1 * b.m[2]);
    const double ry = fma(c1, b.m[3], s1 * b.m[2]);
    r.m[2] = a.m[2] + rx;
    r.m[3] = a.m[3] + ry;
  }
  return r;
}

/* Pose::inverse  (R^T, R^T (-t)) */
template <int D>
RPGO_FN Pose<D> inverse(const Pose<D>& a) {
  Pose<D> r;
  if (D == 3) {
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i) {
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[j * 3 + i];
    }
    const double n0 = -a.m[9], n1 = -a.m[10], n2 = -a.m[11];
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i) r.m[9 + i] = dot3(r.m[i * 3], r.m[i * 3 + 1], r.m[i * 3 + 2], n0, n1, n2);
  } else {
    const double c = a.m[0], s = a.m[1];
    const double nx = -a.m[2], ny = -a.m[3];
    r.m[0] = c;
    r.m[1] = -s;
    r.m[2] = fma(s, ny, c * nx);
    r.m[3] = fma(c, ny, (-s) * nx);
  }
  return r;
}

template <int D>
RPGO_FN Pose<D> between(const Pose<D>& a, const Pose<D>& b) {
  return compose<D>(inverse<D>(a), b);
}

/* ---- adjoint, stored compactly --------------------------------------------------------------
 * D==3: h[0..8] = A = R, h[9..17] = B = [t]x R  (Ad = [A 0; B A]);   D==2: h[0..8] dense 3x3.  */
template <int D>
struct Adj {
  double h[D == 3 ? 18 : 9];
};

template <int D>
RPGO_FN Adj<D> adjoint(const Pose<D>& p) {
  Adj<D> a;
  if (D == 3) {
    const double tx = p.m[9], ty = p.m[10], tz = p.m[11];
    RPGO_UNROLL
    for (int j = 0; j < 3; ++j) {
      const double r0 = p.m[j], r1 = p.m[3 + j], r2 = p.m[6 + j];
      a.h[j] = r0; a.h[3 + j] = r1; a.h[6 + j] = r2;
      /* rows of [t]x = (0,-tz,ty), (tz,0,-tx), (-ty,tx,0); zero terms skipped */
      a.h[9 + j] = fma(ty, r2, (-tz) * r1);
      a.h[12 + j] = fma(-tx, r2, tz * r0);
      a.h[15 + j] = fma(tx, r1, (-ty) * r0);
    }
  } else {
    const double c = p.m[0], s = p.m[1], x = p.m[2], y = p.m[3];
    a.h[0] = c;   a.h[1] = -s;  a.h[2] = y;
    a.h[3] = s;   a.h[4] = c;   a.h[5] = -x;
    a.h[6] = 0.0; a.h[7] = 0.0; a.h[8] = 1.0;
  }
  return a;
}

/* out = H S H^T evaluated as (H S) H^T, k-order, H = Ad.  S, out: N x N row-major.
 * One row of T1 = H S at a time (6 temporaries), so S and out are the only full matrices live. */
template <int D, typename LoadS>
RPGO_FN void hsht(const Adj<D>& H, LoadS S, double* out) {
  if (D == 3) {
    const double* A = H.h;
    const double* B = H.h + 9;
    RPGO_UNROLL
    for (int i = 0; i < 6; ++i) {
      double t[6];
      if (i < 3) {
        RPGO_UNROLL
        for (int j = 0; j < 6; ++j) t[j] = dot3(A[i * 3], A[i * 3 + 1], A[i * 3 + 2], S(0, j), S(1, j), S(2, j));
      } else {
        const int r = i - 3;
        RPGO_UNROLL
        for (int j = 0; j < 6; ++j) {
          double acc = dot3(B[r * 3], B[r * 3 + 1], B[r * 3 + 2], S(0, j), S(1, j), S(2, j));
          acc = fma(A[r * 3], S(3, j), acc);
          acc = fma(A[r * 3 + 1], S(4, j), acc);
          acc = fma(A[r * 3 + 2], S(5, j), acc);
          t[j] = acc;
        }
      }
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) out[i * 6 + j] = dot3(t[0], t[1], t[2], A[j * 3], A[j * 3 + 1], A[j * 3 + 2]);
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) {
        double acc = dot3(t[0], t[1], t[2], B[j * 3], B[j * 3 + 1], B[j * 3 + 2]);
        acc = fma(t[3], A[j * 3], acc);
        acc = fma(t[4], A[j * 3 + 1], acc);
        acc = fma(t[5], A[j * 3 + 2], acc);
        out[i * 6 + 3 + j] = acc;
      }
    }
  } else {
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i) {
      double t[3];
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) t[j] = dot3(H.h[i * 3], H.h[i * 3 + 1], H.h[i * 3 + 2], S(0, j), S(1, j), S(2, j));
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) out[i * 3 + j] = dot3(t[0], t[1], t[2], H.h[j * 3], H.h[j * 3 + 1], H.h[j * 3 + 2]);
    }
  }
}

/* M <- H M H^T in place, the same entries as hsht() (every entry keeps its own k-order chain, so the bits are
 * identical): first T = H M column by column (a column of T depends only on the same column of M), then T H^T row by
 * row.  Only M, H and 12 temporaries are live instead of two full matrices. */
template <int D>
RPGO_FN void hsht_inplace(const Adj<D>& H, double* M) {
  if (D == 3) {
    const double* A = H.h;
    const double* B = H.h + 9;
    RPGO_UNROLL
    for (int c = 0; c < 6; ++c) {
      const double s0 = M[c], s1 = M[6 + c], s2 = M[12 + c], s3 = M[18 + c], s4 = M[24 + c], s5 = M[30 + c];
      RPGO_UNROLL
      for (int r = 0; r < 3; ++r) M[r * 6 + c] = dot3(A[r * 3], A[r * 3 + 1], A[r * 3 + 2], s0, s1, s2);
      RPGO_UNROLL
      for (int r = 0; r < 3; ++r) {
        double acc = dot3(B[r * 3], B[r * 3 + 1], B[r * 3 + 2], s0, s1, s2);
        acc = fma(A[r * 3], s3, acc);
        acc = fma(A[r * 3 + 1], s4, acc);
        acc = fma(A[r * 3 + 2], s5, acc);
        M[(3 + r) * 6 + c] = acc;
      }
    }
    RPGO_UNROLL
    for (int i = 0; i < 6; ++i) {
      const double t0 = M[i * 6], t1 = M[i * 6 + 1], t2 = M[i * 6 + 2], t3 = M[i * 6 + 3], t4 = M[i * 6 + 4], t5 = M[i * 6 + 5];
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) M[i * 6 + j] = dot3(t0, t1, t2, A[j * 3], A[j * 3 + 1], A[j * 3 + 2]);
      RPGO_UNROLL
      for (int j = 0; j < 3; ++j) {
        double acc = dot3(t0, t1, t2, B[j * 3], B[j * 3 + 1], B[j * 3 + 2]);
        acc = fma(t3, A[j * 3], acc);
        acc = fma(t4, A[j * 3 + 1], acc);
        acc = fma(t5, A[j * 3 + 2], acc);
        M[i * 6 + 3 + j] = acc;
      }
    }
  } else {
    /* Ad(Pose2) has the last row (0, 0, 1) (adjoint<2>): row 2 of H M is row 2 of M and column 2 of (H M) H^T is column 2
     * of H M -- the dense products 0*x + 0*y + 1*z return z exactly for finite x, y (only the sign of a zero result can
     * differ, which no later operation observes), so 18 of the 54 operations are skipped, as the zero blocks of the
     * Pose3 adjoint are above */
    RPGO_UNROLL
    for (int c = 0; c < 3; ++c) {
      const double s0 = M[c], s1 = M[3 + c], s2 = M[6 + c];
      RPGO_UNROLL
      for (int r = 0; r < 2; ++r) M[r * 3 + c] = dot3(H.h[r * 3], H.h[r * 3 + 1], H.h[r * 3 + 2], s0, s1, s2);
    }
    RPGO_UNROLL
    for (int i = 0; i < 3; ++i) {
      const double t0 = M[i * 3], t1 = M[i * 3 + 1], t2 = M[i * 3 + 2];
      RPGO_UNROLL
      for (int j = 0; j < 2; ++j) M[i * 3 + j] = dot3(t0, t1, t2, H.h[j * 3], H.h[j * 3 + 1], H.h[j * 3 + 2]);
    }
  }
}

/* Eigen::LLT unblocked (lower): 1 = Success, 0 = NumericalIssue (a pivot <= 0) */
template <int N>
RPGO_FN bool llt_ok(const double* Min) {
  double A[N * N];
  RPGO_UNROLL
  for (int i = 0; i < N; ++i) {
    RPGO_UNROLL
    for (int j = 0; j <= i; ++j) A[i * N + j] = Min[i * N + j];
  }
  bool ok = true;
  RPGO_UNROLL
  for (int k = 0; k < N; ++k) {
    double x = A[k * N + k];
    if (k > 0) {
      double sn = A[k * N] * A[k * N];
      RPGO_UNROLL
      for (int j = 1; j < k; ++j) sn = fma(A[k * N + j], A[k * N + j], sn);
      x = x - sn;
    }
    if (ok && x <= 0.0) ok = false;
    if (!ok) break;
    x = sqrt(x);
    A[k * N + k] = x;
    RPGO_UNROLL
    for (int i = k + 1; i < N; ++i) {
      if (k > 0) {
        double dot = A[i * N] * A[k * N];
        RPGO_UNROLL
        for (int j = 1; j < k; ++j) dot = fma(A[i * N + j], A[k * N + j], dot);
        A[i * N + k] = A[i * N + k] - dot;
      }
      A[i * N + k] = A[i * N + k] / x;
    }
  }
  return ok;
}


/* a / x given r = RN(1/x) (IEEE reciprocal): q0 = RN(a r); rem = a - q0 x (exact, one FMA);
 * q = RN(q0 + rem r).  By Markstein's theorem q == RN(a / x) whenever r is the correctly rounded
 * reciprocal and nothing over/underflows; outside that range we fall back to the plain division.
 * Used where several numerators share one divisor (a column of the LLT / LU factorisations), so the
 * expensive reciprocal is paid once.  tests/test_cpu_math_parity.py checks it against `/` on 10^7
 * adversarial operands and the GPU suite checks the kernels that use it against the kernel that does
 * not (bitset equality over 2*10^8 pairs). */
#if defined(__CUDACC__)
__host__ __device__ __noinline__ double slow_div(double a, double x);
#if defined(RPGO_MATH_IMPL)
__host__ __device__ __noinline__ double slow_div(double a, double x) { return a / x; }
#endif
#else
static double slow_div(double a, double x) { return a / x; }
#endif

/* exponent-range guard evaluated on the integer pipe (device) so that it does not compete for FP64 issue */
RPGO_FN bool quot_in_safe_range(double q0, double a) {
#if defined(__CUDA_ARCH__)
  const int e = (__double2hiint(q0) >> 20) & 0x7ff;
  return (e > 100 && e < 1900) || a == 0.0;
#else
  const double aq = fabs(q0);
  return (aq > 0x1p-922 && aq < 0x1p+877) || a == 0.0;
#endif
}
RPGO_FN double div_by(double a, double x, double r) {
  const double q0 = a * r;
  const double rem = fma(-q0, x, a);
  double q = fma(rem, r, q0);
  if (!quot_in_safe_range(q0, a)) q = slow_div(a, x); /* rare: one shared out-of-line division */
  return q;
}
RPGO_FN bool rcp_safe(double x) {
  const double ax = fabs(x);
  return ax > 1e-140 && ax < 1e140;
}

/* llt_ok with the column divisions done through div_by (same results as llt_ok) */
template <int N>
RPGO_FN bool llt_ok_fast(const double* Min) {
  double A[N * N];
  RPGO_UNROLL
  for (int i = 0; i < N; ++i) {
    RPGO_UNROLL
    for (int j = 0; j <= i; ++j) A[i * N + j] = Min[i * N + j];
  }
  bool ok = true;
  RPGO_UNROLL
  for (int k = 0; k < N; ++k) {
    double x = A[k * N + k];
    if (k > 0) {
      double sn = A[k * N] * A[k * N];
      RPGO_UNROLL
      for (int j = 1; j < k; ++j) sn = fma(A[k * N + j], A[k * N + j], sn);
      x = x - sn;
    }
    if (ok && x <= 0.0) ok = false;
    if (!ok) break;
    x = sqrt(x);
    A[k * N + k] = x;
    if (k + 1 < N) {
      const bool fast = rcp_safe(x);
      const double r = 1.0 / x;
      RPGO_UNROLL
      for (int i = k + 1; i < N; ++i) {
        double v = A[i * N + k];
        if (k > 0) {
          double dot = A[i * N] * A[k * N];
          RPGO_UNROLL
          for (int j = 1; j < k; ++j) dot = fma(A[i * N + j], A[k * N + j], dot);
          v = v - dot;
        }
        if (fast) v = div_by(v, x, r); else v = slow_div(v, x);
        A[i * N + k] = v;
      }
    }
  }
  return ok;
}

/* q = v^T M^-1 v the way the reference computes it: M^-1 by Eigen PartialPivLU + solve(Identity)
 * (column by column), then (v^T M^-1) v.  Row swaps are predicated so that everything stays in
 * registers. */
template <int N>
RPGO_FN double quad_form_inv(const double* Min, const double* v) {
  double lu[N * N];
  int perm[N];
  RPGO_UNROLL
  for (int i = 0; i < N * N; ++i) lu[i] = Min[i];
  RPGO_UNROLL
  for (int i = 0; i < N; ++i) perm[i] = i;
  RPGO_UNROLL
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double biggest = fabs(lu[k * N + k]);
    RPGO_UNROLL
    for (int i = k + 1; i < N; ++i) {
      const double a = fabs(lu[i * N + k]);
      if (a > biggest) { biggest = a; piv = i; }
    }
    if (biggest != 0.0) {
      /* exchange rows k and piv with value selects only (a data-dependent row index would push the
       * whole matrix into local memory) */
      RPGO_UNROLL
      for (int j = 0; j < N; ++j) {
        const double oldk = lu[k * N + j];
        double newk = oldk;
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) {
          const bool sw = (piv == i);
          const double vi = lu[i * N + j];
          newk = sw ? vi : newk;
          lu[i * N + j] = sw ? oldk : vi;
        }
        lu[k * N + j] = newk;
      }
      {
        const int oldk = perm[k];
        int newk = oldk;
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) {
          const bool sw = (piv == i);
          const int vi = perm[i];
          newk = sw ? vi : newk;
          perm[i] = sw ? oldk : vi;
        }
        perm[k] = newk;
      }
      const double pv = lu[k * N + k];
      RPGO_UNROLL
      for (int i = k + 1; i < N; ++i) lu[i * N + k] = lu[i * N + k] / pv;
    }
    RPGO_UNROLL
    for (int i = k + 1; i < N; ++i) {
      RPGO_UNROLL
      for (int j = k + 1; j < N; ++j) lu[i * N + j] = fma(-lu[i * N + k], lu[k * N + j], lu[i * N + j]);
    }
  }
  double rdiag[N];
  RPGO_UNROLL
  for (int i = 0; i < N; ++i) rdiag[i] = 1.0 / lu[i * N + i];
  double w[N]; /* w[c] = sum_i v[i] * inv[i][c] */
  RPGO_UNROLL
  for (int c = 0; c < N; ++c) {
    double x[N];
    RPGO_UNROLL
    for (int i = 0; i < N; ++i) x[i] = (perm[i] == c) ? 1.0 : 0.0;
    RPGO_UNROLL
    for (int i = 0; i < N; ++i) {
      const double b = x[i];
      RPGO_UNROLL
      for (int r = i + 1; r < N; ++r) x[r] = fma(-b, lu[r * N + i], x[r]);
    }
    RPGO_UNROLL
    for (int i = N - 1; i >= 0; --i) {
      const double b = x[i] * rdiag[i];
      x[i] = b;
      RPGO_UNROLL
      for (int r = 0; r < i; ++r) x[r] = fma(-b, lu[r * N + i], x[r]);
    }
    double acc = v[0] * x[0];
    RPGO_UNROLL
    for (int i = 1; i < N; ++i) acc = fma(v[i], x[i], acc);
    w[c] = acc;
  }
  double q = w[0] * v[0];
  RPGO_UNROLL
  for (int j = 1; j < N; ++j) q = fma(w[j], v[j], q);
  return q;
}

template <int N>
RPGO_FN double quad_form_inv_fast(const double* Min, const double* v) {
  double lu[N * N];
  double rdiag[N];
  int perm[N];
  RPGO_UNROLL
  for (int i = 0; i < N * N; ++i) lu[i] = Min[i];
  RPGO_UNROLL
  for (int i = 0; i < N; ++i) perm[i] = i;
  RPGO_UNROLL
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double biggest = fabs(lu[k * N + k]);
    RPGO_UNROLL
    for (int i = k + 1; i < N; ++i) {
      const double a = fabs(lu[i * N + k]);
      if (a > biggest) { biggest = a; piv = i; }
    }
    if (biggest != 0.0) {
      /* exchange rows k and piv with value selects only (a data-dependent row index would push the
       * whole matrix into local memory) */
      RPGO_UNROLL
      for (int j = 0; j < N; ++j) {
        const double oldk = lu[k * N + j];
        double newk = oldk;
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) {
          const bool sw = (piv == i);
          const double vi = lu[i * N + j];
          newk = sw ? vi : newk;
          lu[i * N + j] = sw ? oldk : vi;
        }
        lu[k * N + j] = newk;
      }
      {
        const int oldk = perm[k];
        int newk = oldk;
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) {
          const bool sw = (piv == i);
          const int vi = perm[i];
          newk = sw ? vi : newk;
          perm[i] = sw ? oldk : vi;
        }
        perm[k] = newk;
      }
      const double pv = lu[k * N + k];
      const double rp = 1.0 / pv; /* shared by this column's divisions and by the back substitution */
      rdiag[k] = rp;
      if (rcp_safe(pv)) {
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) lu[i * N + k] = div_by(lu[i * N + k], pv, rp);
      } else {
        RPGO_UNROLL
        for (int i = k + 1; i < N; ++i) lu[i * N + k] = slow_div(lu[i * N + k], pv);
      }
    } else {
      rdiag[k] = 1.0 / lu[k * N + k]; /* zero pivot column: 1/0 exactly as the plain path computes it */
    }
    RPGO_UNROLL
    for (int i = k + 1; i < N; ++i) {
      RPGO_UNROLL
      for (int j = k + 1; j < N; ++j) lu[i * N + j] = fma(-lu[i * N + k], lu[k * N + j], lu[i * N + j]);
    }
  }
  /* inverse = U^-1 (L^-1 P).  Column c of P is e_p with perm[p] == c, so column c of the inverse is
   * U^-1 applied to column p of L^-1.  L^-1's columns have a static triangular zero pattern (the skipped
   * operations of the dense forward substitution are exact no-ops), so they are built without selects;
   * only the final 6 scalars are permuted back into natural order for the k-ordered dot product. */
  double wp[N]; /* wp[p] = sum_i v[i] * inv[i][perm[p]] */
  RPGO_UNROLL
  for (int p = 0; p < N; ++p) {
    double x[N];
    RPGO_UNROLL
    for (int i = 0; i < N; ++i) x[i] = (i == p) ? 1.0 : 0.0;
    RPGO_UNROLL
    for (int i = p; i < N; ++i) {
      const double b = x[i];
      RPGO_UNROLL
      for (int r = i + 1; r < N; ++r) x[r] = fma(-b, lu[r * N + i], x[r]);
    }
    RPGO_UNROLL
    for (int i = N - 1; i >= 0; --i) {
      const double b = x[i] * rdiag[i];
      x[i] = b;
      RPGO_UNROLL
      for (int r = 0; r < i; ++r) x[r] = fma(-b, lu[r * N + i], x[r]);
    }
    double acc = v[0] * x[0];
    RPGO_UNROLL
    for (int i = 1; i < N; ++i) acc = fma(v[i], x[i], acc);
    wp[p] = acc;
  }
  double w[N];
  RPGO_UNROLL
  for (int c = 0; c < N; ++c) {
    double t = wp[0];
    RPGO_UNROLL
    for (int p = 1; p < N; ++p) t = (perm[p] == c) ? wp[p] : t;
    w[c] = t;
  }
  double q = w[0] * v[0];
  RPGO_UNROLL
  for (int j = 1; j < N; ++j) q = fma(w[j], v[j], q);
  return q;
}


/* SO3::Logmap (GTSAM 4.0/4.1 constants) */
RPGO_FN void so3_logmap(const double* R, double* w) {
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

/* Pose3::Logmap / Pose2::Logmap.  v has Dim<D>::N entries in GTSAM tangent order. */
template <int D>
RPGO_FN void logmap(const Pose<D>& p, double* v) {
  if (D == 3) {
    double w[3];
    so3_logmap(p.m, w);
    const double T0 = p.m[9], T1 = p.m[10], T2 = p.m[11];
    const double t = sqrt(fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0])));
    v[0] = w[0]; v[1] = w[1]; v[2] = w[2];
    if (t < 1e-10) {
      v[3] = T0; v[4] = T1; v[5] = T2;
    } else {
      const double wx = w[0] / t, wy = w[1] / t, wz = w[2] / t;
      const double Tan = rpgo_tan(0.5 * t);
      /* W = [0 -wz wy; wz 0 -wx; -wy wx 0]; products in k-order with the zero terms skipped */
      const double WT0 = fma(wy, T2, (-wz) * T1);
      const double WT1 = fma(-wx, T2, wz * T0);
      const double WT2 = fma(wx, T1, (-wy) * T0);
      const double WWT0 = fma(wy, WT2, (-wz) * WT1);
      const double WWT1 = fma(-wx, WT2, wz * WT0);
      const double WWT2 = fma(wx, WT1, (-wy) * WT0);
      const double a = 0.5 * t;
      const double b = 1.0 - t / (2.0 * Tan);
      v[3] = (T0 - a * WT0) + b * WWT0;
      v[4] = (T1 - a * WT1) + b * WWT1;
      v[5] = (T2 - a * WT2) + b * WWT2;
    }
  } else {
    const double c = p.m[0], s = p.m[1], x = p.m[2], y = p.m[3];
    const double w = rpgo_atan2(s, c);
    if (fabs(w) < 1e-10) {
      v[0] = x; v[1] = y; v[2] = w;
    } else {
      const double c_1 = c - 1.0;
      const double det = fma(s, s, c_1 * c_1);
      const double ux = fma(s, y, c * x) - x;
      const double uy = fma(c, y, (-s) * x) - y;
      const double px = fma(-1.0, uy, 0.0 * ux);
      const double py = fma(0.0, uy, 1.0 * ux);
      const double f = w / det;
      v[0] = f * px; v[1] = f * py; v[2] = w;
    }
  }
}

/* ---- T<poseT>: pose + (covariance | node) ----------------------------------------------------- */
template <int D, int MODE>
struct PoseT {
  Pose<D> pose;
  double cov[MODE == MODE_PCM ? Dim<D>::N * Dim<D>::N : 1];
  int node;
  bool rot;
};

/* load an entry from memory with element stride `st` (1 = AoS; tile width = SoA) */
template <int D, int MODE>
RPGO_FN void load_entry(const double* __restrict__ e, int st, PoseT<D, MODE>& o) {
  RPGO_UNROLL
  for (int i = 0; i < Dim<D>::PS; ++i) o.pose.m[i] = e[i * st];
  if (MODE == MODE_PCM) {
    RPGO_UNROLL
    for (int i = 0; i < Dim<D>::N * Dim<D>::N; ++i) o.cov[i] = e[(Dim<D>::OFF_COV + i) * st];
    o.node = 0;
  } else {
    o.node = (int)e[Dim<D>::OFF_NODE * st];
  }
  o.rot = e[Dim<D>::OFF_ROT * st] != 0.0;
}
template <int D, int MODE>
RPGO_FN void store_entry(double* e, int st, const PoseT<D, MODE>& o) {
  RPGO_UNROLL
  for (int i = 0; i < Dim<D>::PS; ++i) e[i * st] = o.pose.m[i];
  if (MODE == MODE_PCM) {
    RPGO_UNROLL
    for (int i = 0; i < Dim<D>::N * Dim<D>::N; ++i) e[(Dim<D>::OFF_COV + i) * st] = o.cov[i];
  }
  e[Dim<D>::OFF_ROT * st] = o.rot ? 1.0 : 0.0;
  e[Dim<D>::OFF_NODE * st] = (double)o.node;
}

/* PoseWithCovariance / PoseWithNode ctor from a BetweenFactor (GeometryUtils.h:91-115, :221-237):
 * NaN rotation covariance => rotation_info = false and only the translation block is kept. */
template <int D, int MODE>
RPGO_FN void from_factor(const double* pose, const double* cov, PoseT<D, MODE>& o) {
  constexpr int N = Dim<D>::N, RD = Dim<D>::RD, TD = Dim<D>::TD;
  RPGO_UNROLL
  for (int i = 0; i < Dim<D>::PS; ++i) o.pose.m[i] = pose[i];
  double tr = cov[0];
  RPGO_UNROLL
  for (int i = 1; i < RD; ++i) tr = tr + cov[i * N + i];
  o.rot = !(tr != tr);
  if (MODE == MODE_PCM) {
    if (o.rot) {
      RPGO_UNROLL
      for (int i = 0; i < N * N; ++i) o.cov[i] = cov[i];
    } else {
      RPGO_UNROLL
      for (int i = 0; i < N * N; ++i) o.cov[i] = 0.0;
      RPGO_UNROLL
      for (int i = 0; i < TD; ++i) {
        RPGO_UNROLL
        for (int j = 0; j < TD; ++j) o.cov[(RD + i) * N + RD + j] = cov[(RD + i) * N + RD + j];
      }
    }
    o.node = 0;
  } else {
    o.node = 1;
  }
}

/* a.compose(b)  GeometryUtils.h:119-129 / :241-249.  Result may alias neither input. */
template <int D, int MODE>
RPGO_FN void pt_compose(const PoseT<D, MODE>& a, const PoseT<D, MODE>& b, PoseT<D, MODE>& o) {
  constexpr int N = Dim<D>::N;
  o.pose = compose<D>(a.pose, b.pose);
  if (MODE == MODE_PCM) {
    const Adj<D> H = adjoint<D>(inverse<D>(b.pose));
    hsht<D>(H, [&](int r, int c) { return a.cov[r * N + c]; }, o.cov);
    RPGO_UNROLL
    for (int i = 0; i < N * N; ++i) o.cov[i] = o.cov[i] + b.cov[i];
    o.node = 0;
  } else {
    o.node = a.node + b.node;
  }
  o.rot = a.rot && b.rot;
}

/* a.inverse(): pose inverted, covariance / node unchanged  (:133-139 / :253-261) */
template <int D, int MODE>
RPGO_FN void pt_inverse_inplace(PoseT<D, MODE>& a) {
  a.pose = inverse<D>(a.pose);
}

/* a.between(b)  GeometryUtils.h:143-170 / :265-273 */
template <int D, int MODE>
RPGO_FN void pt_between(const PoseT<D, MODE>& a, const PoseT<D, MODE>& b, PoseT<D, MODE>& o) {
  constexpr int N = Dim<D>::N;
  o.pose = between<D>(a.pose, b.pose);
  if (MODE == MODE_PCM) {
    {
      const Adj<D> H = adjoint<D>(inverse<D>(o.pose));
      hsht<D>(H, [&](int r, int c) { return a.cov[r * N + c]; }, o.cov);
    }
    RPGO_UNROLL
    for (int i = 0; i < N * N; ++i) o.cov[i] = b.cov[i] - o.cov[i];
    if (!llt_ok<N>(o.cov)) {
      const Adj<D> H = adjoint<D>(inverse<D>(between<D>(b.pose, a.pose)));
      hsht<D>(H, [&](int r, int c) { return b.cov[r * N + c]; }, o.cov);
      RPGO_UNROLL
      for (int i = 0; i < N * N; ++i) o.cov[i] = a.cov[i] - o.cov[i];
    }
    o.node = 0;
  } else {
    const int dn = b.node - a.node;
    o.node = dn < 0 ? -dn : dn;
  }
  o.rot = a.rot && b.rot;
}

/* mahalanobis_norm  GeometryUtils.h:172-186 */
template <int D, bool FAST = false>
RPGO_FN double mahalanobis(const PoseT<D, MODE_PCM>& a) {
  constexpr int N = Dim<D>::N, RD = Dim<D>::RD, TD = Dim<D>::TD;
  double lg[N];
  logmap<D>(a.pose, lg);
  double q;
  if (!a.rot) {
    double blk[TD * TD];
    RPGO_UNROLL
    for (int i = 0; i < TD; ++i) {
      RPGO_UNROLL
      for (int j = 0; j < TD; ++j) blk[i * TD + j] = a.cov[(RD + i) * N + RD + j];
    }
    q = FAST ? quad_form_inv_fast<TD>(blk, lg + RD) : quad_form_inv<TD>(blk, lg + RD);
  } else {
    q = FAST ? quad_form_inv_fast<N>(a.cov, lg) : quad_form_inv<N>(a.cov, lg);
  }
  return sqrt(q);
}

/* avg_trans_norm / avg_rot_norm  GeometryUtils.h:275-288 (incl. the Pose2 head/tail quirk) */
template <int D>
RPGO_FN void simple_norms(const PoseT<D, MODE_SIMPLE>& a, double& trans, double& rotn) {
  constexpr int N = Dim<D>::N, RD = Dim<D>::RD, TD = Dim<D>::TD;
  double lg[N];
  logmap<D>(a.pose, lg);
  double q = lg[N - TD] * lg[N - TD];
  RPGO_UNROLL
  for (int i = 1; i < TD; ++i) q = fma(lg[N - TD + i], lg[N - TD + i], q);
  trans = sqrt(q) / (double)a.node;
  if (!a.rot) {
    rotn = 0.0;
  } else {
    double r = lg[0] * lg[0];
    RPGO_UNROLL
    for (int i = 1; i < RD; ++i) r = fma(lg[i], lg[i], r);
    rotn = sqrt(r) / (double)a.node;
  }
}

/* thresholds as the kernels receive them */
struct Thresholds {
  double odom, lc;                 /* Pcm: Mahalanobis */
  double odom_trans, odom_rot;     /* PcmSimple */
  double dist_trans, dist_rot;
  double band;                     /* near-threshold flag half-width (1e-9) */
};

/* checkOdomConsistent / checkLoopConsistent  Pcm.h:564-596, 638-662.
 * Returns the decision; *dist = Mahalanobis (Pcm) or avg translation (Simple); *near = within band. */
template <int D, int MODE, bool FAST = false>
RPGO_FN bool check_consistent(const PoseT<D, MODE>& r, const Thresholds& th, bool odom, double* dist, bool* near) {
  if (MODE == MODE_PCM) {
    const double d = mahalanobis<D, FAST>(reinterpret_cast<const PoseT<D, MODE_PCM>&>(r));
    const double t = odom ? th.odom : th.lc;
    *dist = d;
    *near = fabs(d - t) < th.band;
    return d < t;
  } else {
    double tr, ro;
    simple_norms<D>(reinterpret_cast<const PoseT<D, MODE_SIMPLE>&>(r), tr, ro);
    const double tt = odom ? th.odom_trans : th.dist_trans;
    const double tro = odom ? th.odom_rot : th.dist_rot;
    *dist = tr;
    *near = (fabs(tr - tt) < th.band) || (fabs(ro - tro) < th.band);
    return tr < tt && ro < tro;
  }
}

/* areLoopsConsistent core  Pcm.h:703-717 given the four trajectory entries (after the key swap) and
 * the two closures.  Entries are read through (pointer, stride) so the same code serves global
 * AoS tables and shared-memory SoA tiles. */
template <int D, int MODE>
RPGO_FN bool pair_check(const double* Ta, int sa, const double* Tb, int sb, const double* lci, int sli,
                        const double* Tc, int sc, const double* Td, int sd, const double* lcj, int slj,
                        const Thresholds& th, double* dist, bool* near) {
  typedef PoseT<D, MODE> PT;
  PT x, y, z;
  {
    PT ea, ec;
    load_entry<D, MODE>(Ta, sa, ea);
    load_entry<D, MODE>(Tc, sc, ec);
    pt_between<D, MODE>(ea, ec, x); /* a_odom_c */
  }
  load_entry<D, MODE>(lcj, slj, y);
  pt_compose<D, MODE>(x, y, z); /* a_path_d = a_odom_c . c_lc_d */
  pt_inverse_inplace<D, MODE>(z);
  load_entry<D, MODE>(lci, sli, y);
  pt_compose<D, MODE>(z, y, x); /* d_path_b = a_path_d^-1 . a_lc_b */
  {
    PT eb, ed;
    load_entry<D, MODE>(Tb, sb, eb);
    load_entry<D, MODE>(Td, sd, ed);
    pt_between<D, MODE>(eb, ed, y); /* b_odom_d */
  }
  pt_compose<D, MODE>(x, y, z); /* loop */
  return check_consistent<D, MODE>(z, th, false, dist, near);
}


/* Landmark re-observation check (incrementLandmarkAdjMatrix, Pcm.h:810-832): observations (i -> l) and (j -> l)
 *   i_odom_j = Trajectory::getBetween(key_i, key_j);  loop = (i_odom_j . j_pose_l)^-1 . i_pose_l
 * `cross` selects getBetween's different-prefix path (GraphUtils.h:43-57) exactly as the reference runs it: inside
 * robot i's Trajectory the foreign keys do not exist, so operator[] yields default entries (E0) for them. */
template <int D, int MODE>
RPGO_FN bool landmark_pair_check(const double* Ei, const double* Ej, const double* Ea0, const double* E0, bool cross,
                                 const double* lci, const double* lcj, const Thresholds& th, double* dist, bool* near) {
  typedef PoseT<D, MODE> PT;
  PT x, y, z;
  if (!cross) {
    PT ei, ej;
    load_entry<D, MODE>(Ei, 1, ei);
    load_entry<D, MODE>(Ej, 1, ej);
    pt_between<D, MODE>(ei, ej, x);
  } else {
    PT ea0, ei, e0, pa, pb, pab, r;
    load_entry<D, MODE>(Ea0, 1, ea0);
    load_entry<D, MODE>(Ei, 1, ei);
    load_entry<D, MODE>(E0, 1, e0);
    pt_between<D, MODE>(ea0, ei, pa);  /* pose_a   = poses[a0].between(poses[key_a]) */
    pt_between<D, MODE>(e0, e0, pb);   /* pose_b   = poses[b0].between(poses[key_b]) : both default entries */
    pt_between<D, MODE>(ea0, e0, pab); /* pose_a0b0 = poses[a0].between(poses[b0]) */
    pt_inverse_inplace<D, MODE>(pa);
    pt_compose<D, MODE>(pa, pab, r);
    pt_compose<D, MODE>(r, pb, x);
  }
  load_entry<D, MODE>(lcj, 1, y);
  pt_compose<D, MODE>(x, y, z); /* i_path_l */
  pt_inverse_inplace<D, MODE>(z);
  load_entry<D, MODE>(lci, 1, y);
  pt_compose<D, MODE>(z, y, x); /* loop */
  return check_consistent<D, MODE>(x, th, false, dist, near);
}

/* ------------------------------------------------------------------------------------------------
 * pair_check_v1: same arithmetic as pair_check (bit-identical results), restructured for the tiled
 * kernel:
 *   - both between() stages and all three compose() stages run through ONE loop body each
 *     (#pragma unroll 1) so the kernel's hot code stays small enough for the instruction cache;
 *   - between()'s "LLT failed -> recompute the other way round" branch (GeometryUtils.h:150-161) is made
 *     warp-convergent: LLT fails at pivot 0 iff cov(0,0) <= 0 (no arithmetic precedes that test), so the
 *     (0,0) element is probed first (21 fused ops) and each lane picks its direction BEFORE the one full
 *     H S H^T it needs; lanes of both kinds then execute the same instructions on per-lane-selected
 *     operands.  Only a failure at a later pivot (numerically indefinite input) takes a divergent path;
 *   - b_odom_d is computed first and parked in `scr` (ENTRY doubles, element stride ss: per-thread
 *     scratch in shared memory) so that a single running covariance lives in registers.
 * Only MODE_PCM uses this path.
 * ---------------------------------------------------------------------------------------------- */
template <int D>
RPGO_FN void load_pose(const double* __restrict__ e, int st, Pose<D>& p) {
  RPGO_UNROLL
  for (int i = 0; i < Dim<D>::PS; ++i) p.m[i] = e[i * st];
}

/* (H S H^T)(0,0) with exactly the operations hsht() performs for that element */
template <int D>
RPGO_FN double hsht00(const Adj<D>& H, const double* __restrict__ S, int st) {
  constexpr int N = Dim<D>::N;
  const double* h = H.h;
  double t[3];
  RPGO_UNROLL
  for (int j = 0; j < 3; ++j) t[j] = dot3(h[0], h[1], h[2], S[(0 * N + j) * st], S[(1 * N + j) * st], S[(2 * N + j) * st]);
  return dot3(t[0], t[1], t[2], h[0], h[1], h[2]);
}

template <int D>
RPGO_FN bool pair_check_v1(const double* Ta, int sa, const double* Tb, int sb, const double* lci, int sli,
                           const double* Tc, int sc, const double* Td, int sd, const double* lcj, int slj,
                           double* scr, int ss, const Thresholds& th, double* dist, bool* near) {
  constexpr int N = Dim<D>::N, NN = N * N, OC = Dim<D>::OFF_COV, OR = Dim<D>::OFF_ROT;
  PoseT<D, MODE_PCM> x; /* running value of the chain */
  bool rot_chain = true;

  /* ---- the two between() stages: s = 0: b_odom_d -> scratch, s = 1: a_odom_c -> x ---- */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int s = 0; s < 2; ++s) {
    const double* pa = s == 0 ? Tb : Ta;
    const int sta = s == 0 ? sb : sa;
    const double* pb = s == 0 ? Td : Tc;
    const int stb = s == 0 ? sd : sc;
    Pose<D> A, B;
    load_pose<D>(pa, sta, A);
    load_pose<D>(pb, stb, B);
    const Pose<D> P = between<D>(A, B);                 /* result pose (both directions keep it) */
    Adj<D> H = adjoint<D>(inverse<D>(P));
    /* probe: forward covariance element (0,0) = cov_b(0,0) - (H cov_a H^T)(0,0) */
    const double c00 = pb[OC * stb] - hsht00<D>(H, pa + OC * sta, sta);
    const bool swapped = c00 <= 0.0;                    /* LLT pivot 0 fails  <=>  recompute the other way */
    if (swapped) H = adjoint<D>(inverse<D>(between<D>(B, A)));
    const double* pS = swapped ? pb : pa;               /* covariance that is propagated */
    const int stS = swapped ? stb : sta;
    const double* pT = swapped ? pa : pb;               /* covariance it is subtracted from */
    const int stT = swapped ? sta : stb;
    double S[NN];
    RPGO_UNROLL
    for (int i = 0; i < NN; ++i) S[i] = pS[(OC + i) * stS];
    hsht<D>(H, [&](int r, int c) { return S[r * N + c]; }, x.cov);
    RPGO_UNROLL
    for (int i = 0; i < NN; ++i) x.cov[i] = pT[(OC + i) * stT] - x.cov[i];
    if (!swapped) {
      if (!llt_ok_fast<N>(x.cov)) {
        /* failure at a later pivot: rare, divergent slow path, identical to the reference's order */
        const Adj<D> H2 = adjoint<D>(inverse<D>(between<D>(B, A)));
        RPGO_UNROLL
        for (int i = 0; i < NN; ++i) S[i] = pb[(OC + i) * stb];
        hsht<D>(H2, [&](int r, int c) { return S[r * N + c]; }, x.cov);
        RPGO_UNROLL
        for (int i = 0; i < NN; ++i) x.cov[i] = pa[(OC + i) * sta] - x.cov[i];
      }
    }
    x.pose = P;
    x.rot = (pa[OR * sta] != 0.0) && (pb[OR * stb] != 0.0);
    x.node = 0;
    if (s == 0) store_entry<D, MODE_PCM>(scr, ss, x);
  }

  /* ---- the three compose() stages: . c_lc_d, (inverse) . a_lc_b, . b_odom_d ---- */
  rot_chain = x.rot;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int t = 0; t < 3; ++t) {
    const double* po = t == 0 ? lcj : (t == 1 ? lci : scr);
    const int sto = t == 0 ? slj : (t == 1 ? sli : ss);
    Pose<D> O;
    load_pose<D>(po, sto, O);
    const Adj<D> H = adjoint<D>(inverse<D>(O));
    double out[NN];
    hsht<D>(H, [&](int r, int c) { return x.cov[r * N + c]; }, out);
    RPGO_UNROLL
    for (int i = 0; i < NN; ++i) x.cov[i] = out[i] + po[(OC + i) * sto];
    x.pose = compose<D>(x.pose, O);
    rot_chain = rot_chain && (po[OR * sto] != 0.0);
    if (t == 0) x.pose = inverse<D>(x.pose);
  }
  x.rot = rot_chain;
  return check_consistent<D, MODE_PCM, true>(x, th, false, dist, near);
}

}  // namespace rpgo
