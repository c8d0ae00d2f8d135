// Host-side instantiation of the PRODUCT's math header (kimera-rpgo_b200/csrc/rpgo_math.cuh) so that the
// CPU test-suite can prove, without a GPU, that its block-structured arithmetic is bit-identical to the
// dense oracle.  Built by tests/test_cpu_math_parity.py with g++ -ffp-contract=off.
#include "rpgo_math.cuh"
#include "rpgo_pair_v2.cuh"

using namespace rpgo;

template <int D, int MODE>
static int pair_t(const double* Ta, const double* Tb, const double* lci, const double* Tc, const double* Td,
                  const double* lcj, const double* thr, double* dist, int* near) {
  Thresholds th;
  th.odom = thr[0]; th.lc = thr[1]; th.odom_trans = thr[2]; th.odom_rot = thr[3]; th.dist_trans = thr[4];
  th.dist_rot = thr[5]; th.band = 1e-9;
  bool nr;
  const bool ok = pair_check<D, MODE>(Ta, 1, Tb, 1, lci, 1, Tc, 1, Td, 1, lcj, 1, th, dist, &nr);
  *near = nr;
  return ok;
}
template <int D, int MODE>
static void compose_t(const double* a, const double* b, double* o) {
  PoseT<D, MODE> x, y, z;
  load_entry<D, MODE>(a, 1, x); load_entry<D, MODE>(b, 1, y);
  pt_compose<D, MODE>(x, y, z);
  store_entry<D, MODE>(o, 1, z);
}
template <int D, int MODE>
static void between_t(const double* a, const double* b, double* o) {
  PoseT<D, MODE> x, y, z;
  load_entry<D, MODE>(a, 1, x); load_entry<D, MODE>(b, 1, y);
  pt_between<D, MODE>(x, y, z);
  store_entry<D, MODE>(o, 1, z);
}
template <int D, int MODE>
static void factor_t(const double* pose, const double* cov, double* o) {
  PoseT<D, MODE> x;
  from_factor<D, MODE>(pose, cov, x);
  for (int i = 0; i < Dim<D>::ENTRY; ++i) o[i] = 0.0;
  store_entry<D, MODE>(o, 1, x);
}

template <int D>
static int pair_v1_t(const double* Ta, const double* Tb, const double* lci, const double* Tc, const double* Td,
                     const double* lcj, const double* thr, double* dist, int* near) {
  Thresholds th;
  th.odom = thr[0]; th.lc = thr[1]; th.odom_trans = thr[2]; th.odom_rot = thr[3]; th.dist_trans = thr[4];
  th.dist_rot = thr[5]; th.band = 1e-9;
  bool nr;
  double scr[64];
  const bool ok = pair_check_v1<D>(Ta, 1, Tb, 1, lci, 1, Tc, 1, Td, 1, lcj, 1, scr, 1, th, dist, &nr);
  *near = nr;
  return ok;
}

template <int D>
static int pair_v2_t(const double* Ta, const double* Tb, const double* lci, const double* Tc, const double* Td,
                     const double* lcj, const double* thr, double* dist, int* bad) {
  Thresholds th;
  th.odom = thr[0]; th.lc = thr[1]; th.odom_trans = thr[2]; th.odom_rot = thr[3]; th.dist_trans = thr[4];
  th.dist_rot = thr[5]; th.band = 1e-9;
  bool nr, bd;
  double scr[64];
  const bool ok = pair_check_v2<D>(Ta, 1, Tb, 1, lci, 1, Tc, 1, Td, 1, lcj, 1, scr, 1, th, dist, &nr, &bd);
  *bad = bd;
  return ok;
}

/* straight-line PcmSimple pair function on the tiled kernel's compact entries (pose, rotation_info, node) */
template <int D>
static int pair_simple_v2_t(const double* Ta, const double* Tb, const double* lci, const double* Tc, const double* Td,
                            const double* lcj, const double* thr, double* dist, int* near, int* bad, int exact) {
  Thresholds th;
  th.odom = thr[0]; th.lc = thr[1]; th.odom_trans = thr[2]; th.odom_rot = thr[3]; th.dist_trans = thr[4];
  th.dist_rot = thr[5]; th.band = 1e-9;
  const double* in[6] = {Ta, Tb, lci, Tc, Td, lcj};
  double c[6][SimpleEntry<D>::E];
  for (int k = 0; k < 6; ++k) {
    for (int i = 0; i < Dim<D>::PS; ++i) c[k][i] = in[k][i];
    c[k][SimpleEntry<D>::OFF_ROT] = in[k][Dim<D>::OFF_ROT];
    c[k][SimpleEntry<D>::OFF_NODE] = in[k][Dim<D>::OFF_NODE];
  }
  bool nr, bd = false;
  bool ok;
  if (exact) ok = pair_check_simple_exact<D>(c[0], 1, c[1], 1, c[2], 1, c[3], 1, c[4], 1, c[5], 1, th, dist, &nr);
  else ok = pair_check_simple_v2<D>(c[0], 1, c[1], 1, c[2], 1, c[3], 1, c[4], 1, c[5], 1, th, dist, &nr, &bd);
  *near = nr;
  *bad = bd;
  return ok;
}

#define DISPATCH(d, m, F, ...)                                   \
  ((d) == 3 ? ((m) == 0 ? F<3, 0>(__VA_ARGS__) : F<3, 1>(__VA_ARGS__)) \
            : ((m) == 0 ? F<2, 0>(__VA_ARGS__) : F<2, 1>(__VA_ARGS__)))

extern "C" {
/* returns the number of mismatches between div_by(a, x, 1/x) and a / x over n operand pairs */
long long shim_div_by_mismatches(long long n, const double* a, const double* x) {
  long long bad = 0;
  for (long long i = 0; i < n; ++i) {
    const double r = 1.0 / x[i];
    const double q = rcp_safe(x[i]) ? div_by(a[i], x[i], r) : a[i] / x[i];
    const double w = a[i] / x[i];
    if (!(q == w) && !(q != q && w != w)) ++bad;
  }
  return bad;
}
int shim_entry_size(int d) { return d == 3 ? Dim<3>::ENTRY : Dim<2>::ENTRY; }
int shim_pair_check(int d, int mode, const double* Ta, const double* Tb, const double* lci, const double* Tc,
                    const double* Td, const double* lcj, const double* thr, double* dist, int* near) {
  return DISPATCH(d, mode, pair_t, Ta, Tb, lci, Tc, Td, lcj, thr, dist, near);
}
int shim_pair_check_v1(int d, const double* Ta, const double* Tb, const double* lci, const double* Tc,
                       const double* Td, const double* lcj, const double* thr, double* dist, int* near) {
  return d == 3 ? pair_v1_t<3>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, near) : pair_v1_t<2>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, near);
}
int shim_pair_check_v2(int d, const double* Ta, const double* Tb, const double* lci, const double* Tc,
                       const double* Td, const double* lcj, const double* thr, double* dist, int* bad) {
  return d == 3 ? pair_v2_t<3>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, bad) : pair_v2_t<2>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, bad);
}
int shim_pair_check_simple_v2(int d, const double* Ta, const double* Tb, const double* lci, const double* Tc, const double* Td,
                              const double* lcj, const double* thr, double* dist, int* near, int* bad, int exact) {
  return d == 3 ? pair_simple_v2_t<3>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, near, bad, exact)
                : pair_simple_v2_t<2>(Ta, Tb, lci, Tc, Td, lcj, thr, dist, near, bad, exact);
}
void shim_compose(int d, int mode, const double* a, const double* b, double* o) { DISPATCH(d, mode, compose_t, a, b, o); }
void shim_between(int d, int mode, const double* a, const double* b, double* o) { DISPATCH(d, mode, between_t, a, b, o); }
void shim_from_factor(int d, int mode, const double* pose, const double* cov, double* o) {
  DISPATCH(d, mode, factor_t, pose, cov, o);
}
}
