/* pcm_oracle.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the PCM outlier-rejection hot path of MIT-SPARK/Kimera-RPGO.
 * Nothing in the product (kimera-rpgo_b200/, include/) may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs load the shared object built from it (oracle/Makefile -> oracle/liboracle.so).
 *
 * What it restates (reference file:line, all under /root/reference):
 *   Pcm::removeOutliers            include/KimeraRPGO/outlier/Pcm.h:148-281
 *   Pcm::parseAndIncrementAdjMatrix                             Pcm.h:414-502
 *   Pcm::updateOdom                                             Pcm.h:516-557
 *   Pcm::isOdomConsistent / checkOdomConsistent                 Pcm.h:564-629
 *   Pcm::areLoopsConsistent / checkLoopConsistent               Pcm.h:638-718
 *   Pcm::incrementAdjMatrix                                     Pcm.h:725-768
 *   Pcm::findInliers / findInliersIncremental                   Pcm.h:851-970
 *   Pcm::buildGraphToOptimize                                   Pcm.h:977-1005
 *   Pcm::removeLastLoopClosure / ignore / revive                Pcm.h:299-383
 *   Trajectory::getBetween         include/KimeraRPGO/utils/GraphUtils.h:37-58
 *   findMaxCliqueHeu[Incremental], findMaxClique   src/utils/GraphUtils.cpp:9-44
 *   CGraphIO::ReadEigenAdjacencyMatrix   include/KimeraRPGO/max_clique_finder/graphIO.cpp:188-230
 *   FMC::maxCliqueHeu[Incremental]       include/KimeraRPGO/max_clique_finder/findCliqueHeu.cpp:32-209
 *   FMC::maxClique / maxCliqueHelper     include/KimeraRPGO/max_clique_finder/findClique.cpp:31-145
 * with the arithmetic of oracle_math.h.  Data structures deliberately mirror the
 * reference (dense double adjacency + distance matrices re-allocated and copied per
 * closure, std::map trajectory) so that timing it gives a "reference-shaped" CPU baseline;
 * orc_set_reference_shaped(h, 0) switches the O(n^2)-per-closure copy off.
 *
 * Parity pinning: see tests/test_oracle_golden.py (reference test expectations) and
 * tests/test_clique_ref.py (bit-equality with the reference's own FMC sources compiled
 * into oracle/_ref/).  Landmark (special-symbol) observations: Pcm.h:207-220, :437-455, :775-844.
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <map>
#include <unordered_map>
#include <vector>

#include "oracle_math.h"

namespace {

enum { MODE_PCM = 0, MODE_SIMPLE = 1 };

/* T<poseT>: PoseWithCovariance (mode 0) or PoseWithNode (mode 1) */
struct OT {
  opwc c;
  opwn s;
};

inline OT ot_default(int d, int mode) {
  OT r;
  memset(&r, 0, sizeof(r));
  if (mode == MODE_PCM) o_pwc_default(d, &r.c); else o_pwn_default(d, &r.s);
  return r;
}
inline OT ot_from_factor(int d, int mode, const double* pose, const double* cov) {
  OT r;
  memset(&r, 0, sizeof(r));
  if (mode == MODE_PCM) o_pwc_from_factor(d, pose, cov, &r.c); else o_pwn_from_factor(d, pose, cov, &r.s);
  return r;
}
inline OT ot_compose(int d, int mode, const OT& a, const OT& b) {
  OT r;
  memset(&r, 0, sizeof(r));
  if (mode == MODE_PCM) o_pwc_compose(d, &a.c, &b.c, &r.c); else o_pwn_compose(d, &a.s, &b.s, &r.s);
  return r;
}
inline OT ot_inverse(int d, int mode, const OT& a) {
  OT r;
  memset(&r, 0, sizeof(r));
  if (mode == MODE_PCM) o_pwc_inverse(d, &a.c, &r.c); else o_pwn_inverse(d, &a.s, &r.s);
  return r;
}
inline OT ot_between(int d, int mode, const OT& a, const OT& b) {
  OT r;
  memset(&r, 0, sizeof(r));
  if (mode == MODE_PCM) o_pwc_between(d, &a.c, &b.c, &r.c); else o_pwn_between(d, &a.s, &b.s, &r.s);
  return r;
}

inline unsigned char key_chr(uint64_t k) { return (unsigned char)(k >> 56); } /* gtsam::Symbol::chr() */

/* ---- clique finder restatement ---------------------------------------------------- */
struct Csr {
  std::vector<int> vtx, edg;
  int max_deg = -1;
};
/* graphIO.cpp:188-230 — lower triangle only (i >= j), nonzero => edge; lists come out ascending */
Csr csr_from_dense(int n, const double* adj) {
  std::vector<std::vector<int>> nl(n);
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) {
      if (i == j || adj[(size_t)i * n + j] == 0) continue;
      nl[i].push_back(j);
      nl[j].push_back(i);
    }
  Csr g;
  g.vtx.push_back(0);
  for (int i = 0; i < n; ++i) {
    g.edg.insert(g.edg.end(), nl[i].begin(), nl[i].end());
    g.vtx.push_back((int)g.edg.size());
  }
  for (int i = 0; i < n; ++i) g.max_deg = std::max(g.max_deg, g.vtx[i + 1] - g.vtx[i]);
  return g;
}
inline int deg(const Csr& g, int v) { return g.vtx[v + 1] - g.vtx[v]; }

/* findCliqueHeu.cpp:32-117 (first = 0, maxClq0 = -1) and :119-209 (first = n - num_new, maxClq0 = prev).
 * Returns maxClq; *out is the WHOLE scratch vector v_i_S at the time of the last improvement
 * (length max_deg + 1), exactly as the reference returns it. */
int heu_restated(const Csr& g, int first, int maxClq0, std::vector<int>* out) {
  const int n = (int)g.vtx.size() - 1;
  int maxClq = maxClq0;
  std::vector<int> S(g.max_deg + 1, 0), S1(g.max_deg + 1, 0);
  for (int v = first; v < n; ++v) {
    if (maxClq > deg(g, v)) continue; /* pruning 1 */
    int pos = 0;
    S[pos++] = v;
    for (int j = g.vtx[v]; j < g.vtx[v + 1]; ++j)
      if (maxClq <= deg(g, g.edg[j])) S[pos++] = g.edg[j]; /* pruning 3 */
    int icc = 0;
    while (pos > 0) {
      icc++;
      const int pick = S[pos - 1];
      int pos1 = 0;
      for (int j = 0; j < pos; ++j)
        for (int k = g.vtx[pick]; k < g.vtx[pick + 1]; ++k)
          if (S[j] == g.edg[k] && maxClq <= deg(g, g.edg[k])) { /* pruning 5 (no-op) */
            S1[pos1++] = S[j];
            break;
          }
      for (int j = 0; j < pos1; ++j) S[j] = S1[j];
      pos = pos1;
    }
    if (maxClq < icc) {
      *out = S;
      maxClq = icc;
    }
  }
  return maxClq;
}

/* findClique.cpp:31-77 */
void exact_helper(const Csr& g, std::vector<int>* U, size_t size, size_t* maxClq, std::vector<int>* inter) {
  if (U->empty()) {
    if (size > *maxClq) {
      *maxClq = size;
      inter->clear();
    }
    return;
  }
  std::vector<int> U_new;
  while (!U->empty()) {
    if (size + U->size() <= *maxClq) return;
    const int index = U->back();
    U->pop_back();
    for (int j = g.vtx[index]; j < g.vtx[index + 1]; ++j)
      if ((size_t)deg(g, g.edg[j]) >= *maxClq)
        for (size_t i = 0; i < U->size(); ++i)
          if (g.edg[j] == (*U)[i]) U_new.push_back(g.edg[j]);
    const size_t prev = *maxClq;
    exact_helper(g, &U_new, size + 1, maxClq, inter);
    if (*maxClq > prev) inter->push_back(index);
    U_new.clear();
  }
}
/* findClique.cpp:80-145 */
int exact_restated(const Csr& g, size_t l_bound, std::vector<int>* out) {
  const int n = (int)g.vtx.size() - 1;
  size_t maxClq = l_bound;
  std::vector<int> U, inter;
  std::vector<int> seen(n, 0);
  for (int i = n - 1; i >= 0; --i) {
    seen[i] = 1;
    const size_t prev = maxClq;
    U.clear();
    if ((size_t)deg(g, i) < maxClq) continue;
    for (int j = g.vtx[i]; j < g.vtx[i + 1]; ++j)
      if (seen[g.edg[j]] != 1 && (size_t)deg(g, g.edg[j]) >= maxClq) U.push_back(g.edg[j]);
    exact_helper(g, &U, 1, &maxClq, &inter);
    if (maxClq > prev) {
      inter.push_back(i);
      *out = inter;
    }
    inter.clear();
  }
  return (int)maxClq;
}

/* src/utils/GraphUtils.cpp:19-27 */
int find_max_clique_heu(int n, const double* adj, std::vector<int>* out) {
  Csr g = csr_from_dense(n, adj);
  return heu_restated(g, 0, -1, out);
}
/* src/utils/GraphUtils.cpp:30-44 */
int find_max_clique_heu_incremental(int n, const double* adj, size_t num_new, size_t prev, std::vector<int>* out) {
  Csr g = csr_from_dense(n, adj);
  /* findCliqueHeu.cpp:141: the first candidate index is computed in size_t; with more new closures than vertices (only
   * possible when the pairwise check is disabled and the matrix stays 1x1) it wraps and the loop never runs */
  if (num_new > (size_t)n) return 0;
  const int r = heu_restated(g, n - (int)num_new, (int)prev, out);
  if ((size_t)r > prev) return r;
  return 0;
}

/* ---- pipeline state ------------------------------------------------------------------ */
struct Factor {
  int type; /* 0 = BetweenFactor<poseT>, 1 = PriorFactor<poseT>, 2 = any other factor */
  uint64_t k1, k2;
  double pose[12];
  double cov[36];
  int64_t id;
};

struct Measurements { /* TypeUtils.h:16-33 */
  std::vector<Factor> factors;
  std::vector<int64_t> consistent; /* factor ids */
  std::vector<double> adj, dist;   /* dense n x n doubles, row-major */
  int n_adj = 1;                   /* matrices start 1x1 zero */
  Measurements() : adj(1, 0.0), dist(1, 0.0) {}
};

struct ObsId {
  unsigned char a, b;
  bool operator<(const ObsId& o) const { return a != o.a ? a < o.a : b < o.b; }
};
inline ObsId make_obs(unsigned char x, unsigned char y) { /* unordered pair, TypeUtils.h:44-58 */
  ObsId o;
  o.a = std::min(x, y);
  o.b = std::max(x, y);
  return o;
}

struct Oracle {
  int d, mode;
  double odom_threshold, lc_threshold, odom_trans_threshold, odom_rot_threshold, dist_trans_threshold,
      dist_rot_threshold;
  bool incremental;
  bool odom_check = true, loop_check = true;
  bool reference_shaped = true;
  double band = 1e-9;

  std::vector<int64_t> nfg_odom, nfg_special;
  std::vector<Factor> special_factors;
  std::map<ObsId, Measurements> loop_closures;       /* reference: unordered_map (iteration order differs) */
  std::vector<ObsId> group_order;                     /* first-seen order, used for output ordering */
  std::map<unsigned char, std::map<uint64_t, OT>> traj; /* odom_trajectories_ */
  std::map<uint64_t, opose> values;
  std::vector<ObsId> lc_in_order;
  std::vector<unsigned char> ignored;
  std::vector<unsigned char> special_symbols;          /* Pcm.h:106 */
  std::map<uint64_t, Measurements> landmarks;          /* Pcm.h:109 (reference: unordered_map) */
  std::vector<uint64_t> landmark_order;                /* first-seen order, for output ordering */
  bool is_special(unsigned char c) const {             /* Pcm.h:506-511 */
    return std::find(special_symbols.begin(), special_symbols.end(), c) != special_symbols.end();
  }
  size_t total_lc = 0, total_good_lc = 0;
  int64_t next_id = 0;
  std::vector<int64_t> output;
  std::vector<std::pair<int64_t, int64_t>> flagged; /* near-threshold (factor id, factor id) */
  uint64_t pair_checks = 0;

  OT get_between(unsigned char prefix, uint64_t ka, uint64_t kb) { /* GraphUtils.h:37-58 */
    auto& poses = traj[prefix]; /* operator[] creates the trajectory, as the reference does */
    auto get = [&](uint64_t k) -> OT& {
      auto it = poses.find(k);
      if (it == poses.end()) it = poses.emplace(k, ot_default(d, mode)).first; /* operator[] inserts default */
      return it->second;
    };
    if (key_chr(ka) == key_chr(kb)) {
      OT a = get(ka);
      OT b = get(kb);
      return ot_between(d, mode, a, b);
    }
    const uint64_t a0 = (uint64_t)key_chr(ka) << 56, b0 = (uint64_t)key_chr(kb) << 56;
    OT pa = ot_between(d, mode, get(a0), get(ka));
    OT pb = ot_between(d, mode, get(b0), get(kb));
    OT pab = ot_between(d, mode, get(a0), get(b0));
    OT r = ot_compose(d, mode, ot_inverse(d, mode, pa), pab);
    return ot_compose(d, mode, r, pb);
  }

  void update_odom(const Factor& f) { /* Pcm.h:516-557 */
    nfg_odom.push_back(f.id);
    const unsigned char prefix = key_chr(f.k2);
    OT delta = ot_from_factor(d, mode, f.pose, f.cov);
    if (traj.find(prefix) == traj.end()) {
      OT init = ot_default(d, mode);
      auto it = values.find(f.k1);
      opose p0;
      if (it != values.end()) p0 = it->second; else o_pose_identity(d, &p0);
      if (mode == MODE_PCM) init.c.pose = p0; else init.s.pose = p0;
      traj[prefix][f.k1] = init;
    }
    auto& poses = traj[prefix];
    auto it = poses.find(f.k1);
    if (it == poses.end()) it = poses.emplace(f.k1, ot_default(d, mode)).first;
    OT prev = it->second;
    poses[f.k2] = ot_compose(d, mode, prev, delta);
  }

  bool check(const OT& r, bool odom, double* dist) { /* Pcm.h:564-596, 638-662 */
    if (mode == MODE_PCM) {
      *dist = o_pwc_mahalanobis(d, &r.c);
      return *dist < (odom ? odom_threshold : lc_threshold);
    }
    *dist = o_pwn_avg_trans(d, &r.s);
    const double rot = o_pwn_avg_rot(d, &r.s);
    return *dist < (odom ? odom_trans_threshold : dist_trans_threshold) &&
           rot < (odom ? odom_rot_threshold : dist_rot_threshold);
  }

  bool is_odom_consistent(const Factor& f, double* dist) { /* Pcm.h:604-629 */
    if (!odom_check) return true;
    OT pij = get_between(key_chr(f.k1), f.k1, f.k2);
    OT pji = ot_inverse(d, mode, ot_from_factor(d, mode, f.pose, f.cov));
    OT r = ot_compose(d, mode, pij, pji);
    return check(r, true, dist);
  }

  bool are_loops_consistent(const Factor& ab, const Factor& cd, double* dist) { /* Pcm.h:670-718 */
    if (!loop_check) return true;
    uint64_t ka = ab.k1, kb = ab.k2, kc = cd.k1, kd = cd.k2;
    OT a_lc_b = ot_from_factor(d, mode, ab.pose, ab.cov);
    OT c_lc_d = ot_from_factor(d, mode, cd.pose, cd.cov);
    if (key_chr(ka) != key_chr(kc)) std::swap(kc, kd); /* keys swapped, measurement NOT inverted :691-698 */
    OT a_odom_c = get_between(key_chr(ka), ka, kc);
    OT b_odom_d = get_between(key_chr(kb), kb, kd);
    OT a_path_d = ot_compose(d, mode, a_odom_c, c_lc_d);
    OT d_path_b = ot_compose(d, mode, ot_inverse(d, mode, a_path_d), a_lc_b);
    OT loop = ot_compose(d, mode, d_path_b, b_odom_d);
    pair_checks++;
    return check(loop, false, dist);
  }

  void note_band(double dist, int64_t fi, int64_t fj) {
    if (mode == MODE_PCM) {
      if (fabs(dist - lc_threshold) < band) flagged.push_back({fi, fj});
    }
  }

  void increment_adj(const ObsId& id, const Factor& f) { /* Pcm.h:725-768 */
    Measurements& m = loop_closures[id];
    const size_t n = m.factors.size();
    if (reference_shaped) {
      std::vector<double> nadj(n * n, 0.0), ndst(n * n, 0.0);
      if (n > 1) {
        for (size_t i = 0; i + 1 < n; ++i)
          for (size_t j = 0; j + 1 < n; ++j) {
            nadj[i * n + j] = m.adj[i * (n - 1) + j];
            ndst[i * n + j] = m.dist[i * (n - 1) + j];
          }
      }
      m.adj.swap(nadj);
      m.dist.swap(ndst);
    }
    m.n_adj = (int)n;
    if (n > 1) {
      for (size_t i = 0; i + 1 < n; ++i) {
        double dist = 0.0;
        const bool ok = are_loops_consistent(m.factors[i], f, &dist);
        note_band(dist, m.factors[i].id, f.id);
        if (reference_shaped) {
          m.dist[(n - 1) * n + i] = dist;
          m.dist[i * n + (n - 1)] = dist;
          if (ok) {
            m.adj[(n - 1) * n + i] = 1;
            m.adj[i * n + (n - 1)] = 1;
          }
        } else {
          tri_set(m, i, n - 1, ok ? 1.0 : 0.0, dist);
        }
      }
    }
  }

  /* non-reference-shaped storage: packed lower triangle in m.adj / m.dist (index j*(j-1)/2 + i, i<j) */
  static void tri_set(Measurements& m, size_t i, size_t j, double a, double dist) {
    const size_t idx = j * (j - 1) / 2 + i;
    if (m.adj.size() <= idx) {
      m.adj.resize(std::max(idx + 1, m.adj.size() * 2), 0.0);
      m.dist.resize(m.adj.size(), 0.0);
    }
    m.adj[idx] = a;
    m.dist[idx] = dist;
  }
  void dense(const Measurements& m, std::vector<double>* adj, std::vector<double>* dist) const {
    const size_t n = (size_t)m.n_adj;
    if (reference_shaped) {
      *adj = std::vector<double>(m.adj.begin(), m.adj.begin() + n * n);
      if (dist) *dist = std::vector<double>(m.dist.begin(), m.dist.begin() + n * n);
      return;
    }
    adj->assign(n * n, 0.0);
    if (dist) dist->assign(n * n, 0.0);
    for (size_t j = 1; j < n; ++j)
      for (size_t i = 0; i < j; ++i) {
        const size_t idx = j * (j - 1) / 2 + i;
        (*adj)[i * n + j] = (*adj)[j * n + i] = m.adj[idx];
        if (dist) (*dist)[i * n + j] = (*dist)[j * n + i] = m.dist[idx];
      }
  }

  /* Pcm.h:775-844.  Observations are (pose key -> landmark key).  A malformed observation (landmark key in
   * front) makes the reference return before it stores the grown matrices; after that its next growth step
   * copies a wrongly sized block (undefined behaviour), so the oracle simply refuses the malformed factor. */
  void increment_landmark_adj(uint64_t lkey) {
    Measurements& m = landmarks[lkey];
    const size_t n = m.factors.size();
    std::vector<double> nadj(n * n, 0.0), ndst(n * n, 0.0);
    if (n > 1) {
      const size_t o = (size_t)m.n_adj;
      for (size_t i = 0; i < o && i + 1 < n; ++i)
        for (size_t j = 0; j < o && j + 1 < n; ++j) {
          nadj[i * n + j] = m.adj[i * o + j];
          ndst[i * n + j] = m.dist[i * o + j];
        }
      const Factor& fj = m.factors[n - 1];
      for (size_t i = 0; i + 1 < n; ++i) {
        const Factor& fi = m.factors[i];
        const uint64_t keyi = fi.k1, keyj = fj.k1;
        OT i_pose_l = ot_from_factor(d, mode, fi.pose, fi.cov);
        OT j_pose_l = ot_from_factor(d, mode, fj.pose, fj.cov);
        OT i_odom_j = get_between(key_chr(keyi), keyi, keyj); /* cross-prefix path of GraphUtils.h:43-57 included */
        OT i_path_l = ot_compose(d, mode, i_odom_j, j_pose_l);
        OT loop = ot_compose(d, mode, ot_inverse(d, mode, i_path_l), i_pose_l);
        double dist;
        const bool ok = check(loop, false, &dist);
        pair_checks++;
        note_band(dist, fi.id, fj.id);
        ndst[(n - 1) * n + i] = ndst[i * n + (n - 1)] = dist;
        if (ok) nadj[(n - 1) * n + i] = nadj[i * n + (n - 1)] = 1;
      }
    }
    m.adj.swap(nadj);
    m.dist.swap(ndst);
    m.n_adj = (int)n;
  }

  void parse_and_increment(const std::vector<Factor>& lcs, std::map<ObsId, size_t>* num_new) { /* Pcm.h:414-502 */
    for (const Factor& f : lcs) {
      if (values.find(f.k1) == values.end() || values.find(f.k2) == values.end()) continue; /* :431-435 */
      if (is_special(key_chr(f.k1)) || is_special(key_chr(f.k2))) { /* landmark re-observation :440-455 */
        const uint64_t lkey = is_special(key_chr(f.k1)) ? f.k1 : f.k2;
        if (f.k1 == lkey) continue; /* malformed (see increment_landmark_adj) */
        if (landmarks.find(lkey) == landmarks.end()) landmark_order.push_back(lkey);
        landmarks[lkey].factors.push_back(f);
        total_lc++;
        increment_landmark_adj(lkey);
        continue;
      }
      double odom_dist;
      bool ok;
      if (key_chr(f.k1) == key_chr(f.k2)) ok = is_odom_consistent(f, &odom_dist); else ok = true;
      if (!ok) continue; /* dropped entirely :487-493 */
      ObsId id = make_obs(key_chr(f.k1), key_chr(f.k2));
      (*num_new)[id]++;
      if (loop_closures.find(id) == loop_closures.end()) group_order.push_back(id);
      loop_closures[id].factors.push_back(f);
      lc_in_order.push_back(id);
      total_lc++;
      if (loop_check) increment_adj(id, f);
    }
  }

  void find_inliers() { /* Pcm.h:851-899 */
    total_good_lc = 0;
    for (auto& kv : loop_closures) {
      Measurements& m = kv.second;
      size_t num_inliers;
      if (loop_check) {
        std::vector<int> idx;
        m.consistent.clear();
        std::vector<double> adj;
        dense(m, &adj, nullptr);
        num_inliers = (size_t)find_max_clique_heu(m.n_adj, adj.data(), &idx);
        for (size_t i = 0; i < num_inliers; ++i) m.consistent.push_back(m.factors[idx[i]].id);
      } else {
        m.consistent.clear();
        for (auto& f : m.factors) m.consistent.push_back(f.id);
        num_inliers = m.factors.size();
      }
      total_good_lc += num_inliers;
    }
    landmark_inliers();
  }

  /* Pcm.h:878-895 (also :950-966 in the incremental variant): every landmark, full heuristic */
  void landmark_inliers() {
    for (auto& kv : landmarks) {
      Measurements& m = kv.second;
      std::vector<int> idx;
      const size_t n = (size_t)m.n_adj;
      std::vector<double> adj(m.adj.begin(), m.adj.begin() + n * n);
      const size_t k = (size_t)find_max_clique_heu((int)n, adj.data(), &idx);
      m.consistent.clear();
      for (size_t i = 0; i < k; ++i) m.consistent.push_back(m.factors[idx[i]].id);
      total_good_lc += k;
    }
  }

  void find_inliers_incremental(const std::map<ObsId, size_t>& num_new) { /* Pcm.h:906-970 */
    total_good_lc = 0;
    for (auto& kv : num_new) {
      Measurements& m = loop_closures[kv.first];
      std::vector<int> idx;
      const size_t prev = m.consistent.size();
      std::vector<double> adj;
      dense(m, &adj, nullptr);
      const size_t num_inliers = (size_t)find_max_clique_heu_incremental(m.n_adj, adj.data(), kv.second, prev, &idx);
      if (num_inliers > 0) {
        m.consistent.clear();
        for (size_t i = 0; i < num_inliers; ++i) m.consistent.push_back(m.factors[idx[i]].id);
      }
    }
    for (auto& kv : loop_closures) total_good_lc += kv.second.consistent.size();
    landmark_inliers();
  }

  void build_graph() { /* Pcm.h:977-1005: odom, special, consistent LCs of non-ignored groups */
    output.clear();
    output.insert(output.end(), nfg_odom.begin(), nfg_odom.end());
    output.insert(output.end(), nfg_special.begin(), nfg_special.end());
    for (const ObsId& id : group_order) {
      auto it = loop_closures.find(id);
      if (it == loop_closures.end()) continue;
      if (std::find(ignored.begin(), ignored.end(), id.a) != ignored.end()) continue;
      if (std::find(ignored.begin(), ignored.end(), id.b) != ignored.end()) continue;
      output.insert(output.end(), it->second.consistent.begin(), it->second.consistent.end());
    }
    for (uint64_t lk : landmark_order) { /* Pcm.h:996-1002 */
      auto it = landmarks.find(lk);
      if (it != landmarks.end()) output.insert(output.end(), it->second.consistent.begin(), it->second.consistent.end());
    }
  }

  bool remove_outliers(std::vector<Factor>& nf, const std::vector<std::pair<uint64_t, opose>>& nv) { /* Pcm.h:148-281 */
    std::map<uint64_t, int> new_keys;
    for (auto& kv : nv) {
      values[kv.first] = kv.second;
      new_keys[kv.first] = 1;
    }
    if (nf.empty()) return false;
    bool do_optimize = false;
    std::vector<Factor> lcs;
    for (Factor& f : nf) {
      f.id = next_id++;
      if (f.type == 0) {
        if (is_special(key_chr(f.k1)) || is_special(key_chr(f.k2))) { /* Pcm.h:180-188 */
          if (new_keys.count(f.k1) || new_keys.count(f.k2)) {
            /* FIRST_LANDMARK_OBSERVATION :207-220 */
            const uint64_t lkey = is_special(key_chr(f.k1)) ? f.k1 : f.k2;
            if (landmarks.find(lkey) == landmarks.end()) landmark_order.push_back(lkey);
            Measurements m;
            m.factors.push_back(f);
            m.consistent.push_back(f.id);
            landmarks[lkey] = m;
            total_lc++;
          } else if (f.k1 != f.k2) {
            lcs.push_back(f);
          }
        } else if (f.k1 + 1 == f.k2 && new_keys.count(f.k2)) {
          update_odom(f);
        } else {
          if (f.k1 != f.k2) lcs.push_back(f);
        }
      } else {
        nfg_special.push_back(f.id);
        special_factors.push_back(f);
        do_optimize = true;
      }
    }
    if (!lcs.empty()) {
      std::map<ObsId, size_t> num_new;
      parse_and_increment(lcs, &num_new);
      if (incremental) find_inliers_incremental(num_new); else find_inliers();
      do_optimize = true;
    }
    build_graph();
    return do_optimize;
  }

  /* Pcm.h:299-353.  Returns 1 and the removed edge keys, or 0 */
  int remove_last(const ObsId& id, uint64_t* k1, uint64_t* k2) {
    auto it = loop_closures.find(id);
    if (it == loop_closures.end()) return 0;
    Measurements& m = it->second;
    const size_t numLC = (size_t)m.n_adj;
    if (numLC <= 0) return 0;
    const size_t num_lc = m.factors.size();
    if (num_lc == 0) return 0; /* reference would index factors[-1]; guarded */
    *k1 = m.factors[num_lc - 1].k1;
    *k2 = m.factors[num_lc - 1].k2;
    const int64_t removed_id = m.factors[num_lc - 1].id;
    m.factors.pop_back();
    if (m.factors.size() < 2) {
      m.consistent.clear();
      for (auto& f : m.factors) m.consistent.push_back(f.id);
    } else if (!loop_check) {
      /* With the pairwise check disabled the reference's matrices stay 1x1 (Pcm.h:484-486), so Pcm.h:320-323 takes a
       * 0x0 block and findMaxCliqueHeu returns -1 as a size_t: undefined behaviour.  Defined here (and in the product's
       * host mirrors) as what findInliers does in that configuration (Pcm.h:870-873: every factor is an inlier); in
       * incremental mode the previous inlier set is kept, minus the removed factor. */
      if (incremental) {
        m.consistent.erase(std::remove(m.consistent.begin(), m.consistent.end(), removed_id), m.consistent.end());
      } else {
        m.consistent.clear();
        for (auto& f : m.factors) m.consistent.push_back(f.id);
      }
    } else {
      std::vector<double> adj, dist;
      dense(m, &adj, &dist);
      const size_t nn = numLC - 1;
      std::vector<double> a2(nn * nn), d2(nn * nn);
      for (size_t i = 0; i < nn; ++i)
        for (size_t j = 0; j < nn; ++j) {
          a2[i * nn + j] = adj[i * numLC + j];
          d2[i * nn + j] = dist[i * numLC + j];
        }
      if (reference_shaped) {
        m.adj = a2;
        m.dist = d2;
      }
      m.n_adj = (int)nn;
      std::vector<int> idx;
      const size_t k = (size_t)find_max_clique_heu((int)nn, a2.data(), &idx);
      m.consistent.clear();
      for (size_t i = 0; i < k; ++i) m.consistent.push_back(m.factors[idx[i]].id);
    }
    build_graph();
    return 1;
  }
};

}  // namespace

extern "C" {

/* thr = {odom_threshold, lc_threshold, odom_trans, odom_rot, dist_trans, dist_rot}  (SolverParams.h:33-56) */
void* orc_create(int d, int mode, const double* thr, int incremental) {
  Oracle* o = new Oracle();
  o->d = d;
  o->mode = mode;
  o->odom_threshold = thr[0];
  o->lc_threshold = thr[1];
  o->odom_trans_threshold = thr[2];
  o->odom_rot_threshold = thr[3];
  o->dist_trans_threshold = thr[4];
  o->dist_rot_threshold = thr[5];
  o->incremental = incremental != 0;
  /* Pcm.h:74-82 */
  if (thr[0] < 0 || thr[3] < 0 || thr[2] < 0) o->odom_check = false;
  if (thr[1] < 0 || thr[5] < 0 || thr[4] < 0) o->loop_check = false;
  return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
void orc_set_special_symbols(void* h, int n, const unsigned char* syms) {
  Oracle* o = (Oracle*)h;
  o->special_symbols.assign(syms, syms + n);
}
/* N4, multirobotValueInitialization front half (Pcm.h:1024-1055): T_w0_wi = T_w0_front . T_front_back . T_wi_back^-1 for
 * every consistent closure of group (r0, ri).  Returns the number written, or -1 when a trajectory key is missing
 * (the reference's .at() throws and the robot is skipped). */
int orc_frame_align(void* h, int r0, int ri, double* out) {
  Oracle* o = (Oracle*)h;
  auto it = o->loop_closures.find(make_obs((unsigned char)r0, (unsigned char)ri));
  if (it == o->loop_closures.end()) return -1;
  const int ps = o->d == 3 ? 12 : 4;
  int n = 0;
  for (int64_t fid : it->second.consistent) {
    const Factor* f = nullptr;
    for (const Factor& q : it->second.factors) if (q.id == fid) f = &q;
    uint64_t front = f->k1, back = f->k2;
    opose T_fb;
    memset(&T_fb, 0, sizeof(T_fb));
    memcpy(T_fb.m, f->pose, sizeof(double) * ps);
    if (key_chr(front) != (unsigned char)r0) {
      std::swap(front, back);
      opose inv;
      o_pose_inverse(o->d, &T_fb, &inv);
      T_fb = inv;
    }
    auto t0 = o->traj.find((unsigned char)r0), ti = o->traj.find((unsigned char)ri);
    if (t0 == o->traj.end() || ti == o->traj.end()) return -1;
    auto pf = t0->second.find(front), pb = ti->second.find(back);
    if (pf == t0->second.end() || pb == ti->second.end()) return -1;
    const opose& Tf = o->mode == MODE_PCM ? pf->second.c.pose : pf->second.s.pose;
    const opose& Tb = o->mode == MODE_PCM ? pb->second.c.pose : pb->second.s.pose;
    opose tmp, Tbinv, res;
    o_pose_compose(o->d, &Tf, &T_fb, &tmp);
    o_pose_inverse(o->d, &Tb, &Tbinv);
    o_pose_compose(o->d, &tmp, &Tbinv, &res);
    memcpy(out + (size_t)n * ps, res.m, sizeof(double) * ps);
    ++n;
  }
  return n;
}
/* getRobotOdomValues (Pcm.h:1074-1082): transform . pose for every trajectory entry, ascending key order */
int orc_robot_odom_values(void* h, int prefix, const double* transform, uint64_t* keys, double* poses) {
  Oracle* o = (Oracle*)h;
  auto t = o->traj.find((unsigned char)prefix);
  if (t == o->traj.end()) return -1;
  const int ps = o->d == 3 ? 12 : 4;
  opose T;
  o_pose_identity(o->d, &T);
  if (transform) memcpy(T.m, transform, sizeof(double) * ps);
  int n = 0;
  for (auto& kv : t->second) {
    const opose& P = o->mode == MODE_PCM ? kv.second.c.pose : kv.second.s.pose;
    opose r;
    o_pose_compose(o->d, &T, &P, &r);
    if (keys) keys[n] = kv.first;
    if (poses) memcpy(poses + (size_t)n * ps, r.m, sizeof(double) * ps);
    ++n;
  }
  return n;
}
int orc_num_landmarks(void* h) { return (int)((Oracle*)h)->landmark_order.size(); }
/* landmark l in first-seen order: key, number of observations, number of inliers */
void orc_landmark_info(void* h, int l, uint64_t* key, int* n, int* n_inliers) {
  Oracle* o = (Oracle*)h;
  const uint64_t k = o->landmark_order[l];
  const Measurements& m = o->landmarks[k];
  *key = k;
  *n = (int)m.factors.size();
  *n_inliers = (int)m.consistent.size();
}
int orc_landmark_adj(void* h, int l, unsigned char* adj, double* dist) {
  Oracle* o = (Oracle*)h;
  const Measurements& m = o->landmarks[o->landmark_order[l]];
  const size_t n = (size_t)m.n_adj;
  for (size_t i = 0; i < n * n; ++i) {
    if (adj) adj[i] = m.adj[i] != 0.0;
    if (dist) dist[i] = m.dist[i];
  }
  return (int)n;
}
void orc_landmark_ids(void* h, int l, long long* factor_ids, long long* inlier_ids) {
  Oracle* o = (Oracle*)h;
  const Measurements& m = o->landmarks[o->landmark_order[l]];
  if (factor_ids) for (size_t i = 0; i < m.factors.size(); ++i) factor_ids[i] = m.factors[i].id;
  if (inlier_ids) for (size_t i = 0; i < m.consistent.size(); ++i) inlier_ids[i] = m.consistent[i];
}
void orc_set_reference_shaped(void* h, int on) { ((Oracle*)h)->reference_shaped = on != 0; }

/* One removeOutliers() call.  Factors: type (0 between, 1 prior, 2 other), keys, pose (12 or 4
 * doubles each, stride ps), covariance (n*n doubles each).  Values: keys + poses. */
int orc_update(void* h, int nf, const int* types, const uint64_t* k1, const uint64_t* k2, const double* poses,
               const double* covs, int nv, const uint64_t* vkeys, const double* vposes) {
  Oracle* o = (Oracle*)h;
  const int ps = o->d == 3 ? 12 : 4, n = o_n(o->d);
  std::vector<Factor> fs(nf);
  for (int i = 0; i < nf; ++i) {
    memset(&fs[i], 0, sizeof(Factor));
    fs[i].type = types[i];
    fs[i].k1 = k1[i];
    fs[i].k2 = k2[i];
    memcpy(fs[i].pose, poses + (size_t)i * ps, sizeof(double) * ps);
    memcpy(fs[i].cov, covs + (size_t)i * n * n, sizeof(double) * n * n);
  }
  std::vector<std::pair<uint64_t, opose>> nvs(nv);
  for (int i = 0; i < nv; ++i) {
    nvs[i].first = vkeys[i];
    memset(&nvs[i].second, 0, sizeof(opose));
    memcpy(nvs[i].second.m, vposes + (size_t)i * ps, sizeof(double) * ps);
  }
  return o->remove_outliers(fs, nvs) ? 1 : 0;
}
/* Pcm::areLoopsConsistent (Pcm.h:670-718) for explicit pairs of a closure table, against the trajectories this oracle
 * has folded so far: pair t = (closure pi[t] as a_lc_b — the OLDER one —, closure pj[t] as c_lc_d).  Lets the parity tests
 * check sampled pairs of matrices far too large to build on one core.  ok_out[t] = decision, dist_out[t] = the distance
 * it was taken on, band_out[t] = 1 when |dist - threshold| < band (the near-threshold flag, PCM mode). */
void orc_check_pairs(void* h, int n, const uint64_t* k1, const uint64_t* k2, const double* poses, const double* covs,
                     long long m, const int* pi, const int* pj, unsigned char* ok_out, double* dist_out, unsigned char* band_out) {
  Oracle* o = (Oracle*)h;
  const int ps = o->d == 3 ? 12 : 4, nn = o_n(o->d) * o_n(o->d);
  auto make = [&](int i) {
    Factor f;
    memset(&f, 0, sizeof(Factor));
    f.type = 0;
    f.k1 = k1[i];
    f.k2 = k2[i];
    memcpy(f.pose, poses + (size_t)i * ps, sizeof(double) * ps);
    memcpy(f.cov, covs + (size_t)i * nn, sizeof(double) * nn);
    f.id = i;
    return f;
  };
  (void)n;
  for (long long t = 0; t < m; ++t) {
    const Factor a = make(pi[t]), c = make(pj[t]);
    double dist = 0.0;
    const bool ok = o->are_loops_consistent(a, c, &dist);
    if (ok_out) ok_out[t] = ok ? 1 : 0;
    if (dist_out) dist_out[t] = dist;
    if (band_out) band_out[t] = (o->mode == MODE_PCM && fabs(dist - o->lc_threshold) < o->band) ? 1 : 0;
  }
}
long long orc_num_lc(void* h) { return (long long)((Oracle*)h)->total_lc; }
long long orc_num_inliers(void* h) { return (long long)((Oracle*)h)->total_good_lc; }
long long orc_num_odom(void* h) { return (long long)((Oracle*)h)->nfg_odom.size(); }
long long orc_num_special(void* h) { return (long long)((Oracle*)h)->nfg_special.size(); }
long long orc_num_values(void* h) { return (long long)((Oracle*)h)->values.size(); }
long long orc_pair_checks(void* h) { return (long long)((Oracle*)h)->pair_checks; }
long long orc_output_size(void* h) { return (long long)((Oracle*)h)->output.size(); }
void orc_output_ids(void* h, long long* ids) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->output.size(); ++i) ids[i] = o->output[i];
}
int orc_num_groups(void* h) { return (int)((Oracle*)h)->group_order.size(); }
/* group g in first-seen order: prefixes and sizes */
void orc_group_info(void* h, int g, int* c1, int* c2, int* n, int* n_inliers) {
  Oracle* o = (Oracle*)h;
  const ObsId id = o->group_order[g];
  const Measurements& m = o->loop_closures[id];
  *c1 = id.a;
  *c2 = id.b;
  *n = (int)m.factors.size();
  *n_inliers = (int)m.consistent.size();
}
/* dense adjacency (uint8) and distances (double), n*n each, n = adjacency dimension */
int orc_group_adj(void* h, int g, unsigned char* adj, double* dist) {
  Oracle* o = (Oracle*)h;
  const Measurements& m = o->loop_closures[o->group_order[g]];
  std::vector<double> a, dd;
  o->dense(m, &a, &dd);
  const size_t n = (size_t)m.n_adj;
  for (size_t i = 0; i < n * n; ++i) {
    if (adj) adj[i] = a[i] != 0.0;
    if (dist) dist[i] = dd[i];
  }
  return (int)n;
}
void orc_group_factor_ids(void* h, int g, long long* ids) {
  Oracle* o = (Oracle*)h;
  const Measurements& m = o->loop_closures[o->group_order[g]];
  for (size_t i = 0; i < m.factors.size(); ++i) ids[i] = m.factors[i].id;
}
void orc_group_inlier_ids(void* h, int g, long long* ids) {
  Oracle* o = (Oracle*)h;
  const Measurements& m = o->loop_closures[o->group_order[g]];
  for (size_t i = 0; i < m.consistent.size(); ++i) ids[i] = m.consistent[i];
}
long long orc_num_flagged(void* h) { return (long long)((Oracle*)h)->flagged.size(); }
void orc_flagged(void* h, long long* pairs) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->flagged.size(); ++i) {
    pairs[2 * i] = o->flagged[i].first;
    pairs[2 * i + 1] = o->flagged[i].second;
  }
}
int orc_remove_last(void* h, int c1, int c2, uint64_t* k1, uint64_t* k2) {
  return ((Oracle*)h)->remove_last(make_obs((unsigned char)c1, (unsigned char)c2), k1, k2);
}
int orc_remove_last_any(void* h, uint64_t* k1, uint64_t* k2) { /* Pcm.h:346-353 */
  Oracle* o = (Oracle*)h;
  if (o->lc_in_order.empty()) return 0;
  ObsId last = o->lc_in_order.back();
  o->lc_in_order.pop_back();
  return o->remove_last(last, k1, k2);
}
void orc_ignore_prefix(void* h, int c) { /* Pcm.h:357-365 */
  Oracle* o = (Oracle*)h;
  if (std::find(o->ignored.begin(), o->ignored.end(), (unsigned char)c) == o->ignored.end())
    o->ignored.push_back((unsigned char)c);
  o->build_graph();
}
void orc_revive_prefix(void* h, int c) { /* Pcm.h:369-377 */
  Oracle* o = (Oracle*)h;
  o->ignored.erase(std::remove(o->ignored.begin(), o->ignored.end(), (unsigned char)c), o->ignored.end());
  o->build_graph();
}
/* trajectory entry lookup for parity tests: returns 1 if present */
int orc_traj_get(void* h, uint64_t key, double* pose, double* cov, int* node, int* rot_info) {
  Oracle* o = (Oracle*)h;
  auto t = o->traj.find(key_chr(key));
  if (t == o->traj.end()) return 0;
  auto it = t->second.find(key);
  if (it == t->second.end()) return 0;
  const int ps = o->d == 3 ? 12 : 4, n = o_n(o->d);
  if (o->mode == MODE_PCM) {
    memcpy(pose, it->second.c.pose.m, sizeof(double) * ps);
    if (cov) memcpy(cov, it->second.c.cov, sizeof(double) * n * n);
    if (node) *node = 0;
    if (rot_info) *rot_info = it->second.c.rotation_info;
  } else {
    memcpy(pose, it->second.s.pose.m, sizeof(double) * ps);
    if (node) *node = it->second.s.node;
    if (rot_info) *rot_info = it->second.s.rotation_info;
  }
  return 1;
}

/* ---- primitive-level entry points (golden-vector tests) ------------------------------- */
void orc_pose_compose(int d, const double* a, const double* b, double* out) {
  opose pa, pb, r;
  memset(&pa, 0, sizeof(pa)); memset(&pb, 0, sizeof(pb));
  const int ps = d == 3 ? 12 : 4;
  memcpy(pa.m, a, sizeof(double) * ps); memcpy(pb.m, b, sizeof(double) * ps);
  o_pose_compose(d, &pa, &pb, &r);
  memcpy(out, r.m, sizeof(double) * ps);
}
void orc_pose_inverse(int d, const double* a, double* out) {
  opose pa, r;
  memset(&pa, 0, sizeof(pa));
  const int ps = d == 3 ? 12 : 4;
  memcpy(pa.m, a, sizeof(double) * ps);
  o_pose_inverse(d, &pa, &r);
  memcpy(out, r.m, sizeof(double) * ps);
}
void orc_logmap(int d, const double* a, double* v) {
  opose pa;
  memset(&pa, 0, sizeof(pa));
  memcpy(pa.m, a, sizeof(double) * (d == 3 ? 12 : 4));
  o_logmap(d, &pa, v);
}
static void load_pwc(int d, const double* pose, const double* cov, int rot, opwc* p) {
  memset(p, 0, sizeof(*p));
  memcpy(p->pose.m, pose, sizeof(double) * (d == 3 ? 12 : 4));
  memcpy(p->cov, cov, sizeof(double) * o_n(d) * o_n(d));
  p->rotation_info = rot;
}
static void store_pwc(int d, const opwc* p, double* pose, double* cov, int* rot) {
  memcpy(pose, p->pose.m, sizeof(double) * (d == 3 ? 12 : 4));
  memcpy(cov, p->cov, sizeof(double) * o_n(d) * o_n(d));
  if (rot) *rot = p->rotation_info;
}
void orc_pwc_compose(int d, const double* pa, const double* ca, int ra, const double* pb, const double* cb, int rb,
                     double* po, double* co, int* ro) {
  opwc a, b, r;
  load_pwc(d, pa, ca, ra, &a); load_pwc(d, pb, cb, rb, &b);
  o_pwc_compose(d, &a, &b, &r);
  store_pwc(d, &r, po, co, ro);
}
void orc_pwc_between(int d, const double* pa, const double* ca, int ra, const double* pb, const double* cb, int rb,
                     double* po, double* co, int* ro) {
  opwc a, b, r;
  load_pwc(d, pa, ca, ra, &a); load_pwc(d, pb, cb, rb, &b);
  o_pwc_between(d, &a, &b, &r);
  store_pwc(d, &r, po, co, ro);
}
void orc_pwc_inverse(int d, const double* pa, const double* ca, int ra, double* po, double* co, int* ro) {
  opwc a, r;
  load_pwc(d, pa, ca, ra, &a);
  o_pwc_inverse(d, &a, &r);
  store_pwc(d, &r, po, co, ro);
}
double orc_pwc_mahalanobis(int d, const double* pa, const double* ca, int ra) {
  opwc a;
  load_pwc(d, pa, ca, ra, &a);
  return o_pwc_mahalanobis(d, &a);
}
void orc_pwn_norms(int d, const double* pa, int node, int rot, double* trans, double* rotn) {
  opwn a;
  memset(&a, 0, sizeof(a));
  memcpy(a.pose.m, pa, sizeof(double) * (d == 3 ? 12 : 4));
  a.node = node;
  a.rotation_info = rot;
  *trans = o_pwn_avg_trans(d, &a);
  *rotn = o_pwn_avg_rot(d, &a);
}
int orc_llt_ok(int n, const double* M) { return o_llt_ok(n, M); }
void orc_lu_inverse(int n, const double* M, double* inv) { o_lu_inverse(n, M, inv); }

/* left fold of P-1 odometry deltas (Pcm.h:545-556).  cum_* has P entries; entry 0 = (init_pose, 0, node 0). */
void orc_traj_fold(int d, int mode, int P, const double* init_pose, const double* dpose, const double* dcov,
                   double* cum_pose, double* cum_cov, int* cum_node, int* cum_rot) {
  const int ps = d == 3 ? 12 : 4, n = o_n(d);
  OT cur = ot_default(d, mode);
  if (mode == MODE_PCM) memcpy(cur.c.pose.m, init_pose, sizeof(double) * ps);
  else memcpy(cur.s.pose.m, init_pose, sizeof(double) * ps);
  for (int k = 0; k < P; ++k) {
    if (k > 0) {
      OT delta = ot_from_factor(d, mode, dpose + (size_t)(k - 1) * ps, dcov + (size_t)(k - 1) * n * n);
      cur = ot_compose(d, mode, cur, delta);
    }
    if (mode == MODE_PCM) {
      memcpy(cum_pose + (size_t)k * ps, cur.c.pose.m, sizeof(double) * ps);
      if (cum_cov) memcpy(cum_cov + (size_t)k * n * n, cur.c.cov, sizeof(double) * n * n);
      if (cum_node) cum_node[k] = 0;
      if (cum_rot) cum_rot[k] = cur.c.rotation_info;
    } else {
      memcpy(cum_pose + (size_t)k * ps, cur.s.pose.m, sizeof(double) * ps);
      if (cum_node) cum_node[k] = cur.s.node;
      if (cum_rot) cum_rot[k] = cur.s.rotation_info;
    }
  }
}

/* ---- clique entry points on a dense uint8 adjacency (n*n) ------------------------------ */
static std::vector<double> to_double(int n, const unsigned char* adj) {
  std::vector<double> a((size_t)n * n);
  for (size_t i = 0; i < (size_t)n * n; ++i) a[i] = adj[i] ? 1.0 : 0.0;
  return a;
}
/* returns size; ids_out receives the first `size` entries of the reference's returned buffer */
int orc_clique_heu(int n, const unsigned char* adj, int* ids_out) {
  std::vector<int> out;
  auto a = to_double(n, adj);
  const int k = find_max_clique_heu(n, a.data(), &out);
  for (int i = 0; i < k && i < (int)out.size(); ++i) ids_out[i] = out[i];
  return k;
}
int orc_clique_heu_incremental(int n, const unsigned char* adj, int num_new, int prev, int* ids_out) {
  std::vector<int> out;
  auto a = to_double(n, adj);
  const int k = find_max_clique_heu_incremental(n, a.data(), (size_t)num_new, (size_t)prev, &out);
  for (int i = 0; i < k && i < (int)out.size(); ++i) ids_out[i] = out[i];
  return k;
}
int orc_clique_exact(int n, const unsigned char* adj, int* ids_out) {
  std::vector<int> out;
  auto a = to_double(n, adj);
  Csr g = csr_from_dense(n, a.data());
  const int k = exact_restated(g, 0, &out);
  for (int i = 0; i < k && i < (int)out.size(); ++i) ids_out[i] = out[i];
  return k;
}

}  /* extern "C" */
