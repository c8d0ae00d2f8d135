// rpgo_host.hpp — C++ host side above the C ABI, mirroring the reference's plug-in interface for the PCM
// path with the same names, argument meaning and return values:
//
//   KimeraRPGO::OutlierRemoval            reference include/KimeraRPGO/outlier/OutlierRemoval.h:19-102
//   KimeraRPGO::PcmParams / RobustSolverParams (setPcm3DParams, setPcm2DParams, setPcmSimple*Params,
//                                         setIncremental, specialSymbols)   include/KimeraRPGO/SolverParams.h:33-279
//   KimeraRPGO::RobustSolver::update / removeLastLoopClosure / ignorePrefix / revivePrefix /
//                                         removePriorFactorsWithPrefix       include/KimeraRPGO/RobustSolver.h:30-139
//   typedefs Pcm2D, Pcm3D, PcmSimple2D, PcmSimple3D                          include/KimeraRPGO/outlier/Pcm.h:1167-1170
//
// GTSAM is not available in this build image, so the handful of GTSAM value types the interface passes
// around (Key/Symbol, Pose2/Pose3, BetweenFactor/PriorFactor, NonlinearFactorGraph, Values, noise-model
// covariance) are represented by the minimal stand-ins in namespace gtsam_lite; INTEGRATION.md shows the same
// adapter written against real GTSAM.  All arithmetic goes through include/rpgo_b200.h to the GPU; the
// optimiser is NOT part of this path (RobustSolver::update here stops where the reference would call
// GTSAM's LM/GN/GNC, src/RobustSolver.cpp:351).
#pragma once

#include <cstdio>
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rpgo_b200.h"

namespace gtsam_lite {

using Key = std::uint64_t;
struct Symbol {
  Key key;
  Symbol(unsigned char c, std::uint64_t j) : key((Key(c) << 56) | j) {}
  explicit Symbol(Key k) : key(k) {}
  unsigned char chr() const { return (unsigned char)(key >> 56); }
  std::uint64_t index() const { return key & ((Key(1) << 56) - 1); }
  operator Key() const { return key; }
};

struct Pose3 {                      // R row-major + t, the ABI's layout
  std::array<double, 12> m{{1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}};
  static constexpr int dimension = 6, storage = 12;
  Pose3() = default;
  Pose3(const std::array<double, 9>& R, double x, double y, double z) {
    for (int i = 0; i < 9; ++i) m[i] = R[i];
    m[9] = x; m[10] = y; m[11] = z;
  }
  static Pose3 Translation(double x, double y, double z) { Pose3 p; p.m[9] = x; p.m[10] = y; p.m[11] = z; return p; }
  static std::array<double, 9> Rz(double t) { return {{std::cos(t), -std::sin(t), 0, std::sin(t), std::cos(t), 0, 0, 0, 1}}; }
};
struct Pose2 {                      // cos, sin, x, y
  std::array<double, 4> m{{1, 0, 0, 0}};
  static constexpr int dimension = 3, storage = 4;
  Pose2() = default;
  Pose2(double x, double y, double theta) : m{{std::cos(theta), std::sin(theta), x, y}} {}
};

struct Factor {
  virtual ~Factor() = default;
  virtual std::vector<Key> keys() const = 0;
  Key front() const { return keys().front(); }
  Key back() const { return keys().back(); }
};
template <class P>
struct BetweenFactor : Factor {
  Key k1, k2;
  P measured;
  std::vector<double> covariance;   // dimension x dimension, row-major (Gaussian::covariance())
  BetweenFactor(Key a, Key b, const P& z, const std::vector<double>& cov) : k1(a), k2(b), measured(z), covariance(cov) {}
  std::vector<Key> keys() const override { return {k1, k2}; }
};
template <class P>
struct PriorFactor : Factor {
  Key k;
  P prior;
  std::vector<double> covariance;
  PriorFactor(Key a, const P& z, const std::vector<double>& cov) : k(a), prior(z), covariance(cov) {}
  std::vector<Key> keys() const override { return {k}; }
};
inline std::vector<double> IsotropicVariance(int dim, double v) {   // noiseModel::Isotropic::Variance
  std::vector<double> c(dim * dim, 0.0);
  for (int i = 0; i < dim; ++i) c[i * dim + i] = v;
  return c;
}

class NonlinearFactorGraph {
 public:
  using sharedFactor = std::shared_ptr<Factor>;
  template <class F>
  void add(const F& f) { v_.push_back(std::make_shared<F>(f)); }
  void add(const sharedFactor& f) { v_.push_back(f); }
  void add(const NonlinearFactorGraph& g) { v_.insert(v_.end(), g.v_.begin(), g.v_.end()); }
  size_t size() const { return v_.size(); }
  const sharedFactor& operator[](size_t i) const { return v_[i]; }
  void pop_back() { v_.pop_back(); }
  std::vector<sharedFactor>::const_iterator begin() const { return v_.begin(); }
  std::vector<sharedFactor>::const_iterator end() const { return v_.end(); }
 private:
  std::vector<sharedFactor> v_;
};

template <class P>
class ValuesT {
 public:
  void insert(Key k, const P& p) {
    if (!m_.emplace(k, p).second) throw std::runtime_error("Values: key already exists");  // gtsam::ValuesKeyAlreadyExists
  }
  void insert(const ValuesT& o) { for (auto& kv : o.m_) insert(kv.first, kv.second); }
  bool exists(Key k) const { return m_.count(k) != 0; }
  const P& at(Key k) const { return m_.at(k); }
  size_t size() const { return m_.size(); }
 private:
  std::map<Key, P> m_;
};

}  // namespace gtsam_lite

namespace KimeraRPGO {

using gtsam_lite::Key;

enum class Solver { LM, GN };
enum class OutlierRemovalMethod { NONE, PCM2D, PCM3D, PCM_Simple2D, PCM_Simple3D };
enum class Verbosity { UPDATE, QUIET, VERBOSE };

struct PcmParams {   // defaults: SolverParams.h:35-42; a threshold < 0 disables that check (Pcm.h:74-82)
  double odom_threshold = 10.0, lc_threshold = 5.0;
  double odom_trans_threshold = 0.05, odom_rot_threshold = 0.005;
  double dist_trans_threshold = 0.01, dist_rot_threshold = 0.001;
  bool incremental = false;
};

struct RobustSolverParams {
  Solver solver = Solver::LM;
  OutlierRemovalMethod outlierRemovalMethod = OutlierRemovalMethod::PCM3D;
  std::vector<char> specialSymbols;
  Verbosity verbosity = Verbosity::UPDATE;
  PcmParams pcm_params;

  void setNoRejection(Verbosity v = Verbosity::UPDATE) { outlierRemovalMethod = OutlierRemovalMethod::NONE; verbosity = v; }
  void setIncremental() { pcm_params.incremental = true; }
  void setPcm2DParams(double odomThreshold, double lcThreshold, Verbosity v = Verbosity::UPDATE) {
    setMahalanobis(OutlierRemovalMethod::PCM2D, odomThreshold, lcThreshold, v);
  }
  void setPcm3DParams(double odomThreshold, double lcThreshold, Verbosity v = Verbosity::UPDATE) {
    setMahalanobis(OutlierRemovalMethod::PCM3D, odomThreshold, lcThreshold, v);
  }
  void setPcmSimple2DParams(double trans, double rot, Verbosity v = Verbosity::UPDATE) {
    setSimple(OutlierRemovalMethod::PCM_Simple2D, trans, rot, trans, rot, v);
  }
  void setPcmSimple3DParams(double trans, double rot, Verbosity v = Verbosity::UPDATE) {
    setSimple(OutlierRemovalMethod::PCM_Simple3D, trans, rot, trans, rot, v);
  }
  void setPcmSimple2DParams(double tO, double rO, double tP, double rP, Verbosity v = Verbosity::UPDATE) {
    setSimple(OutlierRemovalMethod::PCM_Simple2D, tO, rO, tP, rP, v);
  }
  void setPcmSimple3DParams(double tO, double rO, double tP, double rP, Verbosity v = Verbosity::UPDATE) {
    setSimple(OutlierRemovalMethod::PCM_Simple3D, tO, rO, tP, rP, v);
  }

 private:
  void setMahalanobis(OutlierRemovalMethod m, double o, double l, Verbosity v) {
    outlierRemovalMethod = m; pcm_params.odom_threshold = o; pcm_params.lc_threshold = l; verbosity = v;
  }
  void setSimple(OutlierRemovalMethod m, double tO, double rO, double tP, double rP, Verbosity v) {
    outlierRemovalMethod = m;
    pcm_params.odom_trans_threshold = tO; pcm_params.odom_rot_threshold = rO;
    pcm_params.dist_trans_threshold = tP; pcm_params.dist_rot_threshold = rP;
    verbosity = v;
  }
};

struct ObservationId {   // unordered prefix pair, TypeUtils.h:44-58
  char id1, id2;
  ObservationId(char a, char b) : id1(a), id2(b) {}
  bool operator==(const ObservationId& o) const { return (id1 == o.id1 && id2 == o.id2) || (id1 == o.id2 && id2 == o.id1); }
};
struct Edge {
  gtsam_lite::Symbol from_key, to_key;
  Edge(Key a, Key b) : from_key(a), to_key(b) {}
};
using EdgePtr = std::unique_ptr<const Edge>;

template <class P>
class OutlierRemovalT {   // OutlierRemoval.h:19-102
 public:
  using Graph = gtsam_lite::NonlinearFactorGraph;
  using Values = gtsam_lite::ValuesT<P>;
  virtual ~OutlierRemovalT() = default;
  virtual size_t getNumLC() = 0;
  virtual size_t getNumLCInliers() = 0;
  virtual size_t getNumOdomFactors() = 0;
  virtual size_t getNumSpecialFactors() = 0;
  virtual bool removeOutliers(const Graph& new_factors, const Values& new_values, Graph* nfg, Values* values) = 0;
  virtual EdgePtr removeLastLoopClosure(ObservationId, Graph*) { return nullptr; }
  virtual EdgePtr removeLastLoopClosure(Graph*) { return nullptr; }
  virtual void ignoreLoopClosureWithPrefix(char, Graph*) {}
  virtual void reviveLoopClosureWithPrefix(char, Graph*) {}
  virtual std::vector<char> getIgnoredPrefixes() { return {}; }
  virtual void removePriorFactorsWithPrefix(const char&, Graph*) {}
  void setQuiet() { debug_ = false; }
 protected:
  bool debug_ = true;
};

// Pcm<poseT, T> on the GPU.  MODE: RPGO_MODE_PCM (PoseWithCovariance) or RPGO_MODE_SIMPLE (PoseWithNode).
template <class P, int MODE>
class PcmGpu : public OutlierRemovalT<P> {
 public:
  using Graph = gtsam_lite::NonlinearFactorGraph;
  using Values = gtsam_lite::ValuesT<P>;
  using Between = gtsam_lite::BetweenFactor<P>;
  using Prior = gtsam_lite::PriorFactor<P>;

  explicit PcmGpu(PcmParams p, const std::vector<char>& special_symbols = {}) : params_(p), special_symbols_(special_symbols) {
    rpgo_cfg c;
    rpgo_default_cfg(&c);
    c.dim = P::dimension == 6 ? 3 : 2;
    c.mode = MODE;
    c.odom_threshold = p.odom_threshold;             c.lc_threshold = p.lc_threshold;
    c.odom_trans_threshold = p.odom_trans_threshold; c.odom_rot_threshold = p.odom_rot_threshold;
    c.dist_trans_threshold = p.dist_trans_threshold; c.dist_rot_threshold = p.dist_rot_threshold;
    c.incremental = p.incremental ? 1 : 0;
    const int rc = rpgo_create(&c, &h_);
    if (rc != RPGO_OK) throw std::runtime_error("rpgo_create failed (" + std::to_string(rc) + "): no usable CUDA device, no CPU fallback");
    loop_check_ = !(p.lc_threshold < 0 || p.dist_rot_threshold < 0 || p.dist_trans_threshold < 0);
  }
  ~PcmGpu() override { rpgo_destroy(h_); }
  PcmGpu(const PcmGpu&) = delete;
  PcmGpu& operator=(const PcmGpu&) = delete;

  size_t getNumLC() override { return total_lc_; }
  size_t getNumLCInliers() override { return total_good_lc_; }
  size_t getNumOdomFactors() override { return nfg_odom_.size(); }
  size_t getNumSpecialFactors() override { return nfg_special_.size(); }

  // Pcm.h:148-281
  bool removeOutliers(const Graph& new_factors, const Values& new_values, Graph* output_nfg, Values* output_values) override {
    output_values->insert(new_values);
    if (new_factors.size() == 0) return false;
    bool do_optimize = false;
    std::vector<std::shared_ptr<Between>> odom, lcs;
    for (const auto& f : new_factors) {
      if (!f) continue;
      auto b = std::dynamic_pointer_cast<Between>(f);
      if (!b) {                                       // NONBETWEEN_FACTORS, Pcm.h:229-232
        nfg_special_.add(f);
        do_optimize = true;
        continue;
      }
      if (isSpecialSymbol(gtsam_lite::Symbol(b->k1).chr()) || isSpecialSymbol(gtsam_lite::Symbol(b->k2).chr())) {
        // landmark observation, Pcm.h:180-188
        if (new_values.exists(b->k1) || new_values.exists(b->k2)) landmarkFirst(b);   // FIRST_LANDMARK_OBSERVATION :207-220
        else if (b->k1 != b->k2) lcs.push_back(b);                                    // re-observation
        continue;
      }
      if (b->k1 + 1 == b->k2 && new_values.exists(b->k2)) odom.push_back(b);   // ODOMETRY, Pcm.h:189-191
      else if (b->k1 != b->k2) lcs.push_back(b);                                // LOOP_CLOSURE, Pcm.h:221-228
    }
    if (!odom.empty()) appendOdom(odom, *output_values);
    if (!lcs.empty()) {
      std::vector<std::shared_ptr<Between>> plain;
      for (auto& f : lcs) {
        if (isSpecialSymbol(gtsam_lite::Symbol(f->k1).chr()) || isSpecialSymbol(gtsam_lite::Symbol(f->k2).chr()))
          landmarkReobserve(f, *output_values);                                       // Pcm.h:437-455
        else
          plain.push_back(f);
      }
      lcs.swap(plain);
      std::map<int, size_t> num_new = appendLoopClosures(lcs, *output_values);
      if (params_.incremental) findInliersIncremental(num_new); else findInliers();
      do_optimize = true;
    }
    *output_nfg = buildGraphToOptimize();
    return do_optimize;
  }

  EdgePtr removeLastLoopClosure(ObservationId id, Graph* updated) override {   // Pcm.h:299-340
    const int g = rpgo_find_group(h_, (uint8_t)id.id1, (uint8_t)id.id2);
    if (g < 0 || g >= (int)groups_.size() || groups_[g].factors.size() == 0) return nullptr;
    return removeLastOf(g, updated);
  }
  EdgePtr removeLastLoopClosure(Graph* updated) override {                       // Pcm.h:346-353
    if (lc_in_order_.empty()) return nullptr;
    const int g = lc_in_order_.back();
    lc_in_order_.pop_back();
    if (groups_[g].factors.size() == 0) return nullptr;
    return removeLastOf(g, updated);
  }
  void ignoreLoopClosureWithPrefix(char prefix, Graph* updated) override {       // Pcm.h:357-365
    if (std::find(ignored_.begin(), ignored_.end(), prefix) == ignored_.end()) ignored_.push_back(prefix);
    *updated = buildGraphToOptimize();
  }
  void reviveLoopClosureWithPrefix(char prefix, Graph* updated) override {       // Pcm.h:369-377
    ignored_.erase(std::remove(ignored_.begin(), ignored_.end(), prefix), ignored_.end());
    *updated = buildGraphToOptimize();
  }
  std::vector<char> getIgnoredPrefixes() override { return ignored_; }
  void removePriorFactorsWithPrefix(const char& prefix, Graph* updated) override {   // Pcm.h:387-408
    Graph kept;
    for (const auto& f : nfg_special_) {
      auto p = std::dynamic_pointer_cast<Prior>(f);
      if (!p || gtsam_lite::Symbol(p->k).chr() != (unsigned char)prefix) kept.add(f);
    }
    nfg_special_ = kept;
    *updated = buildGraphToOptimize();
  }

  // inspection for tests
  rpgo_handle* handle() { return h_; }
  std::vector<int> inlierIndices(int g) const { return groups_[g].inlier_idx; }

 private:
  struct Group {
    Graph factors, consistent_factors;
    std::vector<int> inlier_idx;
    char id1 = 0, id2 = 0;
    bool is_landmark = false;
  };

  bool isSpecialSymbol(unsigned char c) const {
    return std::find(special_symbols_.begin(), special_symbols_.end(), (char)c) != special_symbols_.end();
  }
  void check(int rc, const char* what) {
    if (rc != RPGO_OK) throw std::runtime_error(std::string(what) + ": " + rpgo_last_error(h_));
  }

  void appendOdom(const std::vector<std::shared_ptr<Between>>& fs, const Values& vals) {   // Pcm.h:516-557
    const size_t n = fs.size(), ps = P::storage, nn = P::dimension * P::dimension;
    std::vector<uint64_t> prev(n), next(n);
    std::vector<double> pose(n * ps), cov(n * nn), init(n * ps);
    for (size_t i = 0; i < n; ++i) {
      prev[i] = fs[i]->k1;
      next[i] = fs[i]->k2;
      std::copy(fs[i]->measured.m.begin(), fs[i]->measured.m.end(), pose.begin() + i * ps);
      std::copy(fs[i]->covariance.begin(), fs[i]->covariance.end(), cov.begin() + i * nn);
      const P p0 = vals.exists(prev[i]) ? vals.at(prev[i]) : P();
      std::copy(p0.m.begin(), p0.m.end(), init.begin() + i * ps);
      nfg_odom_.add(std::static_pointer_cast<gtsam_lite::Factor>(fs[i]));
    }
    check(rpgo_odom_append(h_, (int64_t)n, prev.data(), next.data(), pose.data(), cov.data(), init.data()), "rpgo_odom_append");
  }

  std::map<int, size_t> appendLoopClosures(const std::vector<std::shared_ptr<Between>>& in, const Values& vals) {
    std::vector<std::shared_ptr<Between>> fs;
    for (auto& f : in)
      if (vals.exists(f->k1) && vals.exists(f->k2)) fs.push_back(f);                // Pcm.h:431-435
    std::map<int, size_t> num_new;
    const size_t n = fs.size(), ps = P::storage, nn = P::dimension * P::dimension;
    if (n == 0) return num_new;
    std::vector<uint64_t> kf(n), kt(n);
    std::vector<double> pose(n * ps), cov(n * nn);
    std::vector<uint8_t> acc(n);
    std::vector<int32_t> grp(n), idx(n);
    for (size_t i = 0; i < n; ++i) {
      kf[i] = fs[i]->k1;
      kt[i] = fs[i]->k2;
      std::copy(fs[i]->measured.m.begin(), fs[i]->measured.m.end(), pose.begin() + i * ps);
      std::copy(fs[i]->covariance.begin(), fs[i]->covariance.end(), cov.begin() + i * nn);
    }
    check(rpgo_lc_append(h_, (int64_t)n, kf.data(), kt.data(), pose.data(), cov.data(), acc.data(), grp.data(), idx.data(), nullptr),
          "rpgo_lc_append");
    for (size_t i = 0; i < n; ++i) {
      if (!acc[i]) continue;                                                         // dropped: inconsistent with odometry
      if (grp[i] >= (int)groups_.size()) groups_.resize(grp[i] + 1);
      Group& g = groups_[grp[i]];
      uint8_t a, b; int64_t cnt;
      rpgo_group_info(h_, grp[i], &a, &b, &cnt);
      g.id1 = (char)a; g.id2 = (char)b;
      g.factors.add(std::static_pointer_cast<gtsam_lite::Factor>(fs[i]));            // Pcm.h:481
      lc_in_order_.push_back(grp[i]);                                                // Pcm.h:482
      ++total_lc_;
      ++num_new[grp[i]];
    }
    return num_new;
  }

  // ---- landmarks: one group per landmark key (Pcm.h:109) ---------------------------------------------
  Key landmarkKey(const Between& f) const { return isSpecialSymbol(gtsam_lite::Symbol(f.k1).chr()) ? f.k1 : f.k2; }
  int landmarkAppend(const std::shared_ptr<Between>& f, bool reset) {
    const Key lkey = landmarkKey(*f);
    if (f->k1 == lkey) throw std::runtime_error("landmark observations must be stated pose -> landmark (Pcm.h:803-808)");
    int32_t g = -1;
    const uint64_t pk = f->k1;
    check(rpgo_landmark_append(h_, lkey, 1, &pk, f->measured.m.data(), f->covariance.data(), reset ? 1 : 0, &g), "rpgo_landmark_append");
    if (g >= (int)groups_.size()) groups_.resize(g + 1);
    if (landmark_group_.find(lkey) == landmark_group_.end()) { landmark_group_[lkey] = g; landmark_order_.push_back(lkey); }
    groups_[g].is_landmark = true;
    return g;
  }
  void landmarkFirst(const std::shared_ptr<Between>& f) {
    if (f->k1 == landmarkKey(*f)) {
      // stated landmark -> pose: the reference stores it and trips over it at the first re-observation (Pcm.h:803-808);
      // here it is skipped with a warning before anything of this update has been applied
      std::fprintf(stderr, "rpgo: landmark observation stated landmark -> pose skipped (must be pose -> landmark)\n");
      return;
    }
    const int g = landmarkAppend(f, true);
    groups_[g].factors = Graph();
    groups_[g].factors.add(std::static_pointer_cast<gtsam_lite::Factor>(f));
    groups_[g].consistent_factors = groups_[g].factors;
    ++total_lc_;
  }
  void landmarkReobserve(const std::shared_ptr<Between>& f, const Values& vals) {
    if (!vals.exists(f->k1) || !vals.exists(f->k2)) return;                           // Pcm.h:431-435
    if (f->k1 == landmarkKey(*f)) return;
    const int g = landmarkAppend(f, false);
    groups_[g].factors.add(std::static_pointer_cast<gtsam_lite::Factor>(f));
    ++total_lc_;
  }

  void selectInliers(int g, int mode, int64_t n_new, int64_t prev, bool keep_if_zero) {
    Group& m = groups_[g];
    std::vector<int32_t> ids(std::max<size_t>(m.factors.size(), 1));
    int64_t k = 0;
    check(rpgo_find_inliers(h_, g, mode, n_new, prev, ids.data(), &k, nullptr), "rpgo_find_inliers");
    if (k == 0 && keep_if_zero) return;                                              // Pcm.h:936-939
    m.consistent_factors = Graph();
    m.inlier_idx.assign(ids.begin(), ids.begin() + k);
    for (int64_t i = 0; i < k; ++i) m.consistent_factors.add(m.factors[ids[i]]);     // Pcm.h:867-869
  }
  void findInliers() {                                                               // Pcm.h:851-876
    total_good_lc_ = 0;
    // the groups are independent: one batched call searches them concurrently (rpgo_find_inliers_batch)
    std::vector<int32_t> todo;
    std::vector<int64_t> offset;
    int64_t total = 0;
    for (size_t g = 0; g < groups_.size(); ++g) {
      Group& m = groups_[g];
      if (m.factors.size() == 0) { m.consistent_factors = Graph(); continue; }
      if (loop_check_ || m.is_landmark) {                                             // landmarks: Pcm.h:878-895
        todo.push_back((int32_t)g);
        offset.push_back(total);
        total += (int64_t)m.factors.size();
      } else {
        m.consistent_factors = m.factors;
      }
    }
    if (!todo.empty()) {
      std::vector<int32_t> ids((size_t)std::max<int64_t>(total, 1));
      std::vector<int64_t> sizes(todo.size(), 0);
      check(rpgo_find_inliers_batch(h_, (int32_t)todo.size(), todo.data(), RPGO_CLIQUE_HEU, nullptr, nullptr, ids.data(),
                                    offset.data(), sizes.data()), "rpgo_find_inliers_batch");
      for (size_t q = 0; q < todo.size(); ++q) {
        Group& m = groups_[todo[q]];
        m.consistent_factors = Graph();
        m.inlier_idx.assign(ids.begin() + offset[q], ids.begin() + offset[q] + sizes[q]);
        for (int64_t i = 0; i < sizes[q]; ++i) m.consistent_factors.add(m.factors[ids[offset[q] + i]]);   // Pcm.h:867-869
      }
    }
    for (auto& m : groups_) total_good_lc_ += m.consistent_factors.size();
  }
  void findInliersIncremental(const std::map<int, size_t>& num_new) {                // Pcm.h:906-947
    for (auto& kv : num_new) {
      Group& m = groups_[kv.first];
      if (!loop_check_ && !m.is_landmark) {
        // the reference runs findMaxCliqueHeuIncremental on the 1x1 zero matrix the disabled check leaves behind
        // (Pcm.h:484-486): vertex 0 is selected only when exactly one closure is new and nothing was selected before
        if (kv.second == 1 && m.consistent_factors.size() == 0) {
          m.consistent_factors = Graph();
          m.consistent_factors.add(m.factors[0]);
          m.inlier_idx.assign(1, 0);
        }
        continue;
      }
      selectInliers(kv.first, RPGO_CLIQUE_HEU_INCREMENTAL, (int64_t)kv.second, (int64_t)m.consistent_factors.size(), true);
    }
    for (size_t g = 0; g < groups_.size(); ++g)
      if (groups_[g].is_landmark && groups_[g].factors.size() > 0) selectInliers((int)g, RPGO_CLIQUE_HEU, 0, 0, false);   // Pcm.h:950-966
    total_good_lc_ = 0;
    for (auto& m : groups_) total_good_lc_ += m.consistent_factors.size();
  }
  EdgePtr removeLastOf(int g, Graph* updated) {
    Group& m = groups_[g];
    uint64_t k1 = 0, k2 = 0;
    check(rpgo_lc_remove_last(h_, g, &k1, &k2), "rpgo_lc_remove_last");
    m.factors.pop_back();
    if (m.factors.size() < 2) m.consistent_factors = m.factors;                      // Pcm.h:316-317
    else if (!loop_check_ && !m.is_landmark) {
      // no adjacency exists when the pairwise check is disabled; the reference's path (0x0 block into findMaxCliqueHeu) is
      // undefined behaviour.  Defined as findInliers' rule for that configuration (Pcm.h:870-873): every factor an inlier;
      // incremental mode keeps the previous selection, clipped to the remaining factors
      if (params_.incremental) {
        Graph kept;
        for (size_t i = 0; i < m.inlier_idx.size(); ++i)
          if ((size_t)m.inlier_idx[i] < m.factors.size()) kept.add(m.factors[m.inlier_idx[i]]);
        m.consistent_factors = kept;
        m.inlier_idx.erase(std::remove_if(m.inlier_idx.begin(), m.inlier_idx.end(), [&](int32_t i) { return (size_t)i >= m.factors.size(); }),
                           m.inlier_idx.end());
      } else {
        m.consistent_factors = m.factors;
      }
    } else selectInliers(g, RPGO_CLIQUE_HEU, 0, 0, false);
    *updated = buildGraphToOptimize();
    return EdgePtr(new Edge(k1, k2));
  }
  Graph buildGraphToOptimize() {                                                     // Pcm.h:977-1005
    Graph out;
    out.add(nfg_odom_);
    out.add(nfg_special_);
    for (auto& m : groups_) {
      if (m.is_landmark) continue;
      if (std::find(ignored_.begin(), ignored_.end(), m.id1) != ignored_.end()) continue;
      if (std::find(ignored_.begin(), ignored_.end(), m.id2) != ignored_.end()) continue;
      out.add(m.consistent_factors);
    }
    for (Key lk : landmark_order_) out.add(groups_[landmark_group_.at(lk)].consistent_factors);   // Pcm.h:996-1002
    return out;
  }

  rpgo_handle* h_ = nullptr;
  PcmParams params_;
  std::vector<char> special_symbols_;
  bool loop_check_ = true;
  Graph nfg_odom_, nfg_special_;
  std::vector<Group> groups_;
  std::map<Key, int> landmark_group_;
  std::vector<Key> landmark_order_;
  std::vector<int> lc_in_order_;
  std::vector<char> ignored_;
  size_t total_lc_ = 0, total_good_lc_ = 0;
};

using Pcm2D = PcmGpu<gtsam_lite::Pose2, RPGO_MODE_PCM>;
using Pcm3D = PcmGpu<gtsam_lite::Pose3, RPGO_MODE_PCM>;
using PcmSimple2D = PcmGpu<gtsam_lite::Pose2, RPGO_MODE_SIMPLE>;
using PcmSimple3D = PcmGpu<gtsam_lite::Pose3, RPGO_MODE_SIMPLE>;

// RobustSolver facade (RobustSolver.h:30-139, src/RobustSolver.cpp:35-72, :337-446) without the optimiser.
template <class P>
class RobustSolverT {
 public:
  using Graph = gtsam_lite::NonlinearFactorGraph;
  using Values = gtsam_lite::ValuesT<P>;
  explicit RobustSolverT(const RobustSolverParams& params) {
    const bool is3d = P::dimension == 6;
    switch (params.outlierRemovalMethod) {
      case OutlierRemovalMethod::PCM2D:
      case OutlierRemovalMethod::PCM3D:
        outlier_removal_.reset(new PcmGpu<P, RPGO_MODE_PCM>(params.pcm_params, params.specialSymbols));
        break;
      case OutlierRemovalMethod::PCM_Simple2D:
      case OutlierRemovalMethod::PCM_Simple3D:
        outlier_removal_.reset(new PcmGpu<P, RPGO_MODE_SIMPLE>(params.pcm_params, params.specialSymbols));
        break;
      case OutlierRemovalMethod::NONE:
        break;
    }
    (void)is3d;
  }
  // update(): "loadGraph"/"addGraph" of the README are this same call (tests/testLoadGraph.cpp)
  void update(const Graph& factors = Graph(), const Values& values = Values(), bool optimize_graph = true) {
    (void)optimize_graph;   // the LM/GN/GNC solve stays on GTSAM and is outside this path
    if (outlier_removal_) outlier_removal_->removeOutliers(factors, values, &nfg_, &values_);
    else { nfg_.add(factors); values_.insert(values); }
  }
  EdgePtr removeLastLoopClosure(char p1, char p2) {
    return outlier_removal_ ? outlier_removal_->removeLastLoopClosure(ObservationId(p1, p2), &nfg_) : nullptr;
  }
  EdgePtr removeLastLoopClosure() { return outlier_removal_ ? outlier_removal_->removeLastLoopClosure(&nfg_) : nullptr; }
  void ignorePrefix(char p) { if (outlier_removal_) outlier_removal_->ignoreLoopClosureWithPrefix(p, &nfg_); }
  void revivePrefix(char p) { if (outlier_removal_) outlier_removal_->reviveLoopClosureWithPrefix(p, &nfg_); }
  std::vector<char> getIgnoredPrefixes() { return outlier_removal_ ? outlier_removal_->getIgnoredPrefixes() : std::vector<char>(); }
  void removePriorFactorsWithPrefix(const char& p) { if (outlier_removal_) outlier_removal_->removePriorFactorsWithPrefix(p, &nfg_); }
  const Graph& getFactorsUnsafe() const { return nfg_; }
  const Values& calculateEstimate() const { return values_; }   // initial values: no solve on this path
  size_t getNumLC() { return outlier_removal_ ? outlier_removal_->getNumLC() : 0; }
  size_t getNumLCInliers() { return outlier_removal_ ? outlier_removal_->getNumLCInliers() : 0; }
 private:
  std::unique_ptr<OutlierRemovalT<P>> outlier_removal_;
  Graph nfg_;
  Values values_;
};
using RobustSolver = RobustSolverT<gtsam_lite::Pose3>;
using RobustSolver2D = RobustSolverT<gtsam_lite::Pose2>;

}  // namespace KimeraRPGO
