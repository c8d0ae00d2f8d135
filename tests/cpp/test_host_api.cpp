// C++ tests of the reference-facing host API (kimera-rpgo_b200/host/rpgo_host.hpp) on the GPU.
// Each TEST transcribes a reference test (file:line cited) with the same names and expectations.
//   g++ -std=c++17 -Iinclude -Ikimera-rpgo_b200/host tests/cpp/test_host_api.cpp -Lkimera-rpgo_b200 -lrpgo_b200
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "rpgo_host.hpp"

using namespace KimeraRPGO;
using gtsam_lite::BetweenFactor;
using gtsam_lite::IsotropicVariance;
using gtsam_lite::NonlinearFactorGraph;
using gtsam_lite::Pose3;
using gtsam_lite::PriorFactor;
using gtsam_lite::Symbol;
using Values = gtsam_lite::ValuesT<Pose3>;

static int failures = 0;
#define EXPECT(c)                                                          \
  do {                                                                     \
    if (!(c)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #c); ++failures; } \
  } while (0)

static const std::array<double, 9> R90{{0, -1, 0, 1, 0, 0, 0, 0, 1}};
static const std::array<double, 9> RM90{{0, 1, 0, -1, 0, 0, 0, 0, 1}};
static const std::array<double, 9> R180{{-1, 0, 0, 0, -1, 0, 0, 0, 1}};
static const std::array<double, 9> I3{{1, 0, 0, 0, 1, 0, 0, 0, 1}};

// tests/testPcm.cpp:17-86
static void Pcm_OdometryCheck() {
  PcmParams params;
  params.lc_threshold = -1;
  params.odom_threshold = 0.3;
  std::unique_ptr<OutlierRemovalT<Pose3>> pcm(new Pcm3D(params));
  pcm->setQuiet();
  NonlinearFactorGraph nfg;
  Values est;
  Values init_vals;
  NonlinearFactorGraph init_factors;
  init_vals.insert(0, Pose3());
  init_factors.add(PriorFactor<Pose3>(0, Pose3(), IsotropicVariance(6, 0.01)));
  pcm->removeOutliers(init_factors, init_vals, &nfg, &est);
  for (size_t i = 0; i < 3; i++) {
    Values odom_val;
    NonlinearFactorGraph odom_factor;
    Pose3 odom(R90, 1, 0, 0);
    odom_val.insert(i + 1, odom);
    odom_factor.add(BetweenFactor<Pose3>(i, i + 1, odom, IsotropicVariance(6, 0.1)));
    pcm->removeOutliers(odom_factor, odom_val, &nfg, &est);
  }
  EXPECT(size_t(4) == nfg.size());
  EXPECT(size_t(4) == est.size());
  NonlinearFactorGraph lc_factor1;
  lc_factor1.add(BetweenFactor<Pose3>(3, 0, Pose3(Pose3::Rz(1.51), 0.8, 0, 0), IsotropicVariance(6, 0.1)));
  bool do_optimize = pcm->removeOutliers(lc_factor1, Values(), &nfg, &est);
  EXPECT(size_t(5) == nfg.size());
  EXPECT(size_t(4) == est.size());
  EXPECT(do_optimize == true);
  NonlinearFactorGraph lc_factor2;
  lc_factor2.add(BetweenFactor<Pose3>(3, 0, Pose3(Pose3::Rz(1.51), 0.8, 0, 0), IsotropicVariance(6, 0.05)));
  do_optimize = pcm->removeOutliers(lc_factor2, Values(), &nfg, &est);
  EXPECT(size_t(5) == nfg.size());
  EXPECT(size_t(4) == est.size());
  EXPECT(do_optimize == true);
}

// tests/testPcm.cpp:89-192
static void Pcm_ConsistencyCheck() {
  PcmParams params;
  params.lc_threshold = 0.5;
  params.odom_threshold = -1;
  std::unique_ptr<Pcm3D> pcm(new Pcm3D(params));
  NonlinearFactorGraph nfg;
  Values est;
  Values init_vals;
  NonlinearFactorGraph init_factors;
  init_vals.insert(0, Pose3());
  init_factors.add(PriorFactor<Pose3>(0, Pose3(), IsotropicVariance(6, 0.01)));
  pcm->removeOutliers(init_factors, init_vals, &nfg, &est);
  for (size_t i = 0; i < 6; i++) {
    Values odom_val;
    NonlinearFactorGraph odom_factor;
    Pose3 odom(i < 2 ? R90 : I3, 1, 0, 0);
    odom_val.insert(i + 1, odom);
    odom_factor.add(BetweenFactor<Pose3>(i, i + 1, odom, IsotropicVariance(6, 0.1)));
    pcm->removeOutliers(odom_factor, odom_val, &nfg, &est);
  }
  EXPECT(size_t(7) == nfg.size());
  EXPECT(size_t(7) == est.size());
  const auto noiseLc = IsotropicVariance(6, 0.1);
  NonlinearFactorGraph lc1, lc2, lc3, lc4;
  lc1.add(BetweenFactor<Pose3>(3, 0, Pose3(Pose3::Rz(3.1416), 0, 0.9, 0), noiseLc));
  pcm->removeOutliers(lc1, Values(), &nfg, &est);
  lc2.add(BetweenFactor<Pose3>(4, 0, Pose3(Pose3::Rz(3.1416), -1, 0.8, 0), noiseLc));
  bool do_optimize = pcm->removeOutliers(lc2, Values(), &nfg, &est);
  EXPECT(size_t(9) == nfg.size());
  EXPECT(size_t(7) == est.size());
  EXPECT(do_optimize == true);
  lc3.add(BetweenFactor<Pose3>(5, 0, Pose3(Pose3::Rz(0.99 * 3.1416), -1.8, 0.8, 0), noiseLc));
  do_optimize = pcm->removeOutliers(lc3, Values(), &nfg, &est);
  EXPECT(size_t(10) == nfg.size());
  EXPECT(do_optimize == true);
  lc4.add(BetweenFactor<Pose3>(6, 0, Pose3(Pose3::Rz(0.98 * 3.1416), -2.6, 0.6, 0), noiseLc));
  do_optimize = pcm->removeOutliers(lc4, Values(), &nfg, &est);
  EXPECT(size_t(10) == nfg.size());          // lc4 only consistent with lc3: not in the clique
  EXPECT(size_t(7) == est.size());
  EXPECT(pcm->getNumLC() == 4 && pcm->getNumLCInliers() == 3);
  EXPECT((pcm->inlierIndices(0) == std::vector<int>{0, 1, 2}));
}

// tests/testMultiRobot.cpp:22-188 (Pcm3D) and :191-357 (PcmSimple3D)
static void RobustSolver_multiRobot(bool simple) {
  RobustSolverParams params;
  if (simple) params.setPcmSimple3DParams(0.04, 0.01, Verbosity::QUIET);
  else params.setPcm3DParams(3.0, 0.05, Verbosity::QUIET);
  std::unique_ptr<RobustSolver> pgo(new RobustSolver(params));
  const auto noise = IsotropicVariance(6, 0.1);
  auto a = [](size_t i) { return Key(Symbol('a', i)); };
  auto b = [](size_t i) { return Key(Symbol('b', i)); };
  Values init_vals;
  init_vals.insert(a(0), Pose3());
  init_vals.insert(b(0), Pose3::Translation(0, -1, 0));
  pgo->update(NonlinearFactorGraph(), init_vals);
  for (size_t i = 0; i < 3; i++) {
    Values vals;
    NonlinearFactorGraph odom_factors;
    Pose3 odom = Pose3::Translation(1, 0, 0);
    vals.insert(a(i + 1), odom);
    vals.insert(b(i + 1), odom);
    odom_factors.add(BetweenFactor<Pose3>(a(i), a(i + 1), odom, noise));
    odom_factors.add(BetweenFactor<Pose3>(b(i), b(i + 1), odom, noise));
    pgo->update(odom_factors, vals);
  }
  for (size_t i = 3; i < 5; i++) {
    Values v; NonlinearFactorGraph f;
    Pose3 odom(R90, 1, 0, 0);
    v.insert(a(i + 1), odom);
    f.add(BetweenFactor<Pose3>(a(i), a(i + 1), odom, noise));
    pgo->update(f, v);
  }
  for (size_t i = 3; i < 5; i++) {
    Values v; NonlinearFactorGraph f;
    Pose3 odom(RM90, 1, 0, 0);
    v.insert(b(i + 1), odom);
    f.add(BetweenFactor<Pose3>(b(i), b(i + 1), odom, noise));
    pgo->update(f, v);
  }
  NonlinearFactorGraph lc_factors;
  lc_factors.add(BetweenFactor<Pose3>(b(1), a(1), Pose3::Translation(0, 1, 0), noise));
  lc_factors.add(BetweenFactor<Pose3>(b(4), a(4), Pose3(R180, -1, 0, 0), noise));
  pgo->update(lc_factors, Values());
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(12));
  EXPECT(pgo->calculateEstimate().size() == size_t(12));
  lc_factors = NonlinearFactorGraph();
  lc_factors.add(BetweenFactor<Pose3>(a(2), b(2), Pose3::Translation(0, -1, 0), noise));
  lc_factors.add(BetweenFactor<Pose3>(b(5), a(5), Pose3::Translation(0, -3.3, 0), noise));
  pgo->update(lc_factors, Values());
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(12));
  lc_factors = NonlinearFactorGraph();
  lc_factors.add(BetweenFactor<Pose3>(a(4), a(1), Pose3(RM90, 0, 3, 0), noise));
  lc_factors.add(BetweenFactor<Pose3>(a(4), a(2), Pose3(Pose3::Rz(-1.54), 0, 2, 0), noise));
  pgo->update(lc_factors, Values());
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(13));
  EXPECT(pgo->calculateEstimate().size() == size_t(12));
  // tests/testRemoveLastLoopClosure.cpp / testIgnorePrefix.cpp behaviours on the same graph
  const size_t before = pgo->getFactorsUnsafe().size();
  pgo->ignorePrefix('b');
  EXPECT(pgo->getFactorsUnsafe().size() < before);
  EXPECT((pgo->getIgnoredPrefixes() == std::vector<char>{'b'}));
  pgo->revivePrefix('b');
  EXPECT(pgo->getFactorsUnsafe().size() == before);
  EdgePtr e = pgo->removeLastLoopClosure('a', 'b');
  EXPECT(e != nullptr && e->from_key.chr() == 'b' && e->to_key.chr() == 'a' && e->from_key.index() == 5);
  EXPECT(pgo->removeLastLoopClosure('c', 'd') == nullptr);
}

// tests/testLandmark.cpp:23-185 (LandmarkPcm) and :186-348 (LandmarkPcmSimple)
static void RobustSolver_Landmark(bool simple) {
  RobustSolverParams params;
  if (simple) params.setPcmSimple3DParams(0.3, 0.05, Verbosity::QUIET);
  else params.setPcm3DParams(5.0, 2.5, Verbosity::QUIET);
  params.specialSymbols = std::vector<char>{'l'};
  std::unique_ptr<RobustSolver> pgo(new RobustSolver(params));
  const auto noise = IsotropicVariance(6, 0.1);
  auto a = [](size_t i) { return Key(Symbol('a', i)); };
  auto l = [](size_t i) { return Key(Symbol('l', i)); };
  Values init_vals;
  init_vals.insert(a(0), Pose3());
  pgo->update(NonlinearFactorGraph(), init_vals);
  for (size_t i = 0; i < 5; i++) {
    Values v; NonlinearFactorGraph f;
    Pose3 odom(i < 2 ? I3 : R90, 1, 0, 0);
    v.insert(a(i + 1), odom);
    f.add(BetweenFactor<Pose3>(a(i), a(i + 1), odom, noise));
    pgo->update(f, v);
  }
  NonlinearFactorGraph lc_factors;
  lc_factors.add(BetweenFactor<Pose3>(a(3), a(2), Pose3(Pose3::Rz(-1.57), 0, 0.9, 0), noise));
  lc_factors.add(BetweenFactor<Pose3>(a(4), a(1), Pose3(Pose3::Rz(3.14), 2.1, 1.1, 2.5), noise));
  pgo->update(lc_factors, Values());
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(6));
  EXPECT(pgo->calculateEstimate().size() == size_t(6));
  // Diagonal::Precisions((0,0,0,25,25,25)): no rotation information (NaN block), translation variance 1/25
  std::vector<double> lmk_cov(36, 0.0);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lmk_cov[i * 6 + j] = std::nan("");
  for (int i = 3; i < 6; ++i) lmk_cov[i * 6 + i] = 1.0 / 25.0;
  auto observe = [&](Key from, Key lm, double x, double y, double z, bool first) {
    NonlinearFactorGraph f; Values v;
    f.add(BetweenFactor<Pose3>(from, lm, Pose3::Translation(x, y, z), lmk_cov));
    if (first) v.insert(lm, Pose3::Translation(1, 1, 0));
    pgo->update(f, v);
  };
  observe(a(1), l(0), 0, 1, 0, true);
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(7));
  EXPECT(pgo->calculateEstimate().size() == size_t(7));
  observe(a(5), l(0), 0, -1, 0, false);   // consistent re-observation
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(8));
  observe(a(4), l(0), 1, 0, 0, false);    // inconsistent re-observation
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(8));
  EXPECT(pgo->calculateEstimate().size() == size_t(7));
  observe(a(2), l(1), 0, -1, 0, true);
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(9));
  EXPECT(pgo->calculateEstimate().size() == size_t(8));
  observe(a(5), l(1), 2, 0, 0, false);
  EXPECT(pgo->getFactorsUnsafe().size() == size_t(10));
  EXPECT(pgo->calculateEstimate().size() == size_t(8));
}

int main() {
  try {
    Pcm_OdometryCheck();
    Pcm_ConsistencyCheck();
    RobustSolver_multiRobot(false);
    RobustSolver_multiRobot(true);
    RobustSolver_Landmark(false);
    RobustSolver_Landmark(true);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf(failures ? "There were %d failures\n" : "There were no test failures (%d)\n", failures);
  return failures ? 1 : 0;
}
