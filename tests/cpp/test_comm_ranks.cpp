// C++ multi-rank test of the multi-GPU data plane behind the C ABI (include/rpgo_b200.h: rpgo_comm_unique_id /
// rpgo_comm_init): no Python, no torch.  The parent forks one process per rank BEFORE any CUDA call; rank 0 creates
// the NCCL unique id and hands it to the others through pipes (the "whatever means the application has" of the header).
// Every rank then
//   1. builds the same synthetic 3D pose graph (one robot, helix odometry, noisy inlier closures + random outliers),
//   2. runs it through a sharded handle (cfg.rank / cfg.world, rows of the pair matrix split over the ranks, NCCL
//      all-gather of the adjacency inside rpgo_lc_append, clique candidates partitioned with an NCCL all-reduce),
//   3. runs it through a plain single-GPU handle on the same device,
//   4. demands identical adjacency bitsets, degrees, flagged pairs and inlier ids.
// Usage: _test_comm_ranks [world=2] [closures=3000]; exit code 0 = all ranks agree, 77 = fewer GPUs than ranks.
//   g++ -std=c++17 -Iinclude tests/cpp/test_comm_ranks.cpp -Lkimera-rpgo_b200 -lrpgo_b200 -ldl
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rpgo_b200.h"

namespace {

struct Rng {  // splitmix64
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double gauss() {
    const double u = uni() + 1e-300, v = uni();
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
  }
};

struct Pose {
  double R[9], t[3];
};
Pose identity() { return Pose{{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}}; }
Pose mul(const Pose& a, const Pose& b) {
  Pose r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[i * 3 + j] = a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j] + a.R[i * 3 + 2] * b.R[6 + j];
  for (int i = 0; i < 3; ++i) r.t[i] = a.t[i] + a.R[i * 3] * b.t[0] + a.R[i * 3 + 1] * b.t[1] + a.R[i * 3 + 2] * b.t[2];
  return r;
}
Pose inv(const Pose& a) {
  Pose r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) r.t[i] = -(r.R[i * 3] * a.t[0] + r.R[i * 3 + 1] * a.t[1] + r.R[i * 3 + 2] * a.t[2]);
  return r;
}
Pose expw(double wx, double wy, double wz, double tx, double ty, double tz) {  // rotation exp(w) with translation t
  const double th = std::sqrt(wx * wx + wy * wy + wz * wz);
  const double a = th < 1e-12 ? 1.0 : std::sin(th) / th, b = th < 1e-12 ? 0.5 : (1 - std::cos(th)) / (th * th);
  const double K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  Pose r = identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double kk = 0;
      for (int k = 0; k < 3; ++k) kk += K[i * 3 + k] * K[k * 3 + j];
      r.R[i * 3 + j] += a * K[i * 3 + j] + b * kk;
    }
  r.t[0] = tx; r.t[1] = ty; r.t[2] = tz;
  return r;
}
void put(std::vector<double>& v, const Pose& p) {
  v.insert(v.end(), p.R, p.R + 9);
  v.insert(v.end(), p.t, p.t + 3);
}

struct Graph {
  std::vector<uint64_t> o_prev, o_new, l_from, l_to;
  std::vector<double> o_pose, o_cov, o_init, l_pose, l_cov;
};

Graph make_graph(int P, int n, uint64_t seed) {
  Graph g;
  Rng rng{seed};
  const uint64_t pfx = (uint64_t)'a' << 56;
  std::vector<Pose> truth(P);
  truth[0] = identity();
  const Pose step = expw(0.01, 0.0, 2 * 3.141592653589793 / 50, 1.0, 0.0, 0.02);
  for (int k = 1; k < P; ++k) truth[k] = mul(truth[k - 1], step);
  double ocov[36] = {0}, lcov[36] = {0};
  for (int i = 0; i < 3; ++i) { ocov[i * 7] = 1e-4; ocov[(i + 3) * 7] = 1e-3; lcov[i * 7] = 1e-3; lcov[(i + 3) * 7] = 1e-2; }
  Pose dead = identity();
  for (int k = 0; k + 1 < P; ++k) {
    const Pose noise = expw(0.01 * rng.gauss(), 0.01 * rng.gauss(), 0.01 * rng.gauss(), 0.0316 * rng.gauss(), 0.0316 * rng.gauss(),
                            0.0316 * rng.gauss());
    const Pose d = mul(step, noise);
    g.o_prev.push_back(pfx | (uint64_t)k);
    g.o_new.push_back(pfx | (uint64_t)(k + 1));
    put(g.o_pose, d);
    g.o_cov.insert(g.o_cov.end(), ocov, ocov + 36);
    put(g.o_init, dead);
    dead = mul(dead, d);
  }
  for (int q = 0; q < n; ++q) {
    int i = (int)(rng.next() % P), j = (int)(rng.next() % P);
    while (std::abs(i - j) <= 20) j = (int)(rng.next() % P);
    Pose m;
    if (rng.uni() < 0.5) {
      m = expw(3 * rng.gauss(), 3 * rng.gauss(), 3 * rng.gauss(), 20 * rng.uni() - 10, 20 * rng.uni() - 10, 20 * rng.uni() - 10);
    } else {
      const Pose noise = expw(0.0316 * rng.gauss(), 0.0316 * rng.gauss(), 0.0316 * rng.gauss(), 0.1 * rng.gauss(), 0.1 * rng.gauss(),
                              0.1 * rng.gauss());
      m = mul(mul(inv(truth[i]), truth[j]), noise);
    }
    g.l_from.push_back(pfx | (uint64_t)i);
    g.l_to.push_back(pfx | (uint64_t)j);
    put(g.l_pose, m);
    g.l_cov.insert(g.l_cov.end(), lcov, lcov + 36);
  }
  return g;
}

#define CHECK_RC(h, call)                                                                                  \
  do {                                                                                                     \
    const int rc_ = (call);                                                                                \
    if (rc_ != RPGO_OK) {                                                                                  \
      std::printf("[rank %d] %s failed (%d): %s\n", rank, #call, rc_, (h) ? rpgo_last_error(h) : "");      \
      return 1;                                                                                            \
    }                                                                                                      \
  } while (0)

struct Result {
  std::vector<uint64_t> bits;
  std::vector<int32_t> deg, ids, exact_ids;
  int64_t size = 0, exact_size = 0, flagged = 0;
};

int run_graph(rpgo_handle* h, int rank, const Graph& g, int n_first, Result* out) {
  const int64_t P1 = (int64_t)g.o_prev.size(), n = (int64_t)g.l_from.size();
  CHECK_RC(h, rpgo_odom_append(h, P1, g.o_prev.data(), g.o_new.data(), g.o_pose.data(), g.o_cov.data(), g.o_init.data()));
  // two batches: the second one grows the sharded matrix incrementally
  CHECK_RC(h, rpgo_lc_append(h, n_first, g.l_from.data(), g.l_to.data(), g.l_pose.data(), g.l_cov.data(), nullptr, nullptr, nullptr, nullptr));
  CHECK_RC(h, rpgo_lc_append(h, n - n_first, g.l_from.data() + n_first, g.l_to.data() + n_first, g.l_pose.data() + (size_t)n_first * 12,
                             g.l_cov.data() + (size_t)n_first * 36, nullptr, nullptr, nullptr, nullptr));
  int64_t m = 0;
  CHECK_RC(h, rpgo_group_info(h, 0, nullptr, nullptr, &m));
  const int64_t sw = (m + 63) / 64;
  out->bits.assign((size_t)m * sw, 0);
  out->deg.assign((size_t)m, 0);
  out->ids.assign((size_t)m, 0);
  out->exact_ids.assign((size_t)m, 0);
  CHECK_RC(h, rpgo_adj_bits(h, 0, out->bits.data(), sw));
  CHECK_RC(h, rpgo_degrees(h, 0, out->deg.data()));
  CHECK_RC(h, rpgo_near_threshold(h, 0, nullptr, 0, &out->flagged));
  CHECK_RC(h, rpgo_find_inliers(h, 0, RPGO_CLIQUE_HEU, 0, 0, out->ids.data(), &out->size, nullptr));
  out->ids.resize((size_t)out->size);
  return 0;
}

int rank_main(int rank, int world, int n, const int* rd, const int* wr) {
  rpgo_cfg cfg;
  rpgo_default_cfg(&cfg);
  cfg.odom_threshold = -1.0;
  cfg.lc_threshold = 5.0;
  cfg.device = rank;
  cfg.rank = rank;
  cfg.world = world;
  rpgo_handle* hs = nullptr;
  int rc = rpgo_create(&cfg, &hs);
  if (rc != RPGO_OK) {
    std::printf("[rank %d] rpgo_create failed (%d)\n", rank, rc);
    return rc == RPGO_ERR_CUDA || rc == RPGO_ERR_INVALID ? 77 : 1;
  }
  unsigned char id[RPGO_COMM_ID_BYTES];
  if (rank == 0) {
    if (rpgo_comm_unique_id(id) != RPGO_OK) { std::printf("[rank 0] rpgo_comm_unique_id failed\n"); return 1; }
    for (int r = 1; r < world; ++r)
      if (write(wr[r], id, sizeof(id)) != (ssize_t)sizeof(id)) return 1;
  } else {
    size_t got = 0;
    while (got < sizeof(id)) {
      const ssize_t k = read(rd[rank], id + got, sizeof(id) - got);
      if (k <= 0) return 1;
      got += (size_t)k;
    }
  }
  CHECK_RC(hs, rpgo_comm_init(hs, id, rank, world));
  cfg.rank = 0;
  cfg.world = 1;
  rpgo_handle* h1 = nullptr;
  CHECK_RC(h1, rpgo_create(&cfg, &h1));

  const Graph g = make_graph(std::max(400, n), n, 20261017ULL);
  int bad = 0;
  for (int pass = 0; pass < 2; ++pass) {  // second pass after rpgo_reset: same communicator, same arena
    Result a, b;
    if (run_graph(hs, rank, g, n / 2 + 3, &a) || run_graph(h1, rank, g, n / 2 + 3, &b)) return 1;
    const bool same = a.bits == b.bits && a.deg == b.deg && a.flagged == b.flagged && a.size == b.size && a.ids == b.ids;
    std::printf("[rank %d] pass %d: %lld closures, %lld flagged, clique %lld: sharded %s single\n", rank, pass,
                (long long)a.deg.size(), (long long)a.flagged, (long long)a.size, same ? "==" : "!=");
    if (!same) ++bad;
    CHECK_RC(hs, rpgo_reset(hs));
    CHECK_RC(h1, rpgo_reset(h1));
  }
  rpgo_destroy(h1);
  rpgo_destroy(hs);
  return bad ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const int n = argc > 2 ? std::atoi(argv[2]) : 3000;
  if (world < 2 || world > 16) return 2;
  {
    /* probe in a throw-away child (the parent must stay CUDA-free for fork): are there `world` usable sm_100 GPUs? */
    const pid_t p = fork();
    if (p < 0) return 2;
    if (p == 0) {
      int rc = 0;
      for (int r = 0; r < world && rc == 0; ++r) {
        rpgo_cfg cfg;
        rpgo_default_cfg(&cfg);
        cfg.device = r;
        rpgo_handle* h = nullptr;
        if (rpgo_create(&cfg, &h) != RPGO_OK) rc = 77;
        else rpgo_destroy(h);
      }
      _exit(rc);
    }
    int st = 0;
    waitpid(p, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
      std::printf("COMM_RANKS SKIP (not enough sm_100 GPUs; there is no CPU fallback)\n");
      return 77;
    }
  }
  std::vector<int> rd(world, -1), wr(world, -1);
  for (int r = 1; r < world; ++r) {
    int fd[2];
    if (pipe(fd) != 0) return 2;
    rd[r] = fd[0];
    wr[r] = fd[1];
  }
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {
    const pid_t p = fork();
    if (p < 0) return 2;
    if (p == 0) {
      const int rc = rank_main(r, world, n, rd.data(), wr.data());
      std::fflush(stdout);
      _exit(rc);
    }
    kids.push_back(p);
  }
  int worst = 0;
  for (pid_t p : kids) {
    int st = 0;
    waitpid(p, &st, 0);
    const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
    if (rc != 0 && (worst == 0 || worst == 77)) worst = rc;
  }
  std::printf(worst == 0 ? "COMM_RANKS PASS\n" : worst == 77 ? "COMM_RANKS SKIP (not enough sm_100 GPUs)\n" : "COMM_RANKS FAIL (%d)\n", worst);
  return worst;
}
