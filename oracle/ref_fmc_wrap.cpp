// C entry points around the reference's own FMC library (compiled from /root/reference by
// oracle/Makefile).  Mirrors the three wrappers in src/utils/GraphUtils.cpp:9-44 so that tests can
// call exactly what Pcm.h calls.  TEST INFRASTRUCTURE ONLY.
#include <vector>
#include "KimeraRPGO/max_clique_finder/findClique.h"

static Eigen::MatrixXd to_mat(int n, const unsigned char* adj) {
  Eigen::MatrixXd m = Eigen::MatrixXd::Zero(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) m(i, j) = adj[(size_t)i * n + j] ? 1.0 : 0.0;
  return m;
}

extern "C" {
// returns the clique size; buf receives the returned vector (up to cap entries); *buf_len its length
int ref_find_max_clique_heu(int n, const unsigned char* adj, int* buf, int cap, int* buf_len) {
  FMC::CGraphIO gio;
  gio.ReadEigenAdjacencyMatrix(to_mat(n, adj));
  std::vector<int> out;
  int k = FMC::maxCliqueHeu(&gio, &out);
  *buf_len = (int)out.size();
  for (int i = 0; i < (int)out.size() && i < cap; ++i) buf[i] = out[i];
  return k;
}
int ref_find_max_clique_heu_incremental(int n, const unsigned char* adj, int num_new, int prev, int* buf, int cap,
                                        int* buf_len) {
  FMC::CGraphIO gio;
  gio.ReadEigenAdjacencyMatrix(to_mat(n, adj));
  std::vector<int> out;
  int k = FMC::maxCliqueHeuIncremental(&gio, (size_t)num_new, (size_t)prev, &out);
  *buf_len = (int)out.size();
  for (int i = 0; i < (int)out.size() && i < cap; ++i) buf[i] = out[i];
  if ((size_t)k > (size_t)prev) return k;
  return 0;
}
int ref_find_max_clique(int n, const unsigned char* adj, int* buf, int cap, int* buf_len) {
  FMC::CGraphIO gio;
  gio.ReadEigenAdjacencyMatrix(to_mat(n, adj));
  std::vector<int> out;
  int k = FMC::maxClique(&gio, 0, &out);
  *buf_len = (int)out.size();
  for (int i = 0; i < (int)out.size() && i < cap; ++i) buf[i] = out[i];
  return k;
}
}
