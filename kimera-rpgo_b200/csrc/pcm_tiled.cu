/* pcm_tiled.cu — K3, TMA-staged shared-memory tiled variant (placeholder: forwards to the direct kernel). */
#include "kernels.cuh"
namespace rpgo {
void launch_pairwise_tiled(int dim, int mode, GroupView g, const double* traj, int j_begin, Shard sh, Thresholds th,
                           Flagged fl, cudaStream_t st) {
  launch_pairwise_direct(dim, mode, g, traj, j_begin, sh, th, fl, nullptr, st);
}
}  // namespace rpgo
