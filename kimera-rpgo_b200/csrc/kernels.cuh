/* kernels.cuh — internal launcher declarations shared by the .cu files and the C-ABI layer. */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "rpgo_math.cuh"
#include "../../include/rpgo_b200.h"

namespace rpgo {

/* cudaFuncSetAttribute is per device: remember which devices have opted a kernel into its dynamic shared memory */
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (mask.load(std::memory_order_relaxed) & bit) return false;
    mask.fetch_or(bit);
    return true; /* two racing threads both set the attribute: idempotent */
  }
};

/* one odometry step: entries[out_idx] = entries[prev_idx] (or the running value) . T(delta) */
struct FoldChain {
  int32_t first_step;  /* index of the chain's first step in the step arrays */
  int32_t n_steps;
  int32_t start_idx;   /* trajectory entry the chain starts from */
  int32_t pad;
};

/* device view of one ObservationId group */
struct GroupView {
  const double* lc;          /* n x ENTRY closure entries (AoS), after the from-factor constructor */
  const int32_t* idx_front;  /* trajectory entry index of key_from (0 = default/identity entry) */
  const int32_t* idx_back;   /* trajectory entry index of key_to */
  const uint8_t* pfx_front;  /* Symbol::chr() of key_from (decides the Pcm.h:691-698 key swap) */
  const int32_t* idx_a0;     /* landmark groups: trajectory entry of Symbol(chr(pose key), 0); else nullptr */
  uint32_t* bits;            /* adjacency bitset, row stride stride32 32-bit words */
  int64_t stride32;
  int32_t* deg;
  int32_t n;
};

struct Flagged {
  int32_t* pairs;                 /* 2 ints per flagged pair */
  unsigned long long* count;      /* total flagged (may exceed cap) */
  int64_t cap;
};

/* row sharding for multi-GPU: rows are cut into 2*world chunks of chunk_rows; this rank owns chunks
 * rank and 2*world-1-rank.  world == 1 => everything. */
struct Shard {
  int32_t rank, world;
  int64_t chunk_rows;
};

/* K1 */
void launch_traj_fold(int dim, int mode, int n_chains, const FoldChain* chains, const int32_t* out_idx,
                      const double* delta_pose, const double* delta_cov, double* entries, cudaStream_t st);
void launch_traj_scan_phases(int dim, int mode, int n_chains, const FoldChain* chains, const int32_t* out_idx,
                             const double* delta_pose, const double* delta_cov, double* entries, int chunk,
                             int total_steps, int total_chunks, const int32_t* chunk_chain, const int32_t* chunk_first,
                             const int32_t* chain_first_chunk, const int32_t* step_chunk, double* carry,
                             cudaStream_t st);
/* raw (pose, cov) -> entries; K2 odometry check for the ones with check[i] != 0 */
void launch_lc_prepare(int dim, int mode, int n, const double* pose, const double* cov, const int32_t* idx_front,
                       const int32_t* idx_back, const uint8_t* check, const double* traj, Thresholds th,
                       double* entries_out, uint8_t* ok_out, double* dist_out, cudaStream_t st);
void launch_scatter_entries(int dim, int n, const double* entries, const uint64_t* dst_ptrs, cudaStream_t st);
/* K3 */
void launch_pairwise_direct(int dim, int mode, GroupView g, const double* traj, int j_begin, Shard sh, Thresholds th,
                            Flagged fl, double* dist_out, cudaStream_t st);
int tiled_record_doubles(int dim, int mode);
int tiled_column_pitch(int dim, int mode); /* doubles per record of the per-lane column array */
void launch_gather_records(int dim, int mode, GroupView g, const double* traj, int k0, double* aos, double* soa, double* col,
                           cudaStream_t st);
/* variant: 0 = default (phase-shifted warp groups, straight-line pair function); the two cross-check forms
 * RPGO_KERNEL_TILED_ONE_GROUP / RPGO_KERNEL_TILED_V1 select 1 / 2 */
void launch_pairwise_tiled(int dim, int mode, GroupView g, const double* aos, const double* soa, const double* col, int j_begin,
                           Shard sh, Thresholds th, Flagged fl, int variant, cudaStream_t st);
/* N3: landmark re-observation matrix (one thread per pair of observations of the same landmark) */
void launch_landmark_direct(int dim, int mode, GroupView g, const double* traj, int j_begin, Thresholds th, Flagged fl,
                            double* dist_out, cudaStream_t st);
/* N4: frame-alignment measurements and transformed trajectories (pose-only batched kernels) */
void launch_frame_align(int dim, GroupView g, const double* traj, int entry, uint8_t r0, int m, const int32_t* closure_idx,
                        double* out, cudaStream_t st);
void launch_transform_poses(int dim, const double* traj, int entry, int m, const int32_t* entry_idx, const double* transform,
                            double* out, cudaStream_t st);
/* bitset maintenance */
void launch_mirror(uint32_t* bits, int64_t stride32, int n, int j_begin, cudaStream_t st);
void launch_degree(const uint32_t* bits, int64_t stride32, int n, int32_t* deg, cudaStream_t st);
void launch_clear_last(uint32_t* bits, int64_t stride32, int n_after, cudaStream_t st);

/* K4: heuristic clique.  Returns through host pointers (synchronises the stream). */
struct CliqueScratch {
  uint32_t* degmask;        /* stride32 words */
  int32_t* picks;           /* n */
  int32_t* elim;            /* n */
  int32_t* result;          /* n */
  long long* ctl;           /* control block (clique_kernels.cu): epoch words, statistics, result, per-block log selectors */
  uint32_t* rwork;          /* per-block pick logs: rwork_blocks x 2 x n ints */
  int64_t rwork_blocks;
};
/* bytes the control block needs for `blocks` thread blocks */
inline size_t clique_ctl_bytes(int64_t blocks) { return 256 + (size_t)blocks * 4 + 64; }
struct CliqueStats {
  long long row_ands = 0; /* adjacency-row ANDs of the heuristic: algorithmic bytes = row_ands * n / 8 */
  long long chains = 0;   /* greedy chains started */
  int epochs = 0;
};
/* candidate partition of the clique searches over ranks; the incumbent is combined over the handle's NCCL communicator
 * (rpgo_comm_init) or, when the caller brings its own collective, the host function of rpgo_set_exchange */
struct Comm;
struct CliqueShard {
  int rank = 0, world = 1;
  int (*exchange)(void* user, int32_t op, void* buf, int64_t count, int32_t root) = nullptr;
  void* user = nullptr;
  Comm* comm = nullptr;          /* NCCL communicator of the handle (rpgo_comm_init); preferred over `exchange` */
  cudaStream_t comm_stream = nullptr;
  bool active() const { return world > 1 && (comm != nullptr || exchange != nullptr); }
  /* in-place collective on a HOST buffer, same on every rank: RPGO_XCHG_* */
  int xchg(int32_t op, void* buf, int64_t count, int32_t root) const;
};
int clique_heuristic(const uint32_t* bits, int64_t stride32, int n, const int32_t* deg, int first, int maxclq0,
                     CliqueScratch s, int32_t* ids_out_host, int32_t* true_out_host, int64_t* launches,
                     cudaStream_t st, CliqueShard cs = CliqueShard(), CliqueStats* stats_out = nullptr);
int clique_exact(const uint32_t* bits, int64_t stride32, int n, const int32_t* deg, CliqueScratch s,
                 int32_t* ids_out_host, int64_t* launches, cudaStream_t st, CliqueShard cs = CliqueShard());

double fp64_peak_tflops(cudaStream_t st);
int fastmath_check(long long n, unsigned long long seed, unsigned long long* mismatches, unsigned long long* checked,
                   cudaStream_t st);

}  // namespace rpgo
