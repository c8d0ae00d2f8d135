/* rpgo_b200.h — C ABI of the B200-native PCM outlier-rejection hot path.
 *
 * This is the drop-in boundary for the one data-parallel path of MIT-SPARK/Kimera-RPGO that this
 * library rebuilds: the three calls Pcm<poseT,T>::removeOutliers makes into arithmetic —
 *   updateOdom            (reference include/KimeraRPGO/outlier/Pcm.h:516-557)   -> rpgo_odom_append
 *   isOdomConsistent +
 *   incrementAdjMatrix    (Pcm.h:604-629, :725-768, via parseAndIncrementAdjMatrix :414-502)
 *                                                                                 -> rpgo_lc_append
 *   findMaxCliqueHeu / findMaxCliqueHeuIncremental / findMaxClique
 *                         (src/utils/GraphUtils.cpp:9-44, called from Pcm.h:865, :922, :328)
 *                                                                                 -> rpgo_find_inliers
 *   removeLastLoopClosure's matrix shrink (Pcm.h:320-323)                         -> rpgo_lc_remove_last
 * Factor classification, shared_ptr bookkeeping and output-graph assembly stay on the host in the
 * OutlierRemoval subclass that calls this ABI.
 *
 * Conventions: plain C types only; all pointers are HOST pointers to caller-owned buffers unless a
 * name ends in _device; device memory is owned by the library; every function returns an int status
 * (RPGO_OK == 0) and never throws; a handle is single-threaded (one CUDA stream, one GPU).
 * There is no CPU fallback: rpgo_create fails with RPGO_ERR_CUDA when the selected device is not an sm_100 GPU.
 * Every entry point switches to the handle's device for the duration of the call and restores the caller's current
 * device; device memory comes from a library-private stream-ordered pool (the application's default pool is untouched).
 *
 * Data layout:
 *   key        uint64, gtsam::Key (gtsam::Symbol: chr = key >> 56, index = low 56 bits)
 *   pose  3D   12 doubles: R row-major (9) then t (3)          2D   4 doubles: cos, sin, x, y
 *   cov   3D   6x6 row-major, GTSAM tangent order [rot; trans]  2D   3x3 row-major, (x, y, theta)
 *         (== noiseModel::Gaussian::covariance() of the factor; a NaN rotation block selects the
 *          reference's rotation_info=false path, GeometryUtils.h:98-113)
 *   adjacency  bitset, row i = 64-bit little-endian words, bit j of row i <=> closures i and j of the
 *              group are pairwise consistent; symmetric, zero diagonal (Pcm.h:756-763)
 */
#ifndef RPGO_B200_H_
#define RPGO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPGO_OK 0
#define RPGO_ERR_INVALID 1   /* bad argument */
#define RPGO_ERR_CUDA 2      /* CUDA runtime error / no device */
#define RPGO_ERR_NOMEM 3
#define RPGO_ERR_NOT_FOUND 4 /* unknown group / key */

#define RPGO_MODE_PCM 0      /* Pcm2D / Pcm3D: Mahalanobis distance  (Pcm.h:1167-1168) */
#define RPGO_MODE_SIMPLE 1   /* PcmSimple2D / PcmSimple3D: average trans/rot per node (Pcm.h:1169-1170) */

#define RPGO_CLIQUE_HEU 0              /* findMaxCliqueHeu            GraphUtils.cpp:19-27 */
#define RPGO_CLIQUE_HEU_INCREMENTAL 1  /* findMaxCliqueHeuIncremental GraphUtils.cpp:30-44 */
#define RPGO_CLIQUE_EXACT 2            /* findMaxClique               GraphUtils.cpp:9-17  */

#define RPGO_TRAJ_FOLD 0  /* strict left fold, bit-identical to the reference's order (Pcm.h:553) */
#define RPGO_TRAJ_SCAN 1  /* chunked parallel prefix scan (re-associated; deterministic) */

#define RPGO_KERNEL_AUTO 0
#define RPGO_KERNEL_DIRECT 1  /* v0: one thread per pair, operands gathered from global memory */
#define RPGO_KERNEL_TILED 2   /* TMA-staged shared-memory tiles (what AUTO selects) */
/* cross-check forms of the tiled kernel, used by the parity tests only: one warp group with the straight-line pair
 * function, and one warp group with the plain (branchy) pair function */
#define RPGO_KERNEL_TILED_ONE_GROUP 24
#define RPGO_KERNEL_TILED_V1 22

typedef struct rpgo_handle rpgo_handle;

/* Mirrors KimeraRPGO::PcmParams (reference include/KimeraRPGO/SolverParams.h:33-56). A threshold < 0
 * disables the corresponding check exactly as Pcm.h:74-82 does. */
typedef struct rpgo_cfg {
  int32_t dim;   /* 2 or 3 */
  int32_t mode;  /* RPGO_MODE_* */
  double odom_threshold;
  double lc_threshold;
  double odom_trans_threshold;
  double odom_rot_threshold;
  double dist_trans_threshold;
  double dist_rot_threshold;
  int32_t incremental; /* PcmParams::incremental (informational; the caller picks the clique mode) */
  int32_t device;      /* CUDA ordinal, -1 = current device */
  int32_t traj_mode;   /* RPGO_TRAJ_* */
  int32_t kernel;      /* RPGO_KERNEL_* for the pairwise kernel */
  int32_t rank;        /* multi-GPU: this handle computes row chunks {rank, 2*world-1-rank} of every group */
  int32_t world;       /* 1 = single GPU */
  double band;         /* near-threshold flag half width; <= 0 selects 1e-9 */
  int32_t scan_chunk;  /* RPGO_TRAJ_SCAN chunk length; <= 0 selects 64 */
  int32_t reserved;
} rpgo_cfg;

int rpgo_default_cfg(rpgo_cfg* cfg);
int rpgo_create(const rpgo_cfg* cfg, rpgo_handle** out);
void rpgo_destroy(rpgo_handle* h);
/* back to the freshly created state (= a new Pcm object, Pcm.h:64-96), keeping stream, device arena, pinned staging and
 * communicator: what a long-lived solver calls between independent graphs */
int rpgo_reset(rpgo_handle* h);
const char* rpgo_last_error(const rpgo_handle* h);
/* block until all work queued on the handle's stream has finished */
int rpgo_sync(rpgo_handle* h);
/* the handle's cudaStream_t (as an opaque pointer), for event timing and collectives issued by the caller */
void* rpgo_stream(rpgo_handle* h);

/* ---- K1: odometry trajectory cache (Pcm::updateOdom, Pcm.h:516-557) --------------------------
 * Appends n odometry factors in arrival order: poses[new_key] = poses[prev_key].compose(T(delta)).
 * init_pose (n poses, may be NULL = identity) is consulted only for a factor that introduces a new
 * prefix: it is values.at(prev_key), which seeds the trajectory with zero covariance (:534-542). */
int rpgo_odom_append(rpgo_handle* h, int64_t n, const uint64_t* prev_key, const uint64_t* new_key,
                     const double* delta_pose, const double* delta_cov, const double* init_pose);

/* parity / debug: cumulative entry of `key` (cov may be NULL; node/rot_info may be NULL) */
int rpgo_traj_get(rpgo_handle* h, uint64_t key, double* pose, double* cov, int32_t* node, int32_t* rot_info);
int64_t rpgo_traj_size(rpgo_handle* h);

/* ---- K2 + K3: loop closures (Pcm::parseAndIncrementAdjMatrix for pose-pose closures, :456-494) -
 * For each of the n closures in arrival order: intra-robot closures are checked against odometry
 * (isOdomConsistent; rejected ones are dropped and not counted); accepted closures join the group of
 * their unordered prefix pair (ObservationId, TypeUtils.h:44-58) and the group's adjacency is extended by
 * one row/column per closure (incrementAdjMatrix).
 * Outputs (each may be NULL): accepted[i] 0/1; group[i] group ordinal (first-seen order) or -1;
 * index[i] position inside the group or -1; odom_dist[i] the odometry-check distance (NaN if not run). */
int rpgo_lc_append(rpgo_handle* h, int64_t n, const uint64_t* key_from, const uint64_t* key_to,
                   const double* pose, const double* cov, uint8_t* accepted, int32_t* group, int32_t* index,
                   double* odom_dist);

/* ---- N3: landmark observations (Pcm.h:207-220 first observation; :437-455 + incrementLandmarkAdjMatrix
 * :775-844 re-observations) -----------------------------------------------------------------------------
 * Appends n observations pose_key[i] -> landmark_key (measured pose / covariance as in rpgo_lc_append) to the
 * landmark's own group and extends its adjacency: observations (i -> l), (j -> l) are consistent iff
 * (getBetween(i, j) . j_pose_l)^-1 . i_pose_l passes the loop-consistency check (getBetween's different-prefix
 * path included, as the reference runs it).  reset != 0 empties the group first: a FIRST_LANDMARK_OBSERVATION
 * replaces landmarks_[key] (Pcm.h:218).  *group_out is an ordinal for rpgo_find_inliers / rpgo_adj_bits /
 * rpgo_group_info (which reports id1 = id2 = chr(landmark_key)). */
int rpgo_landmark_append(rpgo_handle* h, uint64_t landmark_key, int64_t n, const uint64_t* pose_key, const double* pose,
                         const double* cov, int32_t reset, int32_t* group_out);

int32_t rpgo_num_groups(rpgo_handle* h);
/* prefixes (id1 <= id2) and number of stored closures of group g */
int rpgo_group_info(rpgo_handle* h, int32_t g, uint8_t* id1, uint8_t* id2, int64_t* n);
/* group ordinal of the unordered prefix pair, or -1 */
int32_t rpgo_find_group(rpgo_handle* h, uint8_t id1, uint8_t id2);

/* Pcm::removeLastLoopClosure's data part (Pcm.h:314-323): drop the last closure of group g (last row and
 * column of its adjacency). key_from/key_to receive the removed edge. RPGO_ERR_NOT_FOUND if empty. */
int rpgo_lc_remove_last(rpgo_handle* h, int32_t g, uint64_t* key_from, uint64_t* key_to);

/* ---- K4 / K5: inlier selection ------------------------------------------------------------------
 * Runs the reference's clique finder on group g's adjacency.  ids_out (capacity >= group size) receives
 * exactly what Pcm.h:865-869 consumes: the first *size_out entries of the vector the reference returns —
 * for the heuristic modes that is FMC's scratch buffer (findCliqueHeu.cpp:110-113), reproduced verbatim,
 * which is generally not the greedy clique itself; true_clique_out (may be NULL, same capacity)
 * receives the vertices the greedy chain actually selected.
 * HEU_INCREMENTAL: candidates are the last n_new closures, bound starts at prev_size; *size_out = 0 when
 * the clique did not grow (caller keeps its previous inliers, Pcm.h:929-939). */
int rpgo_find_inliers(rpgo_handle* h, int32_t g, int32_t clique_mode, int64_t n_new, int64_t prev_size,
                      int32_t* ids_out, int64_t* size_out, int32_t* true_clique_out);

/* Batched form for the groups of one removeOutliers() call (Pcm::findInliers loops over all ObservationId groups,
 * Pcm.h:858-876; the groups are independent).  Entry k searches group groups[k] with n_new[k] / prev_size[k]
 * (both arrays may be NULL for the non-incremental modes) and writes its ids to ids_out + ids_offset[k] (capacity >=
 * that group's size) and its size to size_out[k] — same values rpgo_find_inliers returns.  The searches run
 * concurrently on internal streams.  With cfg.world > 1 and an exchange function registered, whole groups are assigned
 * to ranks (entry k to rank k mod world) and the results are combined with one all-reduce. */
int rpgo_find_inliers_batch(rpgo_handle* h, int32_t n_groups, const int32_t* groups, int32_t clique_mode,
                            const int64_t* n_new, const int64_t* prev_size, int32_t* ids_out, const int64_t* ids_offset,
                            int64_t* size_out);

/* ---- multi-GPU inlier selection with a caller-provided collective -------------------------------------------
 * Alternative to rpgo_comm_init for applications that cannot use NCCL (the gloo CPU tests do this).
 * With cfg.world > 1 and an exchange function registered, rpgo_find_inliers partitions the clique search's root
 * candidates over the ranks (candidate v belongs to rank v mod world in the heuristic, root n-1-v likewise in the
 * exact search) and calls `fn` on the host, on every rank in the same order, to combine the incumbent:
 *   RPGO_XCHG_MIN_I64 / RPGO_XCHG_MAX_I64: in-place all-reduce of count int64 values in buf;
 *   RPGO_XCHG_BCAST_I32: broadcast count int32 values in buf from rank `root`.
 * fn returns 0 on success.  The library stays free of a communication dependency: the caller implements fn with
 * whatever it already uses (the Python harness: torch.distributed over NCCL/NVLink).  Without a registered
 * function every rank searches all candidates (replicated, same result).  The reference has no counterpart:
 * its clique search is single-threaded (GraphUtils.cpp:9-44). */
#define RPGO_XCHG_MIN_I64 0
#define RPGO_XCHG_MAX_I64 1
#define RPGO_XCHG_BCAST_I32 2
typedef int (*rpgo_exchange_fn)(void* user, int32_t op, void* buf, int64_t count, int32_t root);
int rpgo_set_exchange(rpgo_handle* h, rpgo_exchange_fn fn, void* user);

/* ---- multi-GPU data plane behind the ABI (the reference, Pcm.h, is single-process) ---------------
 * One process (or thread) per GPU, each with its own handle created with cfg.rank / cfg.world.  Rank 0 calls
 * rpgo_comm_unique_id and hands the RPGO_COMM_ID_BYTES bytes to the other ranks by whatever means the application has
 * (MPI, a file, torch.distributed's store); then every rank calls rpgo_comm_init, which builds an NCCL communicator over
 * the handles' GPUs (collective call: all ranks must enter it).  From then on
 *   - rpgo_lc_append computes only this rank's row chunks of every touched group and all-gathers the adjacency rows
 *     with ncclAllGather / grouped ncclBroadcast on the handle's stream (NVLink / NVSwitch), so that every rank ends
 *     with the full symmetric bitset and degrees;
 *   - rpgo_find_inliers partitions the clique search's root candidates over the ranks and combines the incumbent with
 *     ncclAllReduce; rpgo_find_inliers_batch spreads whole groups over the ranks;
 *   - every rank returns identical results.
 * Every rank must issue the same sequence of calls with the same data (the tables are replicated).
 * NCCL is bound at run time (dlopen libnccl.so.2); rpgo_comm_init fails with RPGO_ERR_CUDA when it is not installed. */
#define RPGO_COMM_ID_BYTES 128
int rpgo_comm_unique_id(void* id_out);
int rpgo_comm_init(rpgo_handle* h, const void* id, int32_t rank, int32_t world);
int rpgo_comm_destroy(rpgo_handle* h);
/* measurement hook: all-gather group g's row chunks and rebuild mirror + degrees (what rpgo_lc_append does after K3) */
int rpgo_group_allgather(rpgo_handle* h, int32_t g);

/* ---- N4: multi-robot frame alignment, the batched front half (Pcm::multirobotValueInitialization,
 * Pcm.h:1024-1055; the GNC pose averaging that follows stays on GTSAM) -------------------------------------
 * For the m closures closure_idx[] of group g (normally its inliers) writes T_w0_wi = T_w0_front . T_front_back .
 * T_wi_back^-1, swapping the keys and inverting the measurement when the closure is stated ri -> r0 (:1043-1048).
 * Returns RPGO_ERR_NOT_FOUND if a key has no trajectory entry (the reference's .at() throws and the robot is
 * skipped, :1060-1064). */
int rpgo_frame_align_measurements(rpgo_handle* h, int32_t g, uint8_t r0, int64_t m, const int32_t* closure_idx,
                                  double* T_w0_wi_out);
/* getRobotOdomValues (Pcm.h:1074-1082): transform . pose for every trajectory entry of `prefix` in ascending
 * key order (transform may be NULL = identity).  Call with keys_out = poses_out = NULL to get the count. */
int rpgo_robot_odom_values(rpgo_handle* h, uint8_t prefix, const double* transform, int64_t cap, uint64_t* keys_out,
                           double* poses_out, int64_t* n_out);

/* ---- parity / inspection ---------------------------------------------------------------------- */
/* copy group g's adjacency to the host: n rows of stride_words 64-bit words (stride_words >= ceil(n/64)) */
int rpgo_adj_bits(rpgo_handle* h, int32_t g, uint64_t* rows_out, int64_t stride_words);
/* device view of the same bitset (valid until the next call that grows the group) */
int rpgo_adj_bits_device(rpgo_handle* h, int32_t g, void** bits_device, int64_t* stride_words, int64_t* n);
/* vertex degrees (popcount of each row) */
int rpgo_degrees(rpgo_handle* h, int32_t g, int32_t* deg_out);
/* pairs (i, j), i < j, whose distance lies within cfg.band of the threshold (the explicitly flagged
 * near-threshold set).  pairs_out holds up to cap pairs as 2 ints each; *n_out = total flagged.  With a communicator
 * (rows sharded over ranks) the call is collective and returns the union over the ranks. */
int rpgo_near_threshold(rpgo_handle* h, int32_t g, int32_t* pairs_out, int64_t cap, int64_t* n_out);
/* debug: recompute distances of all pairs of group g into an n x n row-major host matrix (small n only) */
int rpgo_pair_distances(rpgo_handle* h, int32_t g, double* dist_out);

/* ---- measurement hooks --------------------------------------------------------------------------
 * Re-run the pairwise kernel over columns [j_begin, n) of group g from the device-resident tables
 * (inputs already in HBM; same launch rpgo_lc_append issues).  Used by bench.py for kernel-only timing. */
int rpgo_group_recompute(rpgo_handle* h, int32_t g, int64_t j_begin);
/* the pairwise kernel alone (no mirror / degree pass): what roofline.achieved is measured on */
int rpgo_group_pairwise(rpgo_handle* h, int32_t g, int64_t j_begin);
/* multi-GPU: after the caller has all-gathered the upper-triangle row chunks, rebuild the lower
 * triangle and the degrees on this GPU */
int rpgo_group_finalize(rpgo_handle* h, int32_t g);
/* statistics of the last heuristic search of rpgo_find_inliers on this rank: adjacency-row ANDs (one n/8-byte row per
 * greedy step plus the initial row of every evaluated candidate: the algorithmic bytes of SURVEY §8(d) are row_ands * n / 8),
 * greedy chains started, and epochs (dependent passes over the candidates) */
int rpgo_clique_stats(rpgo_handle* h, int64_t* row_ands, int64_t* chains, int32_t* epochs);
/* one bitset pass alone, for bandwidth measurements: which = 0 mirror (lower triangle from the upper one), 1 degrees */
int rpgo_debug_pass(rpgo_handle* h, int32_t g, int32_t which);
/* row-chunk geometry of group g for the all-gather: rows are cut into 2*world chunks of chunk_rows rows */
int rpgo_group_chunking(rpgo_handle* h, int32_t g, int64_t* chunk_rows, int64_t* padded_rows);
/* number of kernels this handle has launched since creation */
int64_t rpgo_launch_count(rpgo_handle* h);
/* FP64 FMA micro-benchmark used as the roofline denominator: returns achieved DFMA TFLOP/s */
int rpgo_fp64_peak(int32_t device, double* tflops_out);

/* test / benchmark hook: create (or replace) the group of prefix pair (id1, id2) with n vertices and the
 * given symmetric adjacency (n rows of stride_words 64-bit words) without running the pairwise kernel,
 * so that the clique kernels can be exercised on arbitrary graphs. */
int rpgo_debug_load_group(rpgo_handle* h, uint8_t id1, uint8_t id2, int64_t n, const uint64_t* rows,
                          int64_t stride_words, int32_t* group_out);

/* test hook: compare the branch-free reciprocal / division / square-root sequences used by the straight-line
 * pair kernel with the built-in IEEE operations on n random operand pairs; *mismatches must come back 0 */
int rpgo_debug_check_fastmath(int64_t n, uint64_t seed, uint64_t* mismatches, uint64_t* checked);

const char* rpgo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RPGO_B200_H_ */
