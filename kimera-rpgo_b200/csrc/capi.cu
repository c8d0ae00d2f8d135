/* capi.cu — host-side state and the C ABI (include/rpgo_b200.h) of the B200 PCM path.
 *
 * Host logic restated here (reference file:line):
 *   trajectory bookkeeping of Pcm::updateOdom                        Pcm.h:516-557
 *   per-closure flow of Pcm::parseAndIncrementAdjMatrix              Pcm.h:456-494
 *   ObservationId grouping                                            TypeUtils.h:44-58, Pcm.h:472-486
 *   std::map::operator[] default-entry semantics of Trajectory       GraphUtils.h:40-42
 *   removeLastLoopClosure's matrix shrink                             Pcm.h:314-323
 * All arithmetic runs in the CUDA kernels (pcm_kernels.cu, pcm_tiled.cu, clique_kernels.cu); there
 * is no host fallback.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <atomic>
#include <climits>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/rpgo_b200.h"
#include "comm.h"
#include "keymap.h"
#include "kernels.cuh"

using namespace rpgo;

namespace {

/* Device memory comes from a per-handle bump arena whose chunks are taken from a stream-ordered memory pool that is
 * PRIVATE to this library (one per device, created on first use, never trimmed: freed chunks stay cached, so the next
 * handle of the process re-uses them).  A PCM session makes hundreds of small growing allocations (36 groups x 10
 * arrays) and plain cudaMalloc costs milliseconds each on this platform.  The application's default pool is not
 * touched. */
constexpr int MAX_DEVICES = 64;
cudaMemPool_t library_pool(int dev) {
  static std::mutex mu;
  static cudaMemPool_t pools[MAX_DEVICES] = {}; /* intentionally never destroyed (no CUDA calls at process exit) */
  if (dev < 0 || dev >= MAX_DEVICES) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    uint64_t thr = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    pools[dev] = pool;
  }
  return pools[dev];
}

struct Arena {
  struct Chunk { char* p; size_t cap, used; };
  std::vector<Chunk> chunks;
  cudaStream_t st = nullptr;
  cudaMemPool_t pool = nullptr;
  static constexpr size_t CHUNK = (size_t)64 << 20;
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    /* best fit over the chunks: after rewind() the same request sequence then lands in the same chunks, and a small
     * request never takes the room a dedicated large chunk was made for (first fit did, and the large request that
     * followed had to grow the pool: 100-300 ms outliers in repeated reset -> append -> select cycles) */
    int best = -1;
    for (size_t i = 0; i < chunks.size(); ++i) {
      const Chunk& c = chunks[i];
      if (c.used + bytes <= c.cap && (best < 0 || c.cap - c.used < chunks[best].cap - chunks[best].used)) best = (int)i;
    }
    if (best >= 0) {
      Chunk& c = chunks[best];
      void* r = c.p + c.used;
      c.used += bytes;
      return r;
    }
    const size_t cap = bytes > CHUNK ? bytes : CHUNK;
    void* p = nullptr;
    if (cudaMallocFromPoolAsync(&p, cap, pool, st) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    chunks.push_back(Chunk{(char*)p, cap, bytes});
    return p;
  }
  void release() {
    for (auto& c : chunks) cudaFreeAsync(c.p, st);
    chunks.clear();
  }
  /* rpgo_reset: every block handed out so far is dead; keep the chunks */
  void rewind() {
    for (auto& c : chunks) c.used = 0;
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  Arena* arena = nullptr;
  /* grow to at least `bytes`; optionally keep the first `keep` bytes; new memory is zeroed.  The old block
   * stays in the arena until the handle dies (growth is geometric, so at most 2x is held). */
  cudaError_t ensure(size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = std::max(bytes, cap * 2);
    ncap = (ncap + 255) & ~size_t(255);
    void* np = arena->alloc(ncap);
    if (!np) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemsetAsync(np, 0, ncap, st);
    if (e != cudaSuccess) return e;
    if (p && keep) {
      e = cudaMemcpyAsync(np, p, std::min(keep, cap), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return e;
    }
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

/* Pinned host staging.  cudaHostAlloc costs ~10 ms for the 44 MB a 50k-closure batch needs, so buffers are recycled
 * through a process-wide free list: a handle borrows one and hands it back when it is destroyed.  Nothing here runs a
 * CUDA call from a static or thread_local destructor (the list itself is leaked at exit on purpose). */
struct PinCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
};
PinCache& pin_cache() {
  static PinCache* c = new PinCache();
  return *c;
}
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  void give_back() {
    if (!p) return;
    PinCache& c = pin_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.free_list.push_back({p, cap});
    p = nullptr;
    cap = 0;
  }
  void* ensure(size_t bytes) {
    if (bytes <= cap) return p;
    give_back();
    PinCache& c = pin_cache();
    {
      std::lock_guard<std::mutex> lk(c.mu);
      int best = -1;
      for (int i = 0; i < (int)c.free_list.size(); ++i)
        if (c.free_list[i].second >= bytes && (best < 0 || c.free_list[i].second < c.free_list[best].second)) best = i;
      if (best >= 0) {
        p = c.free_list[best].first;
        cap = c.free_list[best].second;
        c.free_list.erase(c.free_list.begin() + best);
        return p;
      }
      /* nothing fits: drop the largest cached buffer so that growing sessions do not pile up stale ones */
      if (!c.free_list.empty()) {
        int big = 0;
        for (int i = 1; i < (int)c.free_list.size(); ++i)
          if (c.free_list[i].second > c.free_list[big].second) big = i;
        cudaFreeHost(c.free_list[big].first);
        c.free_list.erase(c.free_list.begin() + big);
      }
    }
    size_t want = (bytes + (bytes >> 2) + 4095) & ~size_t(4095);
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      cap = 0;
      return nullptr;
    }
    cap = want;
    return p;
  }
};

/* host -> pinned staging copies of tens of MB: one core moves ~10 GB/s, four move the 19 MB of a 50k batch in ~0.5 ms */
static void staged_copy(void* dst, const void* src, size_t bytes) {
  constexpr size_t MIN_PER_THREAD = (size_t)2 << 20;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t T = std::min<size_t>({(size_t)4, (size_t)hw, bytes / MIN_PER_THREAD});
  if (T <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t per = ((bytes + T - 1) / T + 63) & ~size_t(63);
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; ++t) {
    const size_t off = t * per;
    if (off >= bytes) break;
    const size_t len = std::min(per, bytes - off);
    try {
      th.emplace_back([=]() { memcpy((char*)dst + off, (const char*)src + off, len); });
    } catch (...) { /* no thread to be had (resource limits of the host application): copy this part inline */
      memcpy((char*)dst + off, (const char*)src + off, len);
    }
  }
  memcpy(dst, src, std::min(per, bytes));
  for (auto& x : th) x.join();
}

/* every entry point runs on the handle's device and leaves the caller's current device as it found it */
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
};

struct Group {
  uint8_t id1 = 0, id2 = 0;
  int64_t n = 0;         /* closures stored */
  int64_t cap = 0;       /* capacity in closures (multiple of 128) */
  int64_t stride32 = 0;  /* adjacency row stride in 32-bit words (= cap / 32) */
  DevBuf lc, idxf, idxb, pfx, bits, deg, fl_pairs, fl_count;
  DevBuf rec_aos, rec_soa, rec_col; /* gathered per-closure records for the tiled kernel: rows, SoA slabs, per-lane columns */
  bool landmark = false;   /* N3: observations of one landmark instead of closures of one prefix pair */
  uint64_t lkey = 0;
  DevBuf idxa0;
  std::vector<int32_t> h_idxa0;
  int64_t gathered = 0;    /* closures [0, gathered) have up-to-date records */
  std::vector<uint64_t> kfrom, kto;
  std::vector<int32_t> h_idxf, h_idxb;
  std::vector<uint8_t> h_pfx;
  explicit Group(Arena* a) {
    for (DevBuf* b : {&lc, &idxf, &idxb, &pfx, &bits, &deg, &fl_pairs, &fl_count, &rec_aos, &rec_soa, &rec_col, &idxa0}) b->arena = a;
  }
};

constexpr int64_t FLAG_CAP = 1 << 20;

}  // namespace

static constexpr int64_t CLIQUE_BLOCKS = 148 * 8;
struct CliqueSet {
  DevBuf degmask, picks, elim, result, ctl, rwork;
  void wire(Arena* a) {
    for (DevBuf* b : {&degmask, &picks, &elim, &result, &ctl, &rwork}) b->arena = a;
  }
};
struct CliqueWorker {
  CliqueSet set;
  cudaStream_t stream = nullptr;
  explicit CliqueWorker(Arena* a) { set.wire(a); }
};

struct rpgo_handle {
  Arena arena;   /* state: trajectory table and per-group buffers; rewound by rpgo_reset */
  Arena scratch; /* staging and clique scratch: no state between calls, so it survives rpgo_reset untouched */
  rpgo_cfg cfg;
  int device = 0;               /* CUDA ordinal this handle lives on (every entry point switches to it) */
  rpgo::Comm* comm = nullptr;   /* NCCL communicator over the world's GPUs (rpgo_comm_init), or null */
  PinBuf pin;                   /* pinned host staging, borrowed from the process-wide cache */
  PinBuf pin_odom[2];           /* staging of rpgo_odom_append (two, used alternately): its own buffer, so the call returns while its H2D copy and the
                                   fold are still running and the caller stages the loop closures into `pin` meanwhile */
  int dim = 3, mode = 0, E = 50, PS = 12, NN = 36;
  bool odom_check = true, loop_check = true;
  Thresholds th;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr}; /* end of the last H2D copy out of pin_odom[i] */
  int odom_turn = 0;
  std::string err;
  int64_t launches = 0;
  CliqueStats last_clique;         /* statistics of the last rpgo_find_inliers heuristic search */
  rpgo_exchange_fn xchg = nullptr; /* incumbent exchange of the sharded clique searches */
  void* xchg_user = nullptr;

  /* trajectory table: entry 0 is the default-constructed T (identity, zero covariance, node 0) */
  DevBuf traj;
  int64_t traj_n = 0;
  KeyMap key2idx;
  std::set<uint8_t> prefixes; /* prefixes for which odom_trajectories_ has an entry */
  std::set<uint64_t> missing_refs; /* keys looked up by a closure while absent (default entry used) */
  bool traj_dirty = false;         /* an entry a stored closure refers to has changed since it was resolved */

  std::vector<Group*> groups;
  std::map<std::pair<uint8_t, uint8_t>, int32_t> gindex;
  std::map<uint64_t, int32_t> lindex; /* landmark key -> group ordinal */

  DevBuf d_stage;
  /* uploads that overlap the trajectory fold: the factors of odometry slice i+1 and the closure batch travel on their own
   * stream into their own device staging buffers while the fold of slice i runs on `stream` */
  cudaStream_t copy_stream = nullptr;
  DevBuf d_stage_odom[2], d_stage_lc;
  cudaEvent_t ev_fold[2] = {nullptr, nullptr}; /* the fold that last read d_stage_odom[i] */
  cudaEvent_t ev_lc = nullptr, ev_grow = nullptr;
  DevBuf d_lcent, d_ok, d_dist, d_scan;
  /* clique scratch: one set for the single-group entry point, one per worker of the batched one */
  CliqueSet cset;
  std::vector<CliqueWorker*> workers;

  void wire() {
    arena.st = stream;
    scratch.st = stream;
    traj.arena = &arena;
    for (DevBuf* b : {&d_stage, &d_stage_odom[0], &d_stage_odom[1], &d_stage_lc, &d_lcent, &d_ok, &d_dist, &d_scan, &cset.degmask, &cset.picks, &cset.elim, &cset.result, &cset.ctl, &cset.rwork})
      b->arena = &scratch;
  }
  ~rpgo_handle() {
    DeviceGuard dg(device);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    if (stream) cudaStreamSynchronize(stream);
    if (comm) rpgo::comm_destroy(comm);
    pin.give_back();
    pin_odom[0].give_back();
    pin_odom[1].give_back();
    for (CliqueWorker* w : workers) {
      if (w->stream) {
        cudaStreamSynchronize(w->stream);
        cudaStreamDestroy(w->stream);
      }
      delete w;
    }
    for (Group* g : groups) delete g;
    if (stream) {
      arena.release();
      scratch.release();
      cudaStreamSynchronize(stream);
      cudaStreamDestroy(stream);
    }
    for (cudaEvent_t e : {ev_stage[0], ev_stage[1], ev_fold[0], ev_fold[1], ev_lc, ev_grow})
      if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
};

#define H_CHECK_CUDA(h, x)                                                           \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      (h)->err = std::string(#x) + ": " + cudaGetErrorString(e_);                    \
      return RPGO_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

static inline uint8_t key_chr(uint64_t k) { return (uint8_t)(k >> 56); }

static int traj_lookup(rpgo_handle* h, uint64_t key) {
  auto it = h->key2idx.find(key);
  return it == h->key2idx.end() ? 0 : it->second; /* missing key -> default entry (operator[] semantics) */
}

static int ensure_traj(rpgo_handle* h, int64_t entries) {
  H_CHECK_CUDA(h, h->traj.ensure((size_t)entries * h->E * sizeof(double), (size_t)h->traj_n * h->E * sizeof(double), h->stream));
  return RPGO_OK;
}

static GroupView group_view(rpgo_handle* h, Group* g) {
  GroupView v;
  v.lc = g->lc.as<double>();
  v.idx_front = g->idxf.as<int32_t>();
  v.idx_back = g->idxb.as<int32_t>();
  v.pfx_front = g->pfx.as<uint8_t>();
  v.bits = g->bits.as<uint32_t>();
  v.stride32 = g->stride32;
  v.deg = g->deg.as<int32_t>();
  v.n = (int32_t)g->n;
  v.idx_a0 = g->idxa0.as<int32_t>();
  return v;
}

static Shard group_shard(rpgo_handle* h, Group* g) {
  Shard s;
  s.rank = h->cfg.rank;
  s.world = h->cfg.world < 1 ? 1 : h->cfg.world;
  int64_t c = (g->n + 2 * s.world - 1) / (2 * s.world);
  c = (c + 31) / 32 * 32;
  s.chunk_rows = c < 32 ? 32 : c;
  return s;
}

/* grow a group's device arrays to hold `need` closures */
static int ensure_group(rpgo_handle* h, Group* g, int64_t need) {
  if (need <= g->cap) return RPGO_OK;
  int64_t ncap = std::max<int64_t>(need, g->cap * 2);
  ncap = (ncap + 127) / 128 * 128;
  cudaStream_t st = h->stream;
  H_CHECK_CUDA(h, g->lc.ensure((size_t)ncap * h->E * sizeof(double), (size_t)g->n * h->E * sizeof(double), st));
  H_CHECK_CUDA(h, g->idxf.ensure((size_t)ncap * 4, (size_t)g->n * 4, st));
  H_CHECK_CUDA(h, g->idxb.ensure((size_t)ncap * 4, (size_t)g->n * 4, st));
  H_CHECK_CUDA(h, g->pfx.ensure((size_t)ncap, (size_t)g->n, st));
  H_CHECK_CUDA(h, g->deg.ensure((size_t)ncap * 4, 0, st));
  if (g->landmark) H_CHECK_CUDA(h, g->idxa0.ensure((size_t)ncap * 4, (size_t)g->n * 4, st));
  if (h->loop_check && !g->landmark) {
    const size_t rn = (size_t)tiled_record_doubles(h->dim, h->mode) * 8;
    H_CHECK_CUDA(h, g->rec_aos.ensure((size_t)ncap * rn + 256, (size_t)g->n * rn, st));
    H_CHECK_CUDA(h, g->rec_soa.ensure((size_t)ncap * rn + 256, (size_t)((g->n + 31) / 32) * 32 * rn, st));
    const size_t cn = (size_t)tiled_column_pitch(h->dim, h->mode) * 8;
    H_CHECK_CUDA(h, g->rec_col.ensure((size_t)ncap * cn + 256, (size_t)g->n * cn, st));
  }
  if (h->loop_check || g->landmark) {
    /* adjacency: ncap rows x ncap/32 words; the row pitch changes, so copy row by row (2D copy) */
    const int64_t nstride = ncap / 32;
    /* in multi-GPU mode rows are padded to 2*world chunks */
    const int64_t rows = ncap + 64 * (int64_t)std::max(1, h->cfg.world);
    const size_t bytes = (size_t)rows * nstride * 4;
    void* nb = h->arena.alloc(bytes);
    if (!nb) { h->err = "device allocation failed (adjacency)"; return RPGO_ERR_NOMEM; }
    H_CHECK_CUDA(h, cudaMemsetAsync(nb, 0, bytes, st));
    if (g->bits.p && g->n > 0)
      H_CHECK_CUDA(h, cudaMemcpy2DAsync(nb, (size_t)nstride * 4, g->bits.p, (size_t)g->stride32 * 4,
                                        (size_t)((g->n + 31) / 32) * 4, (size_t)g->n, cudaMemcpyDeviceToDevice, st));
    g->bits.p = nb;
    g->bits.cap = bytes;
    g->stride32 = nstride;
    if (!g->fl_pairs.p) {
      H_CHECK_CUDA(h, g->fl_pairs.ensure((size_t)FLAG_CAP * 2 * 4, 0, st));
      H_CHECK_CUDA(h, g->fl_count.ensure(8, 0, st));
    }
  }
  g->cap = ncap;
  return RPGO_OK;
}

static Flagged group_flagged(Group* g) {
  Flagged f;
  f.pairs = g->fl_pairs.as<int32_t>();
  f.count = g->fl_count.as<unsigned long long>();
  f.cap = FLAG_CAP;
  return f;
}

static int run_pairwise(rpgo_handle* h, Group* g, int64_t j_begin, double* dist_dev) {
  if (!h->loop_check || g->n < 2) return RPGO_OK;
  GroupView v = group_view(h, g);
  if (g->landmark) {
    launch_landmark_direct(h->dim, h->mode, v, h->traj.as<double>(), (int)j_begin, h->th, group_flagged(g), dist_dev, h->stream);
    h->launches += 1;
    H_CHECK_CUDA(h, cudaGetLastError());
    return RPGO_OK;
  }
  Shard sh = group_shard(h, g);
  int kernel = h->cfg.kernel;
  int variant = 0;
  if (kernel == RPGO_KERNEL_TILED_ONE_GROUP) { variant = 1; kernel = RPGO_KERNEL_TILED; }
  else if (kernel == RPGO_KERNEL_TILED_V1) { variant = 2; kernel = RPGO_KERNEL_TILED; }
  if (dist_dev) kernel = RPGO_KERNEL_DIRECT;
  if (kernel == RPGO_KERNEL_AUTO) kernel = RPGO_KERNEL_TILED;
  if (variant != 0 && h->mode != MODE_PCM) variant = 0; /* the cross-check forms exist for the PCM chain only */
  if (kernel == RPGO_KERNEL_TILED) {
    if (g->gathered < g->n) {
      launch_gather_records(h->dim, h->mode, v, h->traj.as<double>(), (int)g->gathered, g->rec_aos.as<double>(),
                            g->rec_soa.as<double>(), g->rec_col.as<double>(), h->stream);
      g->gathered = g->n;
      h->launches += 1;
    }
    launch_pairwise_tiled(h->dim, h->mode, v, g->rec_aos.as<double>(), g->rec_soa.as<double>(), g->rec_col.as<double>(), (int)j_begin,
                          sh, h->th, group_flagged(g), variant, h->stream);
  } else
    launch_pairwise_direct(h->dim, h->mode, v, h->traj.as<double>(), (int)j_begin, sh, h->th, group_flagged(g), dist_dev,
                           h->stream);
  h->launches += 1;
  H_CHECK_CUDA(h, cudaGetLastError());
  return RPGO_OK;
}

static int finalize_group(rpgo_handle* h, Group* g, int64_t j_begin) {
  if ((!h->loop_check && !g->landmark) || g->n < 1) return RPGO_OK;
  launch_mirror(g->bits.as<uint32_t>(), g->stride32, (int)g->n, (int)j_begin, h->stream);
  launch_degree(g->bits.as<uint32_t>(), g->stride32, (int)g->n, g->deg.as<int32_t>(), h->stream);
  h->launches += 2;
  H_CHECK_CUDA(h, cudaGetLastError());
  return RPGO_OK;
}

int rpgo::CliqueShard::xchg(int32_t op, void* buf, int64_t count, int32_t root) const {
  if (comm) {
    switch (op) {
      case RPGO_XCHG_MIN_I64: return comm_allreduce_i64_host(comm, (long long*)buf, (size_t)count, false, comm_stream);
      case RPGO_XCHG_MAX_I64: return comm_allreduce_i64_host(comm, (long long*)buf, (size_t)count, true, comm_stream);
      case RPGO_XCHG_BCAST_I32: return comm_bcast_host(comm, buf, (size_t)count * 4, root, comm_stream);
      default: return 1;
    }
  }
  if (exchange) return exchange(user, op, buf, count, root);
  return 1;
}

static CliqueShard clique_shard(rpgo_handle* h, cudaStream_t st) {
  CliqueShard cs;
  cs.rank = h->cfg.rank;
  cs.world = h->cfg.world;
  cs.exchange = h->xchg;
  cs.user = h->xchg_user;
  cs.comm = h->comm;
  cs.comm_stream = st;
  return cs;
}

/* the one exchange step of the sharded pair matrix: all-gather of this group's adjacency row chunks over NCCL, in place,
 * on the handle's stream (every rank holds the same group geometry: the tables are replicated) */
static int allgather_group(rpgo_handle* h, Group* g) {
  if (!h->comm) { h->err = "no communicator: call rpgo_comm_init first"; return RPGO_ERR_INVALID; }
  if ((!h->loop_check && !g->landmark) || g->n < 1 || !g->bits.p) return RPGO_OK;
  const Shard s = group_shard(h, g);
  const size_t chunk_bytes = (size_t)s.chunk_rows * (size_t)g->stride32 * 4;
  if (comm_allgather_row_chunks(h->comm, g->bits.p, chunk_bytes, h->stream) != 0) {
    h->err = std::string("adjacency all-gather failed: ") + comm_error(h->comm);
    return RPGO_ERR_CUDA;
  }
  return RPGO_OK;
}

extern "C" {

const char* rpgo_version(void) { return "rpgo_b200 0.1 (sm_100a)"; }

int rpgo_default_cfg(rpgo_cfg* c) {
  if (!c) return RPGO_ERR_INVALID;
  memset(c, 0, sizeof(*c));
  c->dim = 3;
  c->mode = RPGO_MODE_PCM;
  /* PcmParams defaults, SolverParams.h:35-42 */
  c->odom_threshold = 10.0;
  c->lc_threshold = 5.0;
  c->odom_trans_threshold = 0.05;
  c->odom_rot_threshold = 0.005;
  c->dist_trans_threshold = 0.01;
  c->dist_rot_threshold = 0.001;
  c->incremental = 0;
  c->device = -1;
  c->traj_mode = RPGO_TRAJ_FOLD;
  c->kernel = RPGO_KERNEL_AUTO;
  c->rank = 0;
  c->world = 1;
  c->band = 1e-9;
  c->scan_chunk = 64;
  return RPGO_OK;
}

int rpgo_create(const rpgo_cfg* cfg, rpgo_handle** out) {
  if (!cfg || !out) return RPGO_ERR_INVALID;
  if ((cfg->dim != 2 && cfg->dim != 3) || (cfg->mode != 0 && cfg->mode != 1)) return RPGO_ERR_INVALID;
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return RPGO_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return RPGO_ERR_CUDA; } /* no CPU fallback */
  int dev = cfg->device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return RPGO_ERR_CUDA;
  if (dev >= ndev || dev >= MAX_DEVICES) return RPGO_ERR_INVALID;
  {
    /* the library carries sm_100a code only: refuse anything else instead of failing at the first launch */
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
      return RPGO_ERR_CUDA;
    if (major != 10 || minor != 0) return RPGO_ERR_CUDA;
  }
  DeviceGuard dg(dev);
  rpgo_handle* h = new rpgo_handle();
  h->device = dev;
  h->cfg = *cfg;
  h->cfg.device = dev;
  if (h->cfg.band <= 0) h->cfg.band = 1e-9;
  if (h->cfg.scan_chunk <= 0) h->cfg.scan_chunk = 64;
  h->dim = cfg->dim;
  h->mode = cfg->mode;
  h->E = cfg->dim == 3 ? Dim<3>::ENTRY : Dim<2>::ENTRY;
  h->PS = cfg->dim == 3 ? 12 : 4;
  h->NN = cfg->dim == 3 ? 36 : 9;
  /* Pcm.h:74-82 */
  h->odom_check = !(cfg->odom_threshold < 0 || cfg->odom_rot_threshold < 0 || cfg->odom_trans_threshold < 0);
  h->loop_check = !(cfg->lc_threshold < 0 || cfg->dist_rot_threshold < 0 || cfg->dist_trans_threshold < 0);
  h->th.odom = cfg->odom_threshold;
  h->th.lc = cfg->lc_threshold;
  h->th.odom_trans = cfg->odom_trans_threshold;
  h->th.odom_rot = cfg->odom_rot_threshold;
  h->th.dist_trans = cfg->dist_trans_threshold;
  h->th.dist_rot = cfg->dist_rot_threshold;
  h->th.band = h->cfg.band;
  h->arena.pool = h->scratch.pool = library_pool(dev);
  if (!h->arena.pool) {
    delete h;
    return RPGO_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    h->stream = nullptr;
    delete h;
    return RPGO_ERR_CUDA;
  }
  h->wire();
  if (cudaEventCreateWithFlags(&h->ev_stage[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_stage[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fold[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fold[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_lc, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_grow, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return RPGO_ERR_CUDA;
  }
  /* entry 0 = default T: identity pose, zero covariance, node 0, rotation_info true */
  if (h->traj.ensure((size_t)1024 * h->E * sizeof(double), 0, h->stream) != cudaSuccess) {
    delete h;
    return RPGO_ERR_CUDA;
  }
  std::vector<double> e0(h->E, 0.0);
  if (h->dim == 3) { e0[0] = e0[4] = e0[8] = 1.0; e0[Dim<3>::OFF_ROT] = 1.0; }
  else { e0[0] = 1.0; e0[Dim<2>::OFF_ROT] = 1.0; }
  cudaMemcpyAsync(h->traj.p, e0.data(), sizeof(double) * h->E, cudaMemcpyHostToDevice, h->stream);
  cudaStreamSynchronize(h->stream);
  h->traj_n = 1;
  *out = h;
  return RPGO_OK;
}

void rpgo_destroy(rpgo_handle* h) {
  if (!h) return;
  delete h; /* ~rpgo_handle switches to the handle's device, drains the stream and tears the communicator down */
}

/* Back to the state right after rpgo_create (a freshly constructed Pcm object, Pcm.h:64-96) while keeping what is
 * expensive to build: the stream(s), the device arena's memory, the pinned staging buffer and the communicator. */
int rpgo_reset(rpgo_handle* h) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  for (CliqueWorker* w : h->workers) H_CHECK_CUDA(h, cudaStreamSynchronize(w->stream));
  for (Group* g : h->groups) delete g;
  h->groups.clear();
  h->gindex.clear();
  h->lindex.clear();
  h->key2idx.clear();
  h->prefixes.clear();
  h->missing_refs.clear();
  h->traj_dirty = false;
  h->traj.p = nullptr; /* staging buffers and clique scratch sets live in h->scratch and are kept */
  h->traj.cap = 0;
  h->arena.rewind();
  H_CHECK_CUDA(h, h->traj.ensure((size_t)1024 * h->E * sizeof(double), 0, h->stream));
  std::vector<double> e0(h->E, 0.0);
  if (h->dim == 3) { e0[0] = e0[4] = e0[8] = 1.0; e0[Dim<3>::OFF_ROT] = 1.0; }
  else { e0[0] = 1.0; e0[Dim<2>::OFF_ROT] = 1.0; }
  H_CHECK_CUDA(h, cudaMemcpyAsync(h->traj.p, e0.data(), sizeof(double) * h->E, cudaMemcpyHostToDevice, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  h->traj_n = 1;
  h->err.clear();
  return RPGO_OK;
}

const char* rpgo_last_error(const rpgo_handle* h) { return h ? h->err.c_str() : "null handle"; }

int rpgo_sync(rpgo_handle* h) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  return RPGO_OK;
}

void* rpgo_stream(rpgo_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t rpgo_launch_count(rpgo_handle* h) { return h ? h->launches : 0; }
int64_t rpgo_traj_size(rpgo_handle* h) { return h ? (int64_t)h->key2idx.size() : 0; }

/* ---------------------------------------------------------------------------------------------- */
static int odom_append_batch(rpgo_handle* h, int64_t n, const uint64_t* prev_key, const uint64_t* new_key,
                             const double* delta_pose, const double* delta_cov, const double* init_pose);

int rpgo_odom_append(rpgo_handle* h, int64_t n, const uint64_t* prev_key, const uint64_t* new_key,
                     const double* delta_pose, const double* delta_cov, const double* init_pose) {
  if (!h || n < 0) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (n == 0) return RPGO_OK;
  if (!prev_key || !new_key || !delta_pose || !delta_cov) return RPGO_ERR_INVALID;
  for (int64_t k = 0; k < n; ++k) /* the all-ones key is the key table's empty marker (not a valid gtsam::Symbol) */
    if (prev_key[k] == ~0ull || new_key[k] == ~0ull) { h->err = "rpgo_odom_append: key 0xffffffffffffffff is not supported"; return RPGO_ERR_INVALID; }
  /* A long batch is fed to the exact fold in slices (exactly what a caller appending odometry in several calls does):
   * the fold of slice i runs on the GPU while the host resolves the keys of slice i+1 and stages its factors, so the
   * wall time is the fold's own plus one slice of host work instead of the sum of both. */
  constexpr int64_t SLICE = 8192;
  if (h->cfg.traj_mode != RPGO_TRAJ_FOLD || n <= SLICE + SLICE / 2)
    return odom_append_batch(h, n, prev_key, new_key, delta_pose, delta_cov, init_pose);
  for (int64_t k0 = 0; k0 < n; k0 += SLICE) {
    const int64_t m = std::min(SLICE, n - k0);
    const int rc = odom_append_batch(h, m, prev_key + k0, new_key + k0, delta_pose + (size_t)k0 * h->PS,
                                     delta_cov + (size_t)k0 * h->NN, init_pose ? init_pose + (size_t)k0 * h->PS : nullptr);
    if (rc != RPGO_OK) return rc;
  }
  return RPGO_OK;
}

static int odom_append_batch(rpgo_handle* h, int64_t n, const uint64_t* prev_key, const uint64_t* new_key,
                             const double* delta_pose, const double* delta_cov, const double* init_pose) {
  const int E = h->E, PS = h->PS, NN = h->NN;
  cudaStream_t st = h->stream;

  static const bool trace = getenv("RPGO_TRACE") != nullptr;
  auto now_ms = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  double t_mark = trace ? now_ms() : 0.0;
  auto mark = [&](const char* what) {
    if (!trace) return;
    const double t = now_ms();
    fprintf(stderr, "[odom_append n=%lld] %-28s %8.3f ms\n", (long long)n, what, t - t_mark);
    t_mark = t;
  };

  /* host pass: resolve indices, seed new prefixes, build chains (Pcm.h:524-556) */
  struct Step { int32_t src; int32_t out; int64_t k; };
  std::vector<std::vector<Step>> chains;       /* steps per chain */
  std::vector<int32_t> chain_start;            /* start entry per chain */
  int open_chain[256];                         /* prefix -> chain index whose tail can be extended, or -1 */
  std::fill(open_chain, open_chain + 256, -1);
  std::vector<uint64_t> chain_tail_key;        /* per chain: the key its last step produced */
  bool seen_prefix[256] = {false};             /* flat mirror of h->prefixes for the per-step test */
  for (uint8_t c : h->prefixes) seen_prefix[c] = true;
  std::set<int32_t> overwritten;               /* pre-existing entries re-written by this batch (rare) */
  std::set<int32_t> seed_idx;                  /* host-written seed entries: ready before any kernel runs */
  const int64_t first_new_entry = h->traj_n;   /* entries >= this index are produced by this batch */
  h->key2idx.reserve(h->key2idx.size() + (size_t)n + 16);
  std::vector<std::pair<int32_t, int64_t>> seeds; /* (entry, step k) */
  int64_t new_entries = h->traj_n;
  bool need_flush_order = false;

  for (int64_t k = 0; k < n; ++k) {
    const uint8_t prefix = key_chr(new_key[k]);
    if (!seen_prefix[prefix]) {
      /* new prefix: poses[prev_key] = (values.at(prev_key), zero cov)  Pcm.h:534-542 */
      seen_prefix[prefix] = true;
      h->prefixes.insert(prefix);
      int32_t idx;
      auto it = h->key2idx.find(prev_key[k]);
      if (it == h->key2idx.end()) {
        idx = (int32_t)new_entries++;
        h->key2idx[prev_key[k]] = idx;
        if (h->missing_refs.erase(prev_key[k])) h->traj_dirty = true;
      } else {
        idx = it->second;
        h->traj_dirty = true;
      }
      seeds.push_back({idx, k});
      seed_idx.insert(idx);
    }
    const int32_t src = traj_lookup(h, prev_key[k]);
    int32_t out;
    auto it = h->key2idx.find(new_key[k]);
    if (it == h->key2idx.end()) {
      out = (int32_t)new_entries++;
      h->key2idx[new_key[k]] = out;
      if (h->missing_refs.erase(new_key[k])) h->traj_dirty = true; /* a closure resolved this key to the default entry */
    } else {
      out = it->second;
      h->traj_dirty = true; /* poses[new_key] overwritten (Pcm.h:556) */
    }
    const int oc = open_chain[prefix];
    if (oc >= 0 && chain_tail_key[oc] == prev_key[k]) {
      chains[oc].push_back({src, out, k});
      chain_tail_key[oc] = new_key[k];
    } else {
      if ((src >= first_new_entry && !seed_idx.count(src)) || overwritten.count(src)) need_flush_order = true; /* starts from an entry another chain writes */
      chains.push_back({{src, out, k}});
      chain_start.push_back(src);
      open_chain[prefix] = (int)chains.size() - 1;
      chain_tail_key.push_back(new_key[k]);
    }
    if (out < first_new_entry) overwritten.insert(out);
  }
  mark("keys + chains (host)");
  int rc = ensure_traj(h, new_entries);
  if (rc != RPGO_OK) return rc;
  mark("ensure_traj");

  /* seeds: host-built entries, uploaded before any kernel */
  if (!seeds.empty()) {
    std::vector<double> e((size_t)E, 0.0);
    for (auto& s : seeds) {
      std::fill(e.begin(), e.end(), 0.0);
      if (init_pose) {
        memcpy(e.data(), init_pose + (size_t)s.second * PS, sizeof(double) * PS);
      } else {
        if (h->dim == 3) { e[0] = e[4] = e[8] = 1.0; } else { e[0] = 1.0; }
      }
      e[h->dim == 3 ? Dim<3>::OFF_ROT : Dim<2>::OFF_ROT] = 1.0;
      H_CHECK_CUDA(h, cudaMemcpyAsync(h->traj.as<double>() + (size_t)s.first * E, e.data(), sizeof(double) * E,
                                      cudaMemcpyHostToDevice, st));
      H_CHECK_CUDA(h, cudaStreamSynchronize(st)); /* e is reused */
    }
  }

  /* flatten chains; steps of one chain are contiguous */
  const size_t bytes_pose = (size_t)n * PS * 8, bytes_cov = (size_t)n * NN * 8;
  const size_t off_cov = bytes_pose, off_out = off_cov + bytes_cov, off_chain = off_out + (size_t)n * 4;
  const size_t total = off_chain + chains.size() * sizeof(FoldChain) + 64;
  const int turn = (h->odom_turn ^= 1);
  H_CHECK_CUDA(h, cudaEventSynchronize(h->ev_stage[turn])); /* the H2D copy that last read this buffer */
  char* pin = (char*)h->pin_odom[turn].ensure(total);
  if (!pin) { h->err = "pinned staging allocation failed"; return RPGO_ERR_NOMEM; }
  double* p_pose = (double*)pin;
  double* p_cov = (double*)(pin + off_cov);
  int32_t* p_out = (int32_t*)(pin + off_out);
  FoldChain* p_chain = (FoldChain*)(pin + ((off_chain + 15) & ~size_t(15)));
  int64_t pos = 0;
  for (size_t c = 0; c < chains.size(); ++c) {
    p_chain[c].first_step = (int32_t)pos;
    p_chain[c].n_steps = (int32_t)chains[c].size();
    p_chain[c].start_idx = chain_start[c];
    p_chain[c].pad = 0;
    const std::vector<Step>& cs = chains[c];
    for (size_t a = 0; a < cs.size();) {
      /* steps that are consecutive in the input (the usual case: one robot's odometry in order) move as one block */
      size_t b = a + 1;
      while (b < cs.size() && cs[b].k == cs[b - 1].k + 1) ++b;
      staged_copy(p_pose + (size_t)pos * PS, delta_pose + (size_t)cs[a].k * PS, sizeof(double) * PS * (b - a));
      staged_copy(p_cov + (size_t)pos * NN, delta_cov + (size_t)cs[a].k * NN, sizeof(double) * NN * (b - a));
      for (size_t q = a; q < b; ++q) p_out[pos++] = cs[q].out;
      a = b;
    }
  }
  mark("staging copy (host)");
  /* exact fold: the upload goes through the copy stream into one of two device staging buffers, so that it overlaps the
   * fold of the previous slice; `st` only waits for this slice's own upload.  (The scan path synchronises anyway.) */
  const bool overlap = need_flush_order || h->cfg.traj_mode == RPGO_TRAJ_FOLD;
  DevBuf& dstage = overlap ? h->d_stage_odom[turn] : h->d_stage;
  cudaStream_t cs = overlap ? h->copy_stream : st;
  {
    const void* before = dstage.p;
    H_CHECK_CUDA(h, dstage.ensure(total, 0, st));
    if (overlap) {
      if (dstage.p != before) { /* fresh memory is zero-filled on `st` */
        H_CHECK_CUDA(h, cudaEventRecord(h->ev_grow, st));
        H_CHECK_CUDA(h, cudaStreamWaitEvent(cs, h->ev_grow, 0));
      }
      H_CHECK_CUDA(h, cudaStreamWaitEvent(cs, h->ev_fold[turn], 0)); /* the fold that last read this buffer */
    }
  }
  H_CHECK_CUDA(h, cudaMemcpyAsync(dstage.p, pin, total, cudaMemcpyHostToDevice, cs));
  H_CHECK_CUDA(h, cudaEventRecord(h->ev_stage[turn], cs));
  if (overlap) H_CHECK_CUDA(h, cudaStreamWaitEvent(st, h->ev_stage[turn], 0));
  mark("H2D enqueue");
  char* d = (char*)dstage.p;
  const double* d_pose = (const double*)d;
  const double* d_cov = (const double*)(d + off_cov);
  const int32_t* d_out = (const int32_t*)(d + off_out);
  const FoldChain* d_chain = (const FoldChain*)(d + ((off_chain + 15) & ~size_t(15)));

  const int nch = (int)chains.size();
  if (need_flush_order || h->cfg.traj_mode == RPGO_TRAJ_FOLD) {
    if (need_flush_order) {
      /* rare: a chain starts from an entry written by an earlier chain of this batch -> serialise */
      for (int c = 0; c < nch; ++c) {
        launch_traj_fold(h->dim, h->mode, 1, d_chain + c, d_out, d_pose, d_cov, h->traj.as<double>(), st);
        h->launches++;
      }
    } else {
      launch_traj_fold(h->dim, h->mode, nch, d_chain, d_out, d_pose, d_cov, h->traj.as<double>(), st);
      h->launches++;
    }
  } else {
    /* chunked scan */
    const int chunk = h->cfg.scan_chunk;
    std::vector<int32_t> chunk_chain, chunk_first, chain_first_chunk(nch), step_chunk((size_t)n);
    for (int c = 0; c < nch; ++c) {
      chain_first_chunk[c] = (int32_t)chunk_chain.size();
      for (int s0 = 0; s0 < p_chain[c].n_steps; s0 += chunk) {
        const int q = (int)chunk_chain.size();
        chunk_chain.push_back(c);
        chunk_first.push_back(p_chain[c].first_step + s0);
        for (int s = s0; s < std::min(s0 + chunk, p_chain[c].n_steps); ++s) step_chunk[p_chain[c].first_step + s] = q;
      }
    }
    const int total_chunks = (int)chunk_chain.size();
    const size_t ib = ((size_t)total_chunks * 2 + nch + n) * 4;
    const size_t ib_al = (ib + 15) & ~size_t(15);
    DevBuf& scan_buf = h->d_scan;
    H_CHECK_CUDA(h, scan_buf.ensure(ib_al + (size_t)total_chunks * E * 8, 0, st));
    std::vector<int32_t> ints;
    ints.insert(ints.end(), chunk_chain.begin(), chunk_chain.end());
    ints.insert(ints.end(), chunk_first.begin(), chunk_first.end());
    ints.insert(ints.end(), chain_first_chunk.begin(), chain_first_chunk.end());
    ints.insert(ints.end(), step_chunk.begin(), step_chunk.end());
    H_CHECK_CUDA(h, cudaMemcpyAsync(scan_buf.p, ints.data(), ib, cudaMemcpyHostToDevice, st));
    H_CHECK_CUDA(h, cudaStreamSynchronize(st));
    const int32_t* di = scan_buf.as<int32_t>();
    launch_traj_scan_phases(h->dim, h->mode, nch, d_chain, d_out, d_pose, d_cov, h->traj.as<double>(), chunk, (int)n,
                            total_chunks, di, di + total_chunks, di + 2 * total_chunks, di + 2 * total_chunks + nch,
                            (double*)((char*)scan_buf.p + ib_al), st);
    h->launches += 3;
  }
  H_CHECK_CUDA(h, cudaGetLastError());
  if (overlap) H_CHECK_CUDA(h, cudaEventRecord(h->ev_fold[turn], st));
  h->traj_n = new_entries;
  /* no wait here: the H2D copy out of pin_odom and the fold keep running while the caller prepares the loop closures
   * (everything downstream is ordered on the same stream; the next rpgo_odom_append waits for ev_stage before it
   * overwrites the buffer) */
  return RPGO_OK;
}

int rpgo_traj_get(rpgo_handle* h, uint64_t key, double* pose, double* cov, int32_t* node, int32_t* rot_info) {
  if (!h || !pose) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  auto it = h->key2idx.find(key);
  if (it == h->key2idx.end()) return RPGO_ERR_NOT_FOUND;
  std::vector<double> e(h->E);
  H_CHECK_CUDA(h, cudaMemcpyAsync(e.data(), h->traj.as<double>() + (size_t)it->second * h->E, sizeof(double) * h->E,
                                  cudaMemcpyDeviceToHost, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  memcpy(pose, e.data(), sizeof(double) * h->PS);
  const int oc = h->dim == 3 ? Dim<3>::OFF_COV : Dim<2>::OFF_COV;
  const int orr = h->dim == 3 ? Dim<3>::OFF_ROT : Dim<2>::OFF_ROT;
  const int on = h->dim == 3 ? Dim<3>::OFF_NODE : Dim<2>::OFF_NODE;
  if (cov) memcpy(cov, e.data() + oc, sizeof(double) * h->NN);
  if (node) *node = (int32_t)e[on];
  if (rot_info) *rot_info = e[orr] != 0.0;
  return RPGO_OK;
}

/* ---------------------------------------------------------------------------------------------- */
int rpgo_lc_append(rpgo_handle* h, int64_t n, const uint64_t* key_from, const uint64_t* key_to,
                   const double* pose, const double* cov, uint8_t* accepted, int32_t* group, int32_t* index,
                   double* odom_dist) {
  if (!h || n < 0) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (n == 0) return RPGO_OK;
  if (!key_from || !key_to || !pose || !cov) return RPGO_ERR_INVALID;
  const int E = h->E, PS = h->PS, NN = h->NN;
  cudaStream_t st = h->stream;
  static const bool trace = getenv("RPGO_TRACE") != nullptr;
  auto now_ms = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  double t_mark = trace ? now_ms() : 0.0;
  auto mark = [&](const char* what) {
    if (!trace) return;
    const double t = now_ms();
    fprintf(stderr, "[lc_append n=%lld] %-28s %8.3f ms\n", (long long)n, what, t - t_mark);
    t_mark = t;
  };

  if (h->traj_dirty) {
    /* the reference looks trajectory entries up at every pair check (GraphUtils.h:40-42), so closures
     * stored earlier must see entries that appeared / changed since: re-resolve and re-gather */
    for (Group* g : h->groups) {
      for (int64_t k = 0; k < g->n; ++k) {
        g->h_idxf[k] = traj_lookup(h, g->kfrom[k]);
        g->h_idxb[k] = traj_lookup(h, g->kto[k]);
        if (g->landmark) g->h_idxa0[k] = traj_lookup(h, (uint64_t)key_chr(g->kfrom[k]) << 56);
      }
      if (g->landmark && g->n > 0 && g->idxa0.p)
        H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxa0.p, g->h_idxa0.data(), (size_t)g->n * 4, cudaMemcpyHostToDevice, st));
      if (g->n > 0 && g->idxf.p) {
        H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxf.p, g->h_idxf.data(), (size_t)g->n * 4, cudaMemcpyHostToDevice, st));
        H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxb.p, g->h_idxb.data(), (size_t)g->n * 4, cudaMemcpyHostToDevice, st));
      }
      g->gathered = 0;
    }
    H_CHECK_CUDA(h, cudaStreamSynchronize(st));
    h->traj_dirty = false;
  }
  /* stage inputs: pose | cov | idxf | idxb | check */
  const size_t o_cov = (size_t)n * PS * 8, o_if = o_cov + (size_t)n * NN * 8, o_ib = o_if + (size_t)n * 4,
               o_ck = o_ib + (size_t)n * 4, o_dst = (o_ck + (size_t)n + 15) & ~size_t(15),
               total = o_dst + (size_t)n * 8;
  char* pin = (char*)h->pin.ensure(total + (size_t)n * 9 + 64);
  if (!pin) { h->err = "pinned staging allocation failed"; return RPGO_ERR_NOMEM; }
  staged_copy(pin, pose, (size_t)n * PS * 8);
  staged_copy(pin + o_cov, cov, (size_t)n * NN * 8);
  int32_t* p_if = (int32_t*)(pin + o_if);
  int32_t* p_ib = (int32_t*)(pin + o_ib);
  uint8_t* p_ck = (uint8_t*)(pin + o_ck);
  uint64_t* p_dst = (uint64_t*)(pin + o_dst);
  bool any_check = false;
  for (int64_t k = 0; k < n; ++k) {
    const uint8_t cf = key_chr(key_from[k]), cb = key_chr(key_to[k]);
    const bool intra = cf == cb;
    p_ck[k] = (intra && h->odom_check) ? 1 : 0;
    any_check = any_check || p_ck[k];
    if (p_ck[k]) h->prefixes.insert(cf); /* odom_trajectories_[chr] is created by the lookup, Pcm.h:619 */
    p_if[k] = traj_lookup(h, key_from[k]);
    p_ib[k] = traj_lookup(h, key_to[k]);
    if (p_if[k] == 0 && !h->key2idx.count(key_from[k])) h->missing_refs.insert(key_from[k]);
    if (p_ib[k] == 0 && !h->key2idx.count(key_to[k])) h->missing_refs.insert(key_to[k]);
  }
  mark("stage + key lookups (host)");
  {
    const void* before = h->d_stage_lc.p;
    H_CHECK_CUDA(h, h->d_stage_lc.ensure(total, 0, st));
    if (h->d_stage_lc.p != before) {
      H_CHECK_CUDA(h, cudaEventRecord(h->ev_grow, st));
      H_CHECK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_grow, 0));
    }
  }
  H_CHECK_CUDA(h, h->d_lcent.ensure((size_t)n * E * 8, 0, st));
  H_CHECK_CUDA(h, h->d_ok.ensure((size_t)n, 0, st));
  H_CHECK_CUDA(h, h->d_dist.ensure((size_t)n * 8, 0, st));
  /* the closure batch is uploaded on the copy stream: it overlaps a trajectory fold still running on `st` (the previous
   * user of d_stage_lc finished before the last rpgo_lc_append returned: it ends with a stream synchronisation) */
  H_CHECK_CUDA(h, cudaMemcpyAsync(h->d_stage_lc.p, pin, o_dst, cudaMemcpyHostToDevice, h->copy_stream));
  H_CHECK_CUDA(h, cudaEventRecord(h->ev_lc, h->copy_stream));
  H_CHECK_CUDA(h, cudaStreamWaitEvent(st, h->ev_lc, 0));
  char* d = (char*)h->d_stage_lc.p;
  launch_lc_prepare(h->dim, h->mode, (int)n, (const double*)d, (const double*)(d + o_cov), (const int32_t*)(d + o_if),
                    (const int32_t*)(d + o_ib), (const uint8_t*)(d + o_ck), h->traj.as<double>(), h->th,
                    h->d_lcent.as<double>(), h->d_ok.as<uint8_t>(), h->d_dist.as<double>(), st);
  h->launches++;
  H_CHECK_CUDA(h, cudaGetLastError());
  uint8_t* h_ok = (uint8_t*)(pin + total);
  std::vector<double> dist_host;
  if (any_check) {
    H_CHECK_CUDA(h, cudaMemcpyAsync(h_ok, h->d_ok.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (odom_dist) {
      dist_host.resize((size_t)n);
      H_CHECK_CUDA(h, cudaMemcpyAsync(dist_host.data(), h->d_dist.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    H_CHECK_CUDA(h, cudaStreamSynchronize(st));
  } else {
    /* no closure of this batch is odometry-checked (check disabled, or inter-robot closures only): K2 accepts them all
     * and reports NaN distances, so the grouping below does not wait for the trajectory fold and K2 -- it runs on the
     * host while they are still executing, and K3 is enqueued behind them */
    memset(h_ok, 1, (size_t)n);
    if (odom_dist) dist_host.assign((size_t)n, std::nan(""));
  }
  mark("H2D + K2 + D2H (sync)");

  /* host pass 2: grouping in arrival order (Pcm.h:466-486) */
  std::map<int32_t, int64_t> old_n; /* groups touched -> size before this call */
  const size_t groups_at_entry = h->groups.size();
  std::vector<int64_t> need(h->groups.size(), 0);
  for (int64_t k = 0; k < n; ++k) {
    if (odom_dist) odom_dist[k] = dist_host[k];
    if (!h_ok[k]) {
      if (accepted) accepted[k] = 0;
      if (group) group[k] = -1;
      if (index) index[k] = -1;
      p_dst[k] = 0;
      continue;
    }
    const uint8_t cf = key_chr(key_from[k]), cb = key_chr(key_to[k]);
    const std::pair<uint8_t, uint8_t> id(std::min(cf, cb), std::max(cf, cb));
    int32_t gi;
    auto it = h->gindex.find(id);
    if (it == h->gindex.end()) {
      Group* g = new Group(&h->arena);
      g->id1 = id.first;
      g->id2 = id.second;
      gi = (int32_t)h->groups.size();
      h->groups.push_back(g);
      h->gindex[id] = gi;
    } else {
      gi = it->second;
    }
    Group* g = h->groups[gi];
    if (!old_n.count(gi)) old_n[gi] = g->n;
    const int64_t idx = g->n++;
    g->kfrom.push_back(key_from[k]);
    g->kto.push_back(key_to[k]);
    g->h_idxf.push_back(p_if[k]);
    g->h_idxb.push_back(p_ib[k]);
    g->h_pfx.push_back(cf);
    if (h->loop_check && idx >= 1) { /* areLoopsConsistent touches both trajectories, Pcm.h:703-710 */
      h->prefixes.insert(id.first);
      h->prefixes.insert(id.second);
    }
    if (accepted) accepted[k] = 1;
    if (group) group[k] = gi;
    if (index) index[k] = (int32_t)idx;
    p_dst[k] = (uint64_t)idx; /* patched to an address below, once capacities are final */
    /* remember group in the low bits via a side vector */
    need.resize(h->groups.size(), 0);
  }
  mark("grouping (host)");
  /* From here on the host bookkeeping (group sizes, key vectors, possibly new groups) is ahead of the device state.  If
   * an allocation, copy or launch fails the bookkeeping is rolled back, so the handle stays usable and consistent with the
   * caller's own lists (the reference has no failure mode here). */
  auto rollback = [&](size_t first_new_group) {
    for (auto& kv : old_n) {
      if ((size_t)kv.first >= first_new_group) continue;
      Group* g = h->groups[kv.first];
      g->n = kv.second;
      g->kfrom.resize((size_t)kv.second);
      g->kto.resize((size_t)kv.second);
      g->h_idxf.resize((size_t)kv.second);
      g->h_idxb.resize((size_t)kv.second);
      g->h_pfx.resize((size_t)kv.second);
      if (g->gathered > g->n) g->gathered = g->n;
    }
    while (h->groups.size() > first_new_group) {
      Group* g = h->groups.back();
      h->gindex.erase({g->id1, g->id2});
      delete g;
      h->groups.pop_back();
    }
  };
  const size_t first_new_group = groups_at_entry;
  auto device_part = [&]() -> int {
  /* capacities (n was advanced above; ensure_group must see the old n for its copies) */
  for (auto& kv : old_n) {
    Group* g = h->groups[kv.first];
    const int64_t n_new = g->n;
    g->n = kv.second;
    int rc = ensure_group(h, g, n_new);
    g->n = n_new;
    if (rc != RPGO_OK) return rc;
  }
  mark("ensure_group (alloc/copies)");
  /* destination addresses */
  {
    std::map<int32_t, int64_t> cursor = old_n;
    for (int64_t k = 0; k < n; ++k) {
      if (!h_ok[k]) continue;
      const uint8_t cf = key_chr(key_from[k]), cb = key_chr(key_to[k]);
      const int32_t gi = h->gindex[{std::min(cf, cb), std::max(cf, cb)}];
      Group* g = h->groups[gi];
      p_dst[k] = (uint64_t)(uintptr_t)(g->lc.as<double>() + (size_t)(cursor[gi]++) * E);
    }
  }
  H_CHECK_CUDA(h, cudaMemcpyAsync(d + o_dst, p_dst, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  launch_scatter_entries(h->dim, (int)n, h->d_lcent.as<double>(), (const uint64_t*)(d + o_dst), st);
  h->launches++;
  for (auto& kv : old_n) {
    Group* g = h->groups[kv.first];
    const int64_t o = kv.second, m = g->n - o;
    H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxf.as<int32_t>() + o, g->h_idxf.data() + o, (size_t)m * 4, cudaMemcpyHostToDevice, st));
    H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxb.as<int32_t>() + o, g->h_idxb.data() + o, (size_t)m * 4, cudaMemcpyHostToDevice, st));
    H_CHECK_CUDA(h, cudaMemcpyAsync(g->pfx.as<uint8_t>() + o, g->h_pfx.data() + o, (size_t)m, cudaMemcpyHostToDevice, st));
  }
  mark("scatter + index uploads");
  /* K3 per touched group */
  for (auto& kv : old_n) {
    Group* g = h->groups[kv.first];
    int rc = run_pairwise(h, g, kv.second, nullptr);
    if (rc != RPGO_OK) return rc;
    if (h->cfg.world <= 1) {
      rc = finalize_group(h, g, kv.second);
      if (rc != RPGO_OK) return rc;
    } else if (h->comm) {
      /* sharded rows: every rank computed its chunks; exchange them and rebuild mirror + degrees everywhere */
      rc = allgather_group(h, g);
      if (rc == RPGO_OK) rc = finalize_group(h, g, 0);
      if (rc != RPGO_OK) return rc;
    } /* else: the caller runs its own all-gather (rpgo_adj_bits_device) and then rpgo_group_finalize */
  }
  H_CHECK_CUDA(h, cudaGetLastError());
  mark("K3 + finalize launches");
  H_CHECK_CUDA(h, cudaStreamSynchronize(st)); /* host vectors / pinned staging are reused */
  mark("final sync");
  return RPGO_OK;
  };
  const int rc_dev = device_part();
  if (rc_dev != RPGO_OK) {
    cudaStreamSynchronize(st);
    cudaGetLastError();
    rollback(first_new_group);
  }
  return rc_dev;
}

int rpgo_landmark_append(rpgo_handle* h, uint64_t landmark_key, int64_t n, const uint64_t* pose_key, const double* pose,
                         const double* cov, int32_t reset, int32_t* group_out) {
  if (!h || n < 0) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (n > 0 && (!pose_key || !pose || !cov)) return RPGO_ERR_INVALID;
  const int E = h->E, PS = h->PS, NN = h->NN;
  cudaStream_t st = h->stream;
  int32_t gi;
  auto it = h->lindex.find(landmark_key);
  if (it == h->lindex.end()) {
    Group* g = new Group(&h->arena);
    g->landmark = true;
    g->lkey = landmark_key;
    g->id1 = g->id2 = key_chr(landmark_key);
    gi = (int32_t)h->groups.size();
    h->groups.push_back(g);
    h->lindex[landmark_key] = gi;
  } else {
    gi = it->second;
  }
  if (group_out) *group_out = gi;
  Group* g = h->groups[gi];
  if (reset && g->n > 0) {
    if (g->bits.p) H_CHECK_CUDA(h, cudaMemsetAsync(g->bits.p, 0, g->bits.cap, st));
    if (g->fl_count.p) H_CHECK_CUDA(h, cudaMemsetAsync(g->fl_count.p, 0, 8, st));
    g->n = 0;
    g->kfrom.clear(); g->kto.clear(); g->h_idxf.clear(); g->h_idxb.clear(); g->h_pfx.clear(); g->h_idxa0.clear();
  }
  if (n == 0) return RPGO_OK;
  if (h->traj_dirty) {
    /* same late-odometry rule as rpgo_lc_append: stored observations must see entries that appeared since */
    for (Group* q : h->groups) {
      for (int64_t k = 0; k < q->n; ++k) {
        q->h_idxf[k] = traj_lookup(h, q->kfrom[k]);
        q->h_idxb[k] = traj_lookup(h, q->kto[k]);
        if (q->landmark) q->h_idxa0[k] = traj_lookup(h, (uint64_t)key_chr(q->kfrom[k]) << 56);
      }
      if (q->n > 0 && q->idxf.p) {
        H_CHECK_CUDA(h, cudaMemcpyAsync(q->idxf.p, q->h_idxf.data(), (size_t)q->n * 4, cudaMemcpyHostToDevice, st));
        H_CHECK_CUDA(h, cudaMemcpyAsync(q->idxb.p, q->h_idxb.data(), (size_t)q->n * 4, cudaMemcpyHostToDevice, st));
        if (q->landmark && q->idxa0.p)
          H_CHECK_CUDA(h, cudaMemcpyAsync(q->idxa0.p, q->h_idxa0.data(), (size_t)q->n * 4, cudaMemcpyHostToDevice, st));
      }
      q->gathered = 0;
    }
    H_CHECK_CUDA(h, cudaStreamSynchronize(st));
    h->traj_dirty = false;
  }
  const int64_t o = g->n;
  int rc = ensure_group(h, g, o + n);
  if (rc != RPGO_OK) return rc;
  /* stage raw pose | cov, run the factor constructor (no odometry check), scatter into the group */
  const size_t o_cov = (size_t)n * PS * 8, o_if = o_cov + (size_t)n * NN * 8, o_ck = o_if + (size_t)n * 4,
               o_dst = (o_ck + (size_t)n + 15) & ~size_t(15), total = o_dst + (size_t)n * 8;
  char* pin = (char*)h->pin.ensure(total + 64);
  if (!pin) { h->err = "pinned staging allocation failed"; return RPGO_ERR_NOMEM; }
  memcpy(pin, pose, (size_t)n * PS * 8);
  memcpy(pin + o_cov, cov, (size_t)n * NN * 8);
  int32_t* p_if = (int32_t*)(pin + o_if);
  uint8_t* p_ck = (uint8_t*)(pin + o_ck);
  uint64_t* p_dst = (uint64_t*)(pin + o_dst);
  for (int64_t k = 0; k < n; ++k) {
    const uint8_t c = key_chr(pose_key[k]);
    const int32_t idx = traj_lookup(h, pose_key[k]);
    const uint64_t a0 = (uint64_t)c << 56;
    p_if[k] = idx;
    p_ck[k] = 0;
    if (idx == 0 && !h->key2idx.count(pose_key[k])) h->missing_refs.insert(pose_key[k]);
    if (!h->key2idx.count(a0)) h->missing_refs.insert(a0);
    h->prefixes.insert(c); /* odom_trajectories_[chr] is created by the lookup (Pcm.h:823-824) */
    g->kfrom.push_back(pose_key[k]);
    g->kto.push_back(landmark_key);
    g->h_idxf.push_back(idx);
    g->h_idxb.push_back(0);
    g->h_pfx.push_back(c);
    g->h_idxa0.push_back(traj_lookup(h, a0));
    p_dst[k] = (uint64_t)(uintptr_t)(g->lc.as<double>() + (size_t)(o + k) * E);
  }
  H_CHECK_CUDA(h, h->d_stage.ensure(total, 0, st));
  H_CHECK_CUDA(h, h->d_lcent.ensure((size_t)n * E * 8, 0, st));
  H_CHECK_CUDA(h, h->d_ok.ensure((size_t)n, 0, st));
  H_CHECK_CUDA(h, h->d_dist.ensure((size_t)n * 8, 0, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(h->d_stage.p, pin, total, cudaMemcpyHostToDevice, st));
  char* d = (char*)h->d_stage.p;
  launch_lc_prepare(h->dim, h->mode, (int)n, (const double*)d, (const double*)(d + o_cov), (const int32_t*)(d + o_if),
                    (const int32_t*)(d + o_if), (const uint8_t*)(d + o_ck), h->traj.as<double>(), h->th, h->d_lcent.as<double>(),
                    h->d_ok.as<uint8_t>(), h->d_dist.as<double>(), st);
  launch_scatter_entries(h->dim, (int)n, h->d_lcent.as<double>(), (const uint64_t*)(d + o_dst), st);
  h->launches += 2;
  g->n = o + n;
  H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxf.as<int32_t>() + o, g->h_idxf.data() + o, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxb.as<int32_t>() + o, g->h_idxb.data() + o, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(g->pfx.as<uint8_t>() + o, g->h_pfx.data() + o, (size_t)n, cudaMemcpyHostToDevice, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(g->idxa0.as<int32_t>() + o, g->h_idxa0.data() + o, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  rc = run_pairwise(h, g, o, nullptr);
  if (rc != RPGO_OK) return rc;
  rc = finalize_group(h, g, o);
  if (rc != RPGO_OK) return rc;
  H_CHECK_CUDA(h, cudaStreamSynchronize(st));
  return RPGO_OK;
}

int32_t rpgo_num_groups(rpgo_handle* h) { return h ? (int32_t)h->groups.size() : 0; }

int rpgo_group_info(rpgo_handle* h, int32_t g, uint8_t* id1, uint8_t* id2, int64_t* n) {
  if (!h || g < 0 || g >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  if (id1) *id1 = h->groups[g]->id1;
  if (id2) *id2 = h->groups[g]->id2;
  if (n) *n = h->groups[g]->n;
  return RPGO_OK;
}

int32_t rpgo_find_group(rpgo_handle* h, uint8_t id1, uint8_t id2) {
  if (!h) return -1;
  auto it = h->gindex.find({std::min(id1, id2), std::max(id1, id2)});
  return it == h->gindex.end() ? -1 : it->second;
}

int rpgo_lc_remove_last(rpgo_handle* h, int32_t gi, uint64_t* key_from, uint64_t* key_to) {
  if (!h || gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  DeviceGuard dg(h->device);
  Group* g = h->groups[gi];
  if (g->n <= 0 || g->landmark) return RPGO_ERR_NOT_FOUND; /* the reference never removes landmark observations */
  if (key_from) *key_from = g->kfrom.back();
  if (key_to) *key_to = g->kto.back();
  g->kfrom.pop_back();
  g->kto.pop_back();
  g->h_idxf.pop_back();
  g->h_idxb.pop_back();
  g->h_pfx.pop_back();
  g->n -= 1;
  if (g->gathered > g->n) g->gathered = g->n;
  if (h->loop_check && g->bits.p) {
    launch_clear_last(g->bits.as<uint32_t>(), g->stride32, (int)g->n, h->stream);
    launch_degree(g->bits.as<uint32_t>(), g->stride32, (int)g->n, g->deg.as<int32_t>(), h->stream);
    h->launches += 2;
    H_CHECK_CUDA(h, cudaGetLastError());
  }
  return RPGO_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* one clique search on an explicit scratch set / stream (shared by the single and the batched entry point) */
static int run_clique(rpgo_handle* h, Group* g, int32_t clique_mode, int64_t n_new, int64_t prev_size, CliqueSet* cset,
                      cudaStream_t st, CliqueShard cs, int32_t* ids_out, int64_t* size_out, int32_t* true_clique_out,
                      int64_t* launches, std::string* err, CliqueStats* stats = nullptr) {
  const int n = (int)g->n;
  CliqueScratch s;
  s.degmask = cset->degmask.as<uint32_t>();
  s.picks = cset->picks.as<int32_t>();
  s.elim = cset->elim.as<int32_t>();
  s.result = cset->result.as<int32_t>();
  s.ctl = cset->ctl.as<long long>();
  s.rwork = cset->rwork.as<uint32_t>();
  s.rwork_blocks = CLIQUE_BLOCKS;
  int r;
  if (clique_mode == RPGO_CLIQUE_HEU) {
    r = clique_heuristic(g->bits.as<uint32_t>(), g->stride32, n, g->deg.as<int32_t>(), 0, -1, s, ids_out, true_clique_out,
                         launches, st, cs, stats);
    if (r < -1) { *err = "clique_heuristic failed: " + std::to_string(r); return RPGO_ERR_CUDA; }
    *size_out = r;
  } else if (clique_mode == RPGO_CLIQUE_HEU_INCREMENTAL) {
    if (n_new < 0 || n_new > n || prev_size < 0) return RPGO_ERR_INVALID;
    r = clique_heuristic(g->bits.as<uint32_t>(), g->stride32, n, g->deg.as<int32_t>(), (int)(n - n_new), (int)prev_size, s,
                         ids_out, true_clique_out, launches, st, cs, stats);
    if (r < -1) { *err = "clique_heuristic failed: " + std::to_string(r); return RPGO_ERR_CUDA; }
    *size_out = (r > prev_size) ? r : 0; /* GraphUtils.cpp:40-43 */
  } else if (clique_mode == RPGO_CLIQUE_EXACT) {
    r = clique_exact(g->bits.as<uint32_t>(), g->stride32, n, g->deg.as<int32_t>(), s, ids_out, launches, st, cs);
    if (r < 0) { *err = "clique_exact failed: " + std::to_string(r); return RPGO_ERR_CUDA; }
    *size_out = r;
  } else {
    return RPGO_ERR_INVALID;
  }
  return RPGO_OK;
}

static cudaError_t ensure_clique_set(CliqueSet* c, int n, cudaStream_t st) {
  const int W = (n + 31) / 32;
  cudaError_t e;
  if ((e = c->degmask.ensure((size_t)W * 4 + 64, 0, st)) != cudaSuccess) return e;
  if ((e = c->picks.ensure((size_t)n * 4 + 64, 0, st)) != cudaSuccess) return e;
  if ((e = c->elim.ensure((size_t)n * 4 + 64, 0, st)) != cudaSuccess) return e;
  if ((e = c->result.ensure((size_t)n * 4 + 64, 0, st)) != cudaSuccess) return e;
  if ((e = c->ctl.ensure(clique_ctl_bytes(CLIQUE_BLOCKS), 0, st)) != cudaSuccess) return e;
  return c->rwork.ensure((size_t)CLIQUE_BLOCKS * 2 * n * 4 + (size_t)W * 8 + 64, 0, st);
}

int rpgo_find_inliers(rpgo_handle* h, int32_t gi, int32_t clique_mode, int64_t n_new, int64_t prev_size,
                      int32_t* ids_out, int64_t* size_out, int32_t* true_clique_out) {
  if (!h || !ids_out || !size_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  const int n = (int)g->n;
  if (n <= 0 || (!h->loop_check && !g->landmark)) { h->err = "find_inliers: empty group or loop check disabled"; return RPGO_ERR_INVALID; }
  cudaStream_t st = h->stream;
  H_CHECK_CUDA(h, ensure_clique_set(&h->cset, n, st));
  const CliqueShard cs = clique_shard(h, st);
  return run_clique(h, g, clique_mode, n_new, prev_size, &h->cset, st, cs, ids_out, size_out, true_clique_out, &h->launches,
                    &h->err, &h->last_clique);
}

int rpgo_clique_stats(rpgo_handle* h, int64_t* row_ands, int64_t* chains, int32_t* epochs) {
  if (!h) return RPGO_ERR_INVALID;
  if (row_ands) *row_ands = h->last_clique.row_ands;
  if (chains) *chains = h->last_clique.chains;
  if (epochs) *epochs = h->last_clique.epochs;
  return RPGO_OK;
}

/* Batched inlier selection: the groups of one removeOutliers() call are independent (Pcm::findInliers loops over
 * them, Pcm.h:858-876), so their searches run concurrently — a few host threads, each with its own stream and
 * scratch set, pull groups (largest first) from a shared counter.  With cfg.world > 1 and an exchange registered,
 * whole groups are assigned to ranks (entry k -> rank k mod world) and the results are combined with ONE all-reduce. */
int rpgo_find_inliers_batch(rpgo_handle* h, int32_t n_groups, const int32_t* groups, int32_t clique_mode,
                            const int64_t* n_new, const int64_t* prev_size, int32_t* ids_out, const int64_t* ids_offset,
                            int64_t* size_out) {
  if (!h || n_groups < 0 || (n_groups > 0 && (!groups || !ids_out || !ids_offset || !size_out))) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (n_groups == 0) return RPGO_OK;
  int max_n = 0;
  for (int k = 0; k < n_groups; ++k) {
    if (groups[k] < 0 || groups[k] >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
    Group* g = h->groups[groups[k]];
    if (g->n <= 0 || (!h->loop_check && !g->landmark)) { h->err = "find_inliers_batch: empty group or loop check disabled"; return RPGO_ERR_INVALID; }
    max_n = std::max<int>(max_n, (int)g->n);
  }
  const CliqueShard all = clique_shard(h, h->stream);
  const bool spread = all.active();
  const int world = spread ? h->cfg.world : 1, rank = spread ? h->cfg.rank : 0;
  /* workers: bounded by the scratch they need (the pick logs are CLIQUE_BLOCKS x n ints each) */
  const size_t per_set = (size_t)CLIQUE_BLOCKS * 2 * max_n * 4 + (size_t)max_n * 16;
  int T = (int)std::min<size_t>(8, std::max<size_t>(1, ((size_t)2 << 30) / std::max<size_t>(per_set, 1)));
  T = std::min(T, (n_groups + world - 1) / world);
  while ((int)h->workers.size() < T) {
    CliqueWorker* w = new CliqueWorker(&h->scratch);
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; h->err = "stream creation failed"; return RPGO_ERR_CUDA; }
    h->workers.push_back(w);
  }
  /* everything the searches read was produced on the handle's stream */
  cudaEvent_t ready;
  H_CHECK_CUDA(h, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  for (int t = 0; t < T; ++t) H_CHECK_CUDA(h, ensure_clique_set(&h->workers[t]->set, max_n, h->stream)); /* arena: this thread only */
  H_CHECK_CUDA(h, cudaEventRecord(ready, h->stream));
  for (int t = 0; t < T; ++t) H_CHECK_CUDA(h, cudaStreamWaitEvent(h->workers[t]->stream, ready, 0));
  /* this rank's entries, largest group first */
  std::vector<int> order;
  for (int k = 0; k < n_groups; ++k)
    if (k % world == rank) order.push_back(k);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->groups[groups[a]]->n > h->groups[groups[b]]->n; });
  std::atomic<int> next(0);
  std::vector<int> rcs(T, RPGO_OK);
  std::vector<std::string> errs(T);
  std::vector<int64_t> launches(T, 0);
  const int dev = h->device;
  auto body = [&](int t) {
    cudaSetDevice(dev); /* worker threads start on device 0 */
    CliqueWorker* w = h->workers[t];
    for (;;) {
      const int q = next.fetch_add(1);
      if (q >= (int)order.size()) break;
      const int k = order[q];
      Group* g = h->groups[groups[k]];
      const int rc = run_clique(h, g, clique_mode, n_new ? n_new[k] : 0, prev_size ? prev_size[k] : 0, &w->set, w->stream,
                                CliqueShard(), ids_out + ids_offset[k], size_out + k, nullptr, &launches[t], &errs[t]);
      if (rc != RPGO_OK) { rcs[t] = rc; break; }
    }
    cudaStreamSynchronize(w->stream);
  };
  if (T == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(body, t);
    for (auto& x : th) x.join();
  }
  cudaEventDestroy(ready);
  for (int t = 0; t < T; ++t) {
    h->launches += launches[t];
    if (rcs[t] != RPGO_OK) { h->err = errs[t]; return rcs[t]; }
  }
  if (spread) {
    /* one MAX all-reduce over [size_k, ids_k...] for all entries: non-owners contribute the minimum */
    std::vector<long long> buf;
    for (int k = 0; k < n_groups; ++k) {
      const bool mine = (k % world == rank);
      const int64_t cap = h->groups[groups[k]]->n;
      buf.push_back(mine ? (long long)size_out[k] : LLONG_MIN);
      for (int64_t i = 0; i < cap; ++i)
        buf.push_back(mine && i < std::max<int64_t>(size_out[k], 0) ? (long long)ids_out[ids_offset[k] + i] : LLONG_MIN);
    }
    if (all.xchg(RPGO_XCHG_MAX_I64, buf.data(), (int64_t)buf.size(), 0) != 0) { h->err = "exchange failed"; return RPGO_ERR_CUDA; }
    size_t pos = 0;
    for (int k = 0; k < n_groups; ++k) {
      const int64_t cap = h->groups[groups[k]]->n;
      size_out[k] = (int64_t)buf[pos++];
      for (int64_t i = 0; i < cap; ++i, ++pos)
        if (i < std::max<int64_t>(size_out[k], 0)) ids_out[ids_offset[k] + i] = (int32_t)buf[pos];
    }
  }
  return RPGO_OK;
}

int rpgo_set_exchange(rpgo_handle* h, rpgo_exchange_fn fn, void* user) {
  if (!h) return RPGO_ERR_INVALID;
  h->xchg = fn;
  h->xchg_user = user;
  return RPGO_OK;
}

/* ---- multi-GPU data plane (NCCL behind the ABI) -------------------------------------------------- */
int rpgo_comm_unique_id(void* id_out) {
  if (!id_out) return RPGO_ERR_INVALID;
  std::string err;
  if (comm_unique_id(id_out, &err) != 0) {
    fprintf(stderr, "rpgo_comm_unique_id: %s\n", err.c_str());
    return RPGO_ERR_CUDA;
  }
  return RPGO_OK;
}

int rpgo_comm_init(rpgo_handle* h, const void* id, int32_t rank, int32_t world) {
  if (!h || !id) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (rank != h->cfg.rank || world != h->cfg.world || world < 2) {
    h->err = "rpgo_comm_init: rank/world must equal the handle's cfg.rank/cfg.world (world >= 2)";
    return RPGO_ERR_INVALID;
  }
  if (h->comm) { comm_destroy(h->comm); h->comm = nullptr; }
  std::string err;
  if (comm_create(&h->comm, id, rank, world, &err) != 0) {
    h->err = "rpgo_comm_init: " + err;
    h->comm = nullptr;
    return RPGO_ERR_CUDA;
  }
  return RPGO_OK;
}

int rpgo_comm_destroy(rpgo_handle* h) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (h->comm) {
    cudaStreamSynchronize(h->stream);
    comm_destroy(h->comm);
    h->comm = nullptr;
  }
  return RPGO_OK;
}

int rpgo_group_allgather(rpgo_handle* h, int32_t gi) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  if (h->cfg.world <= 1) return finalize_group(h, h->groups[gi], 0);
  int rc = allgather_group(h, h->groups[gi]);
  if (rc == RPGO_OK) rc = finalize_group(h, h->groups[gi], 0);
  return rc;
}

/* ---------------------------------------------------------------------------------------------- */
int rpgo_adj_bits(rpgo_handle* h, int32_t gi, uint64_t* rows_out, int64_t stride_words) {
  if (!h || !rows_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  const int64_t n = g->n;
  if (stride_words < (n + 63) / 64) return RPGO_ERR_INVALID;
  memset(rows_out, 0, (size_t)n * stride_words * 8);
  if (!g->bits.p || n == 0) return RPGO_OK;
  const size_t width = (size_t)((n + 31) / 32) * 4;
  H_CHECK_CUDA(h, cudaMemcpy2DAsync(rows_out, (size_t)stride_words * 8, g->bits.p, (size_t)g->stride32 * 4, width, (size_t)n,
                                    cudaMemcpyDeviceToHost, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  return RPGO_OK;
}

int rpgo_adj_bits_device(rpgo_handle* h, int32_t gi, void** bits_device, int64_t* stride_words, int64_t* n) {
  if (!h) return RPGO_ERR_INVALID;
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (bits_device) *bits_device = g->bits.p;
  if (stride_words) *stride_words = g->stride32 / 2;
  if (n) *n = g->n;
  return RPGO_OK;
}

int rpgo_degrees(rpgo_handle* h, int32_t gi, int32_t* deg_out) {
  if (!h || !deg_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (g->n == 0) return RPGO_OK;
  H_CHECK_CUDA(h, cudaMemcpyAsync(deg_out, g->deg.p, (size_t)g->n * 4, cudaMemcpyDeviceToHost, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  return RPGO_OK;
}

int rpgo_near_threshold(rpgo_handle* h, int32_t gi, int32_t* pairs_out, int64_t cap, int64_t* n_out) {
  if (!h || !n_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  *n_out = 0;
  unsigned long long c = 0;
  if (g->fl_count.p) {
    H_CHECK_CUDA(h, cudaMemcpyAsync(&c, g->fl_count.p, 8, cudaMemcpyDeviceToHost, h->stream));
    H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  const int64_t mine = std::min<int64_t>((int64_t)c, FLAG_CAP);
  if (!(h->cfg.world > 1 && h->comm)) {
    *n_out = (int64_t)c;
    const int64_t m = std::min<int64_t>(mine, cap);
    if (pairs_out && m > 0) {
      H_CHECK_CUDA(h, cudaMemcpyAsync(pairs_out, g->fl_pairs.p, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
      H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RPGO_OK;
  }
  /* sharded rows: every rank flagged the pairs of its own row chunks; the list handed out is the union, rank by rank
   * (collective: every rank calls this with the same arguments) */
  const int W = h->cfg.world;
  std::vector<long long> counts((size_t)2 * W, 0);
  counts[h->cfg.rank] = (long long)c;
  counts[W + h->cfg.rank] = (long long)mine;
  if (comm_allreduce_i64_host(h->comm, counts.data(), counts.size(), true, h->stream) != 0) {
    h->err = std::string("near_threshold exchange failed: ") + comm_error(h->comm);
    return RPGO_ERR_CUDA;
  }
  int64_t total = 0, written = 0;
  for (int r = 0; r < W; ++r) total += counts[r];
  *n_out = total;
  std::vector<int32_t> buf;
  for (int r = 0; r < W; ++r) {
    const int64_t m = counts[W + r];
    if (m == 0) continue;
    buf.assign((size_t)m * 2, 0);
    if (r == h->cfg.rank) {
      H_CHECK_CUDA(h, cudaMemcpyAsync(buf.data(), g->fl_pairs.p, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
      H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (comm_bcast_host(h->comm, buf.data(), (size_t)m * 8, r, h->stream) != 0) {
      h->err = std::string("near_threshold exchange failed: ") + comm_error(h->comm);
      return RPGO_ERR_CUDA;
    }
    const int64_t take = std::min<int64_t>(m, cap - written);
    if (pairs_out && take > 0) memcpy(pairs_out + written * 2, buf.data(), (size_t)take * 8);
    written += std::max<int64_t>(take, 0);
  }
  return RPGO_OK;
}

int rpgo_pair_distances(rpgo_handle* h, int32_t gi, double* dist_out) {
  if (!h || !dist_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  const int64_t n = g->n;
  if (n > 8192) { h->err = "pair_distances is a debug call: n <= 8192"; return RPGO_ERR_INVALID; }
  /* debug buffer: allocated and freed directly (not from the arena, which never releases) */
  struct Tmp {
    double* p = nullptr;
    cudaStream_t st;
    ~Tmp() { if (p) cudaFreeAsync(p, st); }
    double* as() { return p; }
  } dd;
  dd.st = h->stream;
  H_CHECK_CUDA(h, cudaMallocAsync((void**)&dd.p, std::max<size_t>((size_t)n * n * 8, 8), h->stream));
  /* count is reset so that the recompute does not double-count flagged pairs */
  unsigned long long saved = 0;
  if (g->fl_count.p) H_CHECK_CUDA(h, cudaMemcpyAsync(&saved, g->fl_count.p, 8, cudaMemcpyDeviceToHost, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  /* force all rows: temporarily single-GPU sharding */
  const int w = h->cfg.world, r = h->cfg.rank;
  h->cfg.world = 1; h->cfg.rank = 0;
  int rc = run_pairwise(h, g, 0, dd.as());
  h->cfg.world = w; h->cfg.rank = r;
  if (rc != RPGO_OK) return rc;
  if (g->fl_count.p) H_CHECK_CUDA(h, cudaMemcpyAsync(g->fl_count.p, &saved, 8, cudaMemcpyHostToDevice, h->stream));
  H_CHECK_CUDA(h, cudaMemcpyAsync(dist_out, dd.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, h->stream));
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  return RPGO_OK;
}

int rpgo_group_recompute(rpgo_handle* h, int32_t gi, int64_t j_begin) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (j_begin < 0 || j_begin > g->n) return RPGO_ERR_INVALID;
  if (g->fl_count.p) H_CHECK_CUDA(h, cudaMemsetAsync(g->fl_count.p, 0, 8, h->stream));
  int rc = run_pairwise(h, g, j_begin, nullptr);
  if (rc != RPGO_OK) return rc;
  if (h->cfg.world <= 1) rc = finalize_group(h, g, j_begin);
  return rc;
}

int rpgo_group_pairwise(rpgo_handle* h, int32_t gi, int64_t j_begin) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (j_begin < 0 || j_begin > g->n) return RPGO_ERR_INVALID;
  if (g->fl_count.p) H_CHECK_CUDA(h, cudaMemsetAsync(g->fl_count.p, 0, 8, h->stream));
  return run_pairwise(h, g, j_begin, nullptr);
}

int rpgo_group_finalize(rpgo_handle* h, int32_t gi) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  return finalize_group(h, h->groups[gi], 0);
}

int rpgo_debug_pass(rpgo_handle* h, int32_t gi, int32_t which) {
  if (!h) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (!g->bits.p || g->n < 1) return RPGO_ERR_INVALID;
  if (which == 0) launch_mirror(g->bits.as<uint32_t>(), g->stride32, (int)g->n, 0, h->stream);
  else if (which == 1) launch_degree(g->bits.as<uint32_t>(), g->stride32, (int)g->n, g->deg.as<int32_t>(), h->stream);
  else return RPGO_ERR_INVALID;
  h->launches += 1;
  H_CHECK_CUDA(h, cudaGetLastError());
  return RPGO_OK;
}

int rpgo_group_chunking(rpgo_handle* h, int32_t gi, int64_t* chunk_rows, int64_t* padded_rows) {
  if (!h) return RPGO_ERR_INVALID;
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Shard s = group_shard(h, h->groups[gi]);
  if (chunk_rows) *chunk_rows = s.chunk_rows;
  if (padded_rows) *padded_rows = s.chunk_rows * 2 * s.world;
  return RPGO_OK;
}

int rpgo_debug_load_group(rpgo_handle* h, uint8_t id1, uint8_t id2, int64_t n, const uint64_t* rows,
                          int64_t stride_words, int32_t* group_out) {
  if (!h || n < 0 || (n > 0 && !rows) || stride_words < (n + 63) / 64) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (!h->loop_check) return RPGO_ERR_INVALID;
  const std::pair<uint8_t, uint8_t> id(std::min(id1, id2), std::max(id1, id2));
  int32_t gi;
  auto it = h->gindex.find(id);
  if (it == h->gindex.end()) {
    Group* g = new Group(&h->arena);
    g->id1 = id.first;
    g->id2 = id.second;
    gi = (int32_t)h->groups.size();
    h->groups.push_back(g);
    h->gindex[id] = gi;
  } else {
    gi = it->second;
  }
  Group* g = h->groups[gi];
  g->n = 0;
  int rc = ensure_group(h, g, std::max<int64_t>(n, 1));
  if (rc != RPGO_OK) return rc;
  H_CHECK_CUDA(h, cudaMemsetAsync(g->bits.p, 0, g->bits.cap, h->stream));
  if (n > 0)
    H_CHECK_CUDA(h, cudaMemcpy2DAsync(g->bits.p, (size_t)g->stride32 * 4, rows, (size_t)stride_words * 8,
                                      (size_t)((n + 31) / 32) * 4, (size_t)n, cudaMemcpyHostToDevice, h->stream));
  g->n = n;
  g->gathered = 0;
  g->kfrom.assign((size_t)n, 0);
  g->kto.assign((size_t)n, 0);
  g->h_idxf.assign((size_t)n, 0);
  g->h_idxb.assign((size_t)n, 0);
  g->h_pfx.assign((size_t)n, id.first);
  launch_degree(g->bits.as<uint32_t>(), g->stride32, (int)n, g->deg.as<int32_t>(), h->stream);
  h->launches++;
  H_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
  if (group_out) *group_out = gi;
  return RPGO_OK;
}

int rpgo_frame_align_measurements(rpgo_handle* h, int32_t gi, uint8_t r0, int64_t m, const int32_t* closure_idx,
                                  double* T_out) {
  if (!h || m < 0 || (m > 0 && (!closure_idx || !T_out))) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (gi < 0 || gi >= (int32_t)h->groups.size()) return RPGO_ERR_NOT_FOUND;
  Group* g = h->groups[gi];
  if (g->landmark) return RPGO_ERR_INVALID;
  if (m == 0) return RPGO_OK;
  for (int64_t t = 0; t < m; ++t) {
    const int32_t c = closure_idx[t];
    if (c < 0 || c >= g->n) return RPGO_ERR_INVALID;
    if (!h->key2idx.count(g->kfrom[c]) || !h->key2idx.count(g->kto[c])) {
      h->err = "frame alignment: closure key without trajectory entry";
      return RPGO_ERR_NOT_FOUND;
    }
  }
  cudaStream_t st = h->stream;
  const size_t ib = ((size_t)m * 4 + 255) & ~size_t(255);
  H_CHECK_CUDA(h, h->d_stage.ensure(ib + (size_t)m * h->PS * 8, 0, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(h->d_stage.p, closure_idx, (size_t)m * 4, cudaMemcpyHostToDevice, st));
  double* d_out = (double*)((char*)h->d_stage.p + ib);
  launch_frame_align(h->dim, group_view(h, g), h->traj.as<double>(), h->E, r0, (int)m, (const int32_t*)h->d_stage.p, d_out, st);
  h->launches++;
  H_CHECK_CUDA(h, cudaGetLastError());
  H_CHECK_CUDA(h, cudaMemcpyAsync(T_out, d_out, (size_t)m * h->PS * 8, cudaMemcpyDeviceToHost, st));
  H_CHECK_CUDA(h, cudaStreamSynchronize(st));
  return RPGO_OK;
}

int rpgo_robot_odom_values(rpgo_handle* h, uint8_t prefix, const double* transform, int64_t cap, uint64_t* keys_out,
                           double* poses_out, int64_t* n_out) {
  if (!h || !n_out) return RPGO_ERR_INVALID;
  DeviceGuard dg(h->device);
  std::vector<std::pair<uint64_t, int32_t>> ent;
  for (auto& kv : h->key2idx)
    if (key_chr(kv.first) == prefix) ent.push_back({kv.first, kv.second});
  std::sort(ent.begin(), ent.end());
  *n_out = (int64_t)ent.size();
  if (!keys_out && !poses_out) return RPGO_OK;
  if (cap < (int64_t)ent.size() || !keys_out || !poses_out) return RPGO_ERR_INVALID;
  const int64_t m = (int64_t)ent.size();
  if (m == 0) return RPGO_OK;
  std::vector<int32_t> idx((size_t)m);
  for (int64_t i = 0; i < m; ++i) { keys_out[i] = ent[i].first; idx[i] = ent[i].second; }
  std::vector<double> T(h->PS, 0.0);
  if (transform) memcpy(T.data(), transform, sizeof(double) * h->PS);
  else if (h->dim == 3) { T[0] = T[4] = T[8] = 1.0; } else { T[0] = 1.0; }
  cudaStream_t st = h->stream;
  const size_t ib = ((size_t)m * 4 + 255) & ~size_t(255), tb = 256;
  H_CHECK_CUDA(h, h->d_stage.ensure(ib + tb + (size_t)m * h->PS * 8, 0, st));
  char* d = (char*)h->d_stage.p;
  H_CHECK_CUDA(h, cudaMemcpyAsync(d, idx.data(), (size_t)m * 4, cudaMemcpyHostToDevice, st));
  H_CHECK_CUDA(h, cudaMemcpyAsync(d + ib, T.data(), sizeof(double) * h->PS, cudaMemcpyHostToDevice, st));
  double* d_out = (double*)(d + ib + tb);
  launch_transform_poses(h->dim, h->traj.as<double>(), h->E, (int)m, (const int32_t*)d, (const double*)(d + ib), d_out, st);
  h->launches++;
  H_CHECK_CUDA(h, cudaGetLastError());
  H_CHECK_CUDA(h, cudaMemcpyAsync(poses_out, d_out, (size_t)m * h->PS * 8, cudaMemcpyDeviceToHost, st));
  H_CHECK_CUDA(h, cudaStreamSynchronize(st));
  return RPGO_OK;
}

int rpgo_debug_check_fastmath(int64_t n, uint64_t seed, uint64_t* mismatches, uint64_t* checked) {
  if (!mismatches || !checked || n <= 0) return RPGO_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return RPGO_ERR_CUDA;
  unsigned long long m = 0, c = 0;
  const int rc = fastmath_check((long long)n, (unsigned long long)seed, &m, &c, 0);
  *mismatches = m;
  *checked = c;
  return rc == 0 ? RPGO_OK : RPGO_ERR_CUDA;
}

int rpgo_fp64_peak(int32_t device, double* tflops_out) {
  if (!tflops_out) return RPGO_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return RPGO_ERR_CUDA;
  if (device >= ndev) return RPGO_ERR_INVALID;
  int cur = 0;
  cudaGetDevice(&cur);
  DeviceGuard dg(device >= 0 ? device : cur);
  cudaStream_t st;
  if (cudaStreamCreate(&st) != cudaSuccess) return RPGO_ERR_CUDA;
  *tflops_out = fp64_peak_tflops(st);
  cudaStreamDestroy(st);
  return *tflops_out > 0 ? RPGO_OK : RPGO_ERR_CUDA;
}

}  /* extern "C" */
