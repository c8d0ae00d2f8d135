/* clique_kernels.cu — inlier selection over the bitset adjacency (sm_100a).
 *
 *   K4  heuristic max clique   FMC::maxCliqueHeu / maxCliqueHeuIncremental
 *                              (reference include/KimeraRPGO/max_clique_finder/findCliqueHeu.cpp:32-209,
 *                               wrappers src/utils/GraphUtils.cpp:19-44), including the scratch-buffer
 *                               contents the reference hands back (findCliqueHeu.cpp:110-113) which
 *                               Pcm.h:865-869 consumes as "inlier indices".
 *
 * Reference algorithm per candidate v (ascending), with running bound maxClq:
 *     skip if maxClq > deg(v);  S = [v] ++ [u in N(v) ascending : deg(u) >= maxClq]
 *     repeat { pick = S.back(); S = [u in S : u in N(pick)]; icc++ } until S empty
 *     if icc > maxClq { out = scratch buffer; maxClq = icc }
 * Bitset formulation: R = N(v) & {deg >= maxClq}; pick = highest set bit of R; R &= N(pick);
 * icc = 1 + number of picks.  Two exact accelerations:
 *   - bound: steps + |R| + 1 <= maxClq  =>  icc cannot exceed maxClq  =>  the candidate cannot change
 *     the reference's state, stop early;
 *   - window resolution: all picks that fall into the current top 32-bit word of R are resolved by one
 *     warp from a single 32x32 sub-block of the adjacency, then the remaining words are ANDed with all
 *     of those picks' rows in one parallel sweep (one dependent memory round per window, not per pick).
 * The sequential dependence between candidates (maxClq) is honoured by rounds: every candidate is
 * evaluated against a snapshot of maxClq; the first candidate that improves it commits, later
 * candidates are re-evaluated against the new bound (identical to the sequential order of events).
 */
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include <cooperative_groups.h>

#include "kernels.cuh"

namespace rpgo {

static constexpr int HEU_THREADS = 128;       /* block size of the clique kernels for n <= HEU_WIDE_N */
static constexpr int HEU_THREADS_WIDE = 512;  /* wide rows (n > HEU_WIDE_N): 4x the sweep parallelism per chain */
static constexpr int HEU_WIDE_N = 131072; /* measured: 512-thread blocks lose at 50k (27 vs 15 ms), win at 200k (911 vs 1038 ms) */
/* measurement build (-DRPGO_CLIQUE_COUNTERS): ctl[1] chains started, ctl[2] windows, ctl[3] adjacency/degree-mask bytes read */
#ifdef RPGO_CLIQUE_COUNTERS
#define RPGO_COUNT_BYTES(ctl, x) atomicAdd((ctl) + 3, (unsigned long long)(x))
#else
#define RPGO_COUNT_BYTES(ctl, x) ((void)0)
#endif
static constexpr int HEU_PRE = 13; /* words per thread covered by the sweep prefetch (13 * 128 * 32 = 53k vertices) */

__global__ void degmask_kernel(const int32_t* __restrict__ deg, int n, int M, uint32_t* mask, int words) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  uint32_t m = 0;
#pragma unroll 4
  for (int b = 0; b < 32; ++b) {
    const int u = w * 32 + b;
    if (u < n && deg[u] >= M) m |= 1u << b;
  }
  mask[w] = m;
}

/* block-wide (sum, max) reduction; returns to all threads */
__device__ __forceinline__ void block_sum_max(int& s, int& m, int* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  s = __reduce_add_sync(0xffffffffu, s);
  m = __reduce_max_sync(0xffffffffu, m);
  __syncthreads(); /* protect sh from the previous use */
  if (lane == 0) {
    sh[wid] = s;
    sh[32 + wid] = m;
  }
  __syncthreads();
  int ts = 0, tm = -1;
  for (int i = 0; i < nw; ++i) {
    ts += sh[i];
    tm = max(tm, sh[32 + i]);
  }
  s = ts;
  m = tm;
}


static constexpr int HEU_LIST_K = 4;                       /* list elements per thread */
static constexpr int HEU_LIST_MAX = HEU_LIST_K * HEU_THREADS; /* list mode below this many survivors (any block size) */

/* Tail of a greedy chain once at most HEU_LIST_MAX candidates survive: the survivors are written to shared
 * memory as a list in DESCENDING id order, so that "highest set bit of R" becomes "first live list entry"; each
 * pick then costs one adjacency word per live entry instead of a sweep over the whole bitset.  Same picks, same
 * order, same count as the bitset loop.  Returns the number of picks made (-1: a lower candidate already
 * improved, give up; -2: cannot beat M). */
template <int TH>
__device__ int heu_list_tail(const uint32_t* __restrict__ bits, int64_t stride32, const uint32_t* R, int top, int cnt,
                             int steps, int M, int v, unsigned long long* ctl, int32_t* my_picks, int32_t* L, int* sh) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  /* build the list: thread t owns the word range [top - (t+1)*C + 1, top - t*C] scanned from the top */
  const int C = (top + TH) / TH;
  int mine = 0;
  for (int c = 0; c < C; ++c) {
    const int w = top - tid * C - c;
    if (w >= 0) mine += __popc(R[w]);
  }
  /* exclusive prefix over threads (ascending tid = descending ids) */
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();
  if (lane == 31) sh[wid] = incl;
  __syncthreads();
  int off = incl - mine;
  for (int i = 0; i < wid; ++i) off += sh[i];
  for (int c = 0; c < C; ++c) {
    const int w = top - tid * C - c;
    if (w >= 0) {
      uint32_t r = R[w];
      while (r) {
        const int b = 31 - __clz(r);
        r &= ~(1u << b);
        L[off++] = w * 32 + b;
      }
    }
  }
  __syncthreads();
  constexpr int LK = (HEU_LIST_MAX + TH - 1) / TH; /* list elements per thread */
  int u[LK];
  bool alive[LK];
#pragma unroll
  for (int k = 0; k < LK; ++k) {
    const int pos = k * TH + tid;
    alive[k] = pos < cnt;
    u[k] = alive[k] ? L[pos] : 0;
  }
  int m = cnt, made = 0, iter = 0;
  while (m > 0) {
    if (steps + made + m + 1 <= M) return -2;
    /* first live entry (block-wide minimum position) and live count */
    int first = INT_MAX, live = 0;
#pragma unroll
    for (int k = 0; k < LK; ++k) {
      if (alive[k]) {
        first = min(first, k * TH + tid);
        ++live;
      }
    }
    first = __reduce_min_sync(0xffffffffu, first);
    live = __reduce_add_sync(0xffffffffu, live);
    __syncthreads();
    if (lane == 0) {
      sh[wid] = first;
      sh[32 + wid] = live;
    }
    if (tid == 0) sh[16] = ((++iter & 15) == 0 && (long long)(*(volatile unsigned long long*)ctl >> 32) < (long long)v) ? 1 : 0;
    __syncthreads();
    first = INT_MAX;
    live = 0;
#pragma unroll
    for (int i = 0; i < TH / 32; ++i) {
      first = min(first, sh[i]);
      live += sh[32 + i];
    }
    if (sh[16]) return -1;
    m = live;
    if (m == 0) break;
    if (steps + made + m + 1 <= M) return -2;
    const int pick = L[first];
    if (tid == 0) my_picks[steps + made] = pick;
    ++made;
    const uint32_t* prow = bits + (size_t)pick * stride32;
#pragma unroll
    for (int k = 0; k < LK; ++k) {
      if (alive[k]) {
        const int pos = k * TH + tid;
        if (pos == first) alive[k] = false;
        else {
          alive[k] = (prow[u[k] >> 5] >> (u[k] & 31)) & 1u;
          RPGO_COUNT_BYTES(ctl, 4);
        }
      }
    }
    m -= 1; /* refined at the top of the next iteration */
  }
  return made;
}

/* un-mark the cached verdict of every neighbour of the vertices in xs (their filter membership changed) */
__global__ void undead_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n, const int32_t* __restrict__ xs, int nx,
                              int32_t* dead_flags) {
  const int u = xs[blockIdx.x];
  const int W = (n + 31) / 32;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    uint32_t r = bits[(size_t)u * stride32 + w];
    while (r) {
      const int b = __ffs(r) - 1;
      r &= r - 1;
      dead_flags[w * 32 + b] = 0;
    }
  }
  if (threadIdx.x == 0) dead_flags[u] = 0;
}

/* ctl[0]: packed (candidate << 32 | icc) of the lowest-index improving candidate of this round
 *         (ULLONG_MAX = none).  picks_block: per-block pick log (n ints each). */
template <int TH>
__device__ __forceinline__ void heu_round_body(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                               const int32_t* __restrict__ deg, const uint32_t* degmask, int first, int vstep,
                                               int M, unsigned long long* ctl, int32_t* picks_block, int32_t* dead_flags) {
  extern __shared__ uint32_t R[];
  __shared__ int sh[64];
  __shared__ int s_abort;
  __shared__ int32_t s_list[HEU_LIST_MAX];
  const int W = (n + 31) / 32;
  const int tid = threadIdx.x;
  int32_t* my_picks = picks_block + (size_t)blockIdx.x * n;

  /* candidates first, first + vstep, ...: vstep > 1 when the candidates are partitioned over ranks */
  for (long long vv = first + (long long)blockIdx.x * vstep; vv < n; vv += (long long)gridDim.x * vstep) {
    const int v = (int)vv;
    /* a lower-index candidate already improved: everything from here on is re-evaluated next round */
    __syncthreads();
    if (tid == 0) s_abort = ((long long)(*(volatile unsigned long long*)ctl >> 32) < (long long)v) ? 1 : 0;
    __syncthreads();
    if (s_abort) return;
    if (M > deg[v]) continue; /* pruning 1 */
    if (dead_flags[v]) continue; /* evaluated under an equivalent bound before: cannot improve (see host driver) */
    int cnt = 0, top = -1;
    for (int w = tid; w < W; w += blockDim.x) {
      const uint32_t r = bits[(size_t)v * stride32 + w] & degmask[w];
      RPGO_COUNT_BYTES(ctl, 8);
      R[w] = r;
      cnt += __popc(r);
      if (r) top = w;
    }
    block_sum_max(cnt, top, sh);
    if (cnt + 1 <= M) {
      if (tid == 0) dead_flags[v] = 1;
      continue; /* cannot exceed the bound */
    }
#ifdef RPGO_CLIQUE_COUNTERS
    if (tid == 0) atomicAdd(ctl + 1, 1ULL);
#endif
    int steps = 0;
    bool dead = false;
    int iter = 0;
    while (cnt > 0) {
      if (steps + cnt + 1 <= M) { dead = true; break; }
      if (cnt <= HEU_LIST_MAX) {
        const int made = heu_list_tail<TH>(bits, stride32, R, top, cnt, steps, M, v, ctl, my_picks, s_list, sh);
        if (made == -1) return;
        if (made == -2) { dead = true; break; }
        steps += made;
        break;
      }
      if (tid == 0) s_abort = ((++iter & 7) == 0 && (long long)(*(volatile unsigned long long*)ctl >> 32) < (long long)v) ? 1 : 0;
      const int t = top;
#ifdef RPGO_CLIQUE_COUNTERS
      if (tid == 0) atomicAdd(ctl + 2, 1ULL);
#endif
      const uint32_t T = R[t];
      const int lane = tid & 31;
      /* every warp resolves the window redundantly from the same 32x32 adjacency block (no block barrier
       * between resolution and sweep), and the sweep's row words are requested BEFORE the resolution so
       * that both dependent global accesses overlap: one memory round trip per window */
      uint32_t rw = 0;
      if ((T >> lane) & 1u) rw = bits[(size_t)(t * 32 + lane) * stride32 + t];
      if (tid < 32 && ((T >> lane) & 1u)) RPGO_COUNT_BYTES(ctl, 4);
      const int nT = __popc(T);
      uint32_t pre[4][HEU_PRE]; /* up to 4 candidate rows x HEU_PRE words per thread are prefetched */
      const bool prefetch = (nT <= 4) && (t <= HEU_PRE * TH);
      if (prefetch) {
        uint32_t q = T;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int b = q ? 31 - __clz(q) : -1;
          if (b >= 0) q &= ~(1u << b);
#pragma unroll
          for (int k = 0; k < HEU_PRE; ++k) {
            const int w = tid + k * TH;
            pre[c][k] = (b >= 0 && w < t) ? bits[(size_t)(t * 32 + b) * stride32 + w] : 0xffffffffu;
            if (b >= 0 && w < t) RPGO_COUNT_BYTES(ctl, 4);
          }
        }
      }
      uint32_t cur = T, P = 0;
      int k0 = 0;
      while (cur) {
        const int b = 31 - __clz(cur);
        P |= 1u << b;
        const uint32_t rb = __shfl_sync(0xffffffffu, rw, b);
        cur &= rb & ~(1u << b);
        if (tid == 0) my_picks[steps + k0] = t * 32 + b;
        ++k0;
      }
      __syncthreads(); /* all warps have read R[t] / s_abort is visible */
      if (s_abort) return;
      if (tid == 0) R[t] = 0;
      steps += __popc(P);
      cnt = 0;
      top = -1;
      if (prefetch) {
        /* AND mask of the candidate rows that turned out to be picks */
#pragma unroll
        for (int k = 0; k < HEU_PRE; ++k) {
          const int w = tid + k * TH;
          if (w < t) {
            uint32_t r = R[w];
            if (r) {
              uint32_t q = T;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int b = q ? 31 - __clz(q) : -1;
                if (b >= 0) {
                  q &= ~(1u << b);
                  if ((P >> b) & 1u) r &= pre[c][k];
                }
              }
              R[w] = r;
              cnt += __popc(r);
              if (r) top = w;
            }
          }
        }
      } else {
        for (int w = tid; w < t; w += blockDim.x) {
          uint32_t r = R[w];
          if (r) {
            uint32_t q = P;
            while (q && r) {
              const int b = 31 - __clz(q);
              q &= ~(1u << b);
              r &= bits[(size_t)(t * 32 + b) * stride32 + w];
              RPGO_COUNT_BYTES(ctl, 4);
            }
            R[w] = r;
            cnt += __popc(r);
            if (r) top = w;
          }
        }
      }
      block_sum_max(cnt, top, sh);
    }
    if (dead) {
      if (tid == 0) dead_flags[v] = 1;
      continue;
    }
    const int icc = steps + 1;
    if (icc > M) {
      if (tid == 0) atomicMin(ctl, ((unsigned long long)(unsigned)v << 32) | (unsigned)icc);
      return; /* later candidates of this block are > v: re-evaluated next round */
    }
    if (tid == 0) dead_flags[v] = 1; /* completed chain that did not beat M */
  }
}

template <int TH>
__global__ void __launch_bounds__(TH) heu_round_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                                const int32_t* __restrict__ deg,
                                                                const uint32_t* __restrict__ degmask, int first, int vstep,
                                                                int M, unsigned long long* ctl, int32_t* picks_block,
                                                                int32_t* dead_flags) {
  heu_round_body<TH>(bits, stride32, n, deg, degmask, first, vstep, M, ctl, picks_block, dead_flags);
}

/* All rounds of one search in ONE cooperative launch (single-rank searches): the bound update, the winner's pick log and
 * the invalidation of cached verdicts move onto the device, separated by grid-wide barriers, so a search costs one launch
 * and one host synchronisation instead of ~7 API calls per round.  Same rounds, same order of events as the host loop. */
struct HeuResult {
  int M, winner, winner_M, winner_icc, rounds, pad;
};

template <int TH>
__global__ void __launch_bounds__(TH) heu_persistent_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                                     const int32_t* __restrict__ deg, uint32_t* degmask,
                                                                     int first, int maxclq0, unsigned long long* ctl,
                                                                     int32_t* picks_block, int32_t* dead_flags,
                                                                     int32_t* picks_out, HeuResult* result) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int W = (n + 31) / 32;
  const int tid = threadIdx.x;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gthreads = (long long)gridDim.x * blockDim.x;
  int M = maxclq0, start = first < 0 ? 0 : first;
  int winner = -1, winner_M = 0, winner_icc = 0, rounds = 0;
  while (start < n) {
    /* degree filter of this bound, control word */
    for (long long w = gtid; w < W; w += gthreads) {
      uint32_t m = 0;
#pragma unroll 4
      for (int b = 0; b < 32; ++b) {
        const int u = (int)w * 32 + b;
        if (u < n && deg[u] >= M) m |= 1u << b;
      }
      degmask[w] = m;
    }
    if (gtid == 0) ctl[0] = ~0ULL;
    grid.sync();
    heu_round_body<TH>(bits, stride32, n, deg, degmask, start, 1, M, ctl, picks_block, dead_flags);
    grid.sync();
    ++rounds;
    const unsigned long long c = *(volatile unsigned long long*)ctl;
    if (c == ~0ULL) break;
    const int v = (int)(c >> 32), icc = (int)(c & 0xffffffffu);
    /* the winner's block keeps its pick log */
    if ((int)blockIdx.x == (v - start) % (int)gridDim.x) {
      const int32_t* my_picks = picks_block + (size_t)blockIdx.x * n;
      for (int k = tid; k < icc - 1; k += blockDim.x) picks_out[k] = my_picks[k];
    }
    /* a vertex with M <= deg < icc leaves the degree filter: cached verdicts of its neighbours are no longer valid */
    for (int u = blockIdx.x; u < n; u += gridDim.x) {
      const int d = deg[u];
      if (d >= M && d < icc) {
        for (int w = tid; w < W; w += blockDim.x) {
          uint32_t r = bits[(size_t)u * stride32 + w];
          while (r) {
            const int b = __ffs(r) - 1;
            r &= r - 1;
            dead_flags[w * 32 + b] = 0;
          }
        }
        if (tid == 0) dead_flags[u] = 0;
      }
    }
    winner = v;
    winner_M = M;
    winner_icc = icc;
    M = icc;
    start = v + 1;
    grid.sync();
  }
  if (gtid == 0) {
    result->M = M;
    result->winner = winner;
    result->winner_M = winner_M;
    result->winner_icc = winner_icc;
    result->rounds = rounds;
  }
}

/* elimination step of every vertex of the winner's initial list:
 * e(u) = k such that u leaves the list when pick p_k is applied (p_1 > p_2 > ... > p_K), 0 if u was
 * never in the list.  u leaves at the first (= highest-id) pick it is not adjacent to.
 * pset: bitset of picks; above[w]: number of picks in words > w. */
__global__ void heu_elim_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n, int v,
                                const uint32_t* __restrict__ degmask, const uint32_t* __restrict__ pset,
                                const int32_t* __restrict__ above, int32_t* elim) {
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (u >= n) return;
  const int W = (n + 31) / 32;
  const bool in_r0 = ((bits[(size_t)v * stride32 + (u >> 5)] & degmask[u >> 5]) >> (u & 31)) & 1u;
  if (!in_r0) {
    if (lane == 0) elim[u] = 0;
    return;
  }
  int e = 0;
  for (int base = ((W - 1) / 32) * 32; base >= 0; base -= 32) {
    const int w = base + lane;
    uint32_t z = 0;
    if (w < W) z = pset[w] & ~bits[(size_t)u * stride32 + w];
    const unsigned nz = __ballot_sync(0xffffffffu, z != 0);
    if (nz) {
      const int hl = 31 - __clz(nz);
      const uint32_t zz = __shfl_sync(0xffffffffu, z, hl);
      const int ww = base + hl;
      const int b = 31 - __clz(zz);
      const uint32_t pw = pset[ww];
      const int higher = (b == 31) ? 0 : __popc(pw >> (b + 1));
      e = above[ww] + higher + 1;
      break;
    }
  }
  if (lane == 0) elim[u] = e;
}

/* out[p], p = 1..K: the (p-1)-th smallest u with elim[u] > kstar[p] (see DESIGN.md "scratch-buffer replay") */
__global__ void heu_select_kernel(int n, int K, const int32_t* __restrict__ elim, const int32_t* __restrict__ kstar,
                                  int32_t* out) {
  const int p = 1 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p > K) return;
  const int ks = kstar[p];
  int need = p - 1; /* rank to find */
  int found = -1;
  for (int base = 0; base < n; base += 32) {
    const int u = base + lane;
    const bool in = (u < n) && (elim[u] > ks);
    const unsigned m = __ballot_sync(0xffffffffu, in);
    const int c = __popc(m);
    if (need < c) {
      /* the need-th set bit of m */
      unsigned mm = m;
      for (int i = 0; i < need; ++i) mm &= mm - 1;
      found = base + (__ffs(mm) - 1);
      break;
    }
    need -= c;
  }
  if (lane == 0) out[p] = found;
}

#define CUCHECK(x)                                   \
  do {                                               \
    cudaError_t e_ = (x);                            \
    if (e_ != cudaSuccess) return -(int)e_ - 1000;   \
  } while (0)

/* Host driver.  Returns the reference's return value (maxClq); ids_out_host gets the first maxClq
 * entries of the reference's returned buffer; true_out_host (optional) the greedy clique itself. */
int clique_heuristic(const uint32_t* bits, int64_t stride32, int n, const int32_t* deg, int first, int maxclq0,
                     CliqueScratch s, int32_t* ids_out_host, int32_t* true_out_host, int64_t* launches,
                     cudaStream_t st, CliqueShard cs) {
  const bool sharded = cs.active();
  const int world = sharded ? cs.world : 1, rank = sharded ? cs.rank : 0;
  if (n <= 0) return -1;
  const int W = (n + 31) / 32;
  const size_t smem = (size_t)W * sizeof(uint32_t);
  if (smem > 200 * 1024) return -2; /* n > ~1.6M closures in one group: not supported by this kernel */
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    cudaFuncSetAttribute(heu_round_kernel<HEU_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(heu_round_kernel<HEU_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  static const char* wide_env = getenv("RPGO_CLIQUE_WIDE"); /* A/B knob: 0 / 1 forces the block size */
  const bool wide = wide_env ? (wide_env[0] == '1') : (n > HEU_WIDE_N);
  const int threads = wide ? HEU_THREADS_WIDE : HEU_THREADS;
  /* "dead" cache: a candidate that could not beat bound M cannot beat any M' >= M as long as its filtered
   * neighbourhood {u : deg(u) >= M} is unchanged, i.e. as long as no vertex has M <= deg < M'.  The host checks
   * that on the sorted degree list at every bound change and clears the cache otherwise. */
  const bool trace = getenv("RPGO_CLIQUE_TRACE") != nullptr;
  auto now_ms = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  const double t_begin = now_ms();
  double t_round0 = 0;
  static const bool host_loop = getenv("RPGO_CLIQUE_HOSTLOOP") != nullptr;
  const bool persistent = !sharded && !trace && !host_loop;
  std::vector<int32_t> hdeg;
  CUCHECK(cudaMemsetAsync(s.elim, 0, sizeof(int32_t) * n, st));
  if (!persistent) { /* the host loop needs the degrees for the cache invalidation */
    hdeg.resize(n);
    CUCHECK(cudaMemcpyAsync(hdeg.data(), deg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CUCHECK(cudaStreamSynchronize(st));
  }
  std::vector<int32_t> changed; /* vertices whose filter membership changes between two bounds */
  int M = maxclq0;
  int winner = -1, winner_M = 0, winner_icc = 0, winner_block = 0;
  bool winner_local = true;
  int start = first < 0 ? 0 : first;
  if (persistent) {
    /* single rank: all rounds in one cooperative launch */
    static PerDeviceOnce attr2;
    if (attr2.first()) {
      cudaFuncSetAttribute(heu_persistent_kernel<HEU_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(heu_persistent_kernel<HEU_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const void* pk = wide ? (const void*)heu_persistent_kernel<HEU_THREADS_WIDE> : (const void*)heu_persistent_kernel<HEU_THREADS>;
    int dev = 0, sms = 0, bps = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (wide) CUCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, heu_persistent_kernel<HEU_THREADS_WIDE>, HEU_THREADS_WIDE, smem));
    else CUCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, heu_persistent_kernel<HEU_THREADS>, HEU_THREADS, smem));
    if (bps >= 1) {
      int grid = bps * sms;
      if (grid > (int)s.rwork_blocks) grid = (int)s.rwork_blocks;
      /* small groups take a slice of the machine so that independent searches (batched entry point) run side by side */
      const int want = std::max(64, (n - start + 7) / 8);
      if (grid > want) grid = want;
      if (grid > n - start) grid = std::max(1, n - start);
      HeuResult* d_res = (HeuResult*)((unsigned long long*)s.ctl + 4);
      const uint32_t* a_bits = bits;
      int64_t a_stride = stride32;
      int a_n = n, a_first = start, a_M = maxclq0;
      const int32_t* a_deg = deg;
      uint32_t* a_degmask = s.degmask;
      unsigned long long* a_ctl = (unsigned long long*)s.ctl;
      int32_t* a_rwork = (int32_t*)s.rwork;
      int32_t* a_elim = s.elim;
      int32_t* a_picks = s.picks;
      void* args[] = {&a_bits, &a_stride, &a_n, &a_deg, &a_degmask, &a_first, &a_M, &a_ctl, &a_rwork, &a_elim, &a_picks, &d_res};
      CUCHECK(cudaLaunchCooperativeKernel(pk, dim3(grid), dim3(threads), args, smem, st));
      *launches += 1;
      HeuResult hr;
      CUCHECK(cudaMemcpyAsync(&hr, d_res, sizeof(hr), cudaMemcpyDeviceToHost, st));
      CUCHECK(cudaStreamSynchronize(st));
      M = hr.M;
      winner = hr.winner;
      winner_M = hr.winner_M;
      winner_icc = hr.winner_icc;
      start = n; /* the host loop below is skipped */
    }
  }
  std::vector<int32_t> picks_host;
  const int grid_cap = (int)s.rwork_blocks;
  unsigned long long h_ctl;
  if (trace) fprintf(stderr, "[clique] setup (degree copy + sort) %.3f ms\n", now_ms() - t_begin);
  while (start < n) {
    t_round0 = now_ms();
    const unsigned long long none = ~0ULL;
    const unsigned long long init3[4] = {none, 0ULL, 0ULL, 0ULL};
    CUCHECK(cudaMemcpyAsync(s.ctl, init3, sizeof(init3), cudaMemcpyHostToDevice, st));
    degmask_kernel<<<(W + 127) / 128, 128, 0, st>>>(deg, n, M, s.degmask, W);
    /* candidates of this rank: v = start (mod nothing) for one GPU, v = rank (mod world) when partitioned */
    const int v0 = start + (((rank - start) % world) + world) % world;
    int grid = (n - v0 + world - 1) / world;
    if (grid > grid_cap) grid = grid_cap;
    if (grid > 0) {
      if (wide)
        heu_round_kernel<HEU_THREADS_WIDE><<<grid, HEU_THREADS_WIDE, smem, st>>>(bits, stride32, n, deg, s.degmask, v0, world, M,
                                                                             (unsigned long long*)s.ctl, (int32_t*)s.rwork, s.elim);
      else
        heu_round_kernel<HEU_THREADS><<<grid, HEU_THREADS, smem, st>>>(bits, stride32, n, deg, s.degmask, v0, world, M,
                                                                     (unsigned long long*)s.ctl, (int32_t*)s.rwork, s.elim);
      *launches += 1;
    }
    *launches += 1;
    CUCHECK(cudaMemcpyAsync(&h_ctl, s.ctl, sizeof(h_ctl), cudaMemcpyDeviceToHost, st));
    CUCHECK(cudaStreamSynchronize(st));
    bool local_winner = true;
    if (sharded) {
      /* incumbent exchange: the lowest-index improving candidate over all ranks wins the round */
      const unsigned long long mine = h_ctl;
      long long key = (h_ctl == none) ? LLONG_MAX : (long long)h_ctl;
      if (cs.xchg(RPGO_XCHG_MIN_I64, &key, 1, 0) != 0) return -3;
      h_ctl = (key == LLONG_MAX) ? none : (unsigned long long)key;
      local_winner = (h_ctl == mine);
    }
    if (trace) {
      fprintf(stderr, "[clique] rank %d round kernel+sync %.3f ms\n", rank, now_ms() - t_round0);
      unsigned long long c[4];
      cudaMemcpy(c, s.ctl, sizeof(c), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[clique] round start=%d M=%d grid=%d -> improver=%lld icc=%d | chains started %llu, windows %llu, bytes %llu\n", start, M, grid,
              h_ctl == none ? -1LL : (long long)(h_ctl >> 32), (int)(h_ctl & 0xffffffffu), c[1], c[2], c[3]);
    }
    if (h_ctl == none) break;
    winner = (int)(h_ctl >> 32);
    winner_icc = (int)(h_ctl & 0xffffffffu);
    winner_M = M;
    winner_local = local_winner;
    if (local_winner) {
      winner_block = ((winner - v0) / world) % grid;
      /* keep the winner's pick log (its block may be reused next round) */
      CUCHECK(cudaMemcpyAsync(s.picks, (int32_t*)s.rwork + (size_t)winner_block * n,
                              sizeof(int32_t) * (size_t)(winner_icc - 1 > 0 ? winner_icc - 1 : 0), cudaMemcpyDeviceToDevice, st));
    }
    {
      /* any vertex with M <= deg < winner_icc changes the filter: cached verdicts are no longer valid */
      changed.clear();
      for (int u = 0; u < n && changed.size() <= 4096; ++u)
        if (hdeg[u] >= M && hdeg[u] < winner_icc) changed.push_back(u);
      const int nx = (int)changed.size();
      if (nx > 4096) {
        CUCHECK(cudaMemsetAsync(s.elim, 0, sizeof(int32_t) * n, st));
      } else if (nx > 0) {
        /* only candidates adjacent to one of those vertices see a different filtered neighbourhood */
        CUCHECK(cudaMemcpyAsync(s.result, changed.data(), sizeof(int32_t) * nx, cudaMemcpyHostToDevice, st));
        CUCHECK(cudaStreamSynchronize(st)); /* `changed` is reused next round */
        undead_kernel<<<nx, 128, 0, st>>>(bits, stride32, n, s.result, nx, s.elim);
        *launches += 1;
      }
    }
    M = winner_icc;
    start = winner + 1;
  }
  if (winner < 0) return M; /* no candidate improved on maxclq0 (incremental mode) */
  const double t_replay = now_ms();
  if (trace) fprintf(stderr, "[clique] rounds done at %.3f ms\n", t_replay - t_begin);

  const int K = winner_icc - 1;
  if (!winner_local) {
    /* the final winning chain ran on another rank: replay that one candidate here (same bound, same filter) to get
     * its pick log; intermediate winners only moved the bound and need no log */
    const unsigned long long init3[4] = {~0ULL, 0ULL, 0ULL, 0ULL};
    CUCHECK(cudaMemcpyAsync(s.ctl, init3, sizeof(init3), cudaMemcpyHostToDevice, st));
    CUCHECK(cudaMemsetAsync(s.elim + winner, 0, sizeof(int32_t), st));
    degmask_kernel<<<(W + 127) / 128, 128, 0, st>>>(deg, n, winner_M, s.degmask, W);
    if (wide)
      heu_round_kernel<HEU_THREADS_WIDE><<<1, HEU_THREADS_WIDE, smem, st>>>(bits, stride32, n, deg, s.degmask, winner, n, winner_M,
                                                                        (unsigned long long*)s.ctl, (int32_t*)s.rwork, s.elim);
    else
      heu_round_kernel<HEU_THREADS><<<1, HEU_THREADS, smem, st>>>(bits, stride32, n, deg, s.degmask, winner, n, winner_M,
                                                                (unsigned long long*)s.ctl, (int32_t*)s.rwork, s.elim);
    *launches += 2;
    CUCHECK(cudaMemcpyAsync(s.picks, (int32_t*)s.rwork, sizeof(int32_t) * (size_t)(K > 0 ? K : 0), cudaMemcpyDeviceToDevice, st));
  }
  picks_host.resize(K > 0 ? K : 1);
  if (K > 0) CUCHECK(cudaMemcpyAsync(picks_host.data(), s.picks, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
  CUCHECK(cudaStreamSynchronize(st));
  if (true_out_host) {
    true_out_host[0] = winner;
    for (int k = 0; k < K; ++k) true_out_host[1 + k] = picks_host[k];
  }
  ids_out_host[0] = winner;
  if (K == 0) return M;

  /* ---- scratch-buffer replay ---- */
  std::vector<uint32_t> pset(W, 0u);
  for (int k = 0; k < K; ++k) pset[picks_host[k] >> 5] |= 1u << (picks_host[k] & 31);
  std::vector<int32_t> above(W, 0);
  for (int w = W - 2; w >= 0; --w) above[w] = above[w + 1] + __builtin_popcount(pset[w + 1]);
  /* device temporaries reuse s.rwork (the pick logs are no longer needed) */
  uint32_t* d_pset = s.rwork;
  int32_t* d_above = (int32_t*)(s.rwork + W);
  int32_t* d_kstar = d_above + W;
  CUCHECK(cudaMemcpyAsync(d_pset, pset.data(), sizeof(uint32_t) * W, cudaMemcpyHostToDevice, st));
  CUCHECK(cudaMemcpyAsync(d_above, above.data(), sizeof(int32_t) * W, cudaMemcpyHostToDevice, st));
  degmask_kernel<<<(W + 127) / 128, 128, 0, st>>>(deg, n, winner_M, s.degmask, W);
  heu_elim_kernel<<<(n + 7) / 8, 256, 0, st>>>(bits, stride32, n, winner, s.degmask, d_pset, d_above, s.elim);
  *launches += 2;
  std::vector<int32_t> elim(n);
  CUCHECK(cudaMemcpyAsync(elim.data(), s.elim, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  CUCHECK(cudaStreamSynchronize(st));
  /* cnt[k] = #{u : elim[u] > k}, k = 0..K  (list k has 1 + cnt[k] entries) */
  std::vector<int64_t> hist(K + 2, 0);
  for (int u = 0; u < n; ++u)
    if (elim[u] > 0) hist[elim[u] <= K ? elim[u] : K + 1]++;
  std::vector<int64_t> cnt(K + 1, 0);
  {
    int64_t acc = hist[K + 1];
    for (int k = K; k >= 0; --k) {
      cnt[k] = acc; /* elim > k */
      acc += hist[k];
    }
  }
  /* kstar[p] = max{k <= K : cnt[k] >= p}; cnt is non-increasing, cnt[K - p] >= p always */
  std::vector<int32_t> kstar(K + 1, 0);
  {
    int k = K;
    for (int p = 1; p <= K; ++p) {
      while (k > 0 && cnt[k] < p) --k;
      kstar[p] = k;
    }
  }
  CUCHECK(cudaMemcpyAsync(d_kstar, kstar.data(), sizeof(int32_t) * (K + 1), cudaMemcpyHostToDevice, st));
  heu_select_kernel<<<(K + 7) / 8, 256, 0, st>>>(n, K, s.elim, d_kstar, s.result);
  *launches += 1;
  CUCHECK(cudaMemcpyAsync(ids_out_host + 1, s.result + 1, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
  CUCHECK(cudaStreamSynchronize(st));
  if (trace) fprintf(stderr, "[clique] replay %.3f ms, total %.3f ms\n", now_ms() - t_replay, now_ms() - t_begin);
  return M;
}

}  // namespace rpgo
