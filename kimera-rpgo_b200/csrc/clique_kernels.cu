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
 * icc(v) = 1 + number of picks.
 *
 * EPOCHS.  maxClq enters a candidate's chain only through the degree filter {deg >= maxClq} (and the skip test, which is
 * the same set).  Let M be the bound at some point and hi = min{deg(u) : deg(u) >= M}: for every bound M' in [M, hi]
 * the filter is the same set, so every chain evaluated while the bound stays <= hi is a fixed function icc(v) of the
 * candidate alone.  The sequential loop therefore decomposes into epochs (start, M):
 *   (a) v* = the first candidate >= start with icc(v) > hi, if any: everything before it only raised the bound within
 *       [M, hi]; v* commits (bound icc(v*), filter changes) and the next epoch starts at v* + 1;
 *   (b) otherwise the loop ends in this epoch and its result is the FIRST candidate attaining max icc (if > M).
 * Both are order-free searches, so all candidates of an epoch are evaluated concurrently, with two exact prunings on the
 * chain's upper bound ub = picks so far + |R| + 1:  a chain is abandoned when ub <= hi (cannot be v*) AND (ub, v) cannot
 * beat the best completed chain so far (cannot be the first arg-max); and everything beyond a known v* is abandoned.
 * Dense PCM graphs (all degrees far above the clique size) need two epochs — the first ends at the first candidate
 * because the initial bound -1 admits every vertex — where the round-per-improvement formulation of round 1 needed one
 * dependent winner chain per improvement (13 at 50k closures).
 *   - window resolution: all picks that fall into the current top 32-bit word of R are resolved by one
 *     warp from a single 32x32 sub-block of the adjacency, then the remaining words are ANDed with all
 *     of those picks' rows in one parallel sweep (one dependent memory round per window, not per pick);
 *   - verdicts "icc(v) <= hi" are cached across epochs and invalidated only for neighbours of vertices that leave the
 *     filter.
 * Multi-GPU: candidates are partitioned over ranks (v mod world); an epoch ends with ONE all-reduce of the two control
 * words (lowest v*, best chain) over NCCL / the caller's exchange function.
 */
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include <cooperative_groups.h>

#include "comm.h"
#include "kernels.cuh"

namespace rpgo {

static constexpr int HEU_THREADS = 128;       /* block size of the clique kernels for n <= HEU_WIDE_N */
static constexpr int HEU_THREADS_WIDE = 512;  /* wide rows (n > HEU_WIDE_N): 4x the sweep parallelism per chain */
static constexpr int HEU_WIDE_N = 131072;
static constexpr int HEU_PRE = 13; /* words per thread covered by the sweep prefetch (13 * 128 * 32 = 53k vertices) */

typedef unsigned long long ull;

/* control block of one search (device memory, 8-byte words):
 *   set e = epoch % 3 at words [4e .. 4e+2]: vstar, best, hi
 *   [12] row ANDs (statistics: algorithmic bytes = row_ands * n / 8), [13] chains started
 *   [16..] HeuResult;  byte 256..: saved_sel[block] (which of the block's two pick logs holds its best chain) */
static constexpr int CTL_ROW_ANDS = 12, CTL_CHAINS = 13, CTL_RESULT = 16, CTL_SAVED_SEL_BYTES = 256;
static constexpr ull HEU_NONE = ~0ULL;

__device__ __host__ __forceinline__ ull best_key(int icc, int v) {
  return ((ull)(unsigned)(icc + 1) << 32) | (ull)(0xFFFFFFFFu - (unsigned)v);
}
/* can a chain of candidate v whose length is at most ub still matter in this epoch? */
__device__ __forceinline__ bool heu_relevant(int ub, int v, int hi, ull best) { return ub > hi || best_key(ub, v) > best; }

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

/* shared snapshot of the epoch's control words, refreshed by thread 0 between two block barriers */
struct HeuShared {
  ull best;
  int vstar_v; /* candidate index of the lowest known v*, INT_MAX if none */
  int pad;
};
__device__ __forceinline__ void heu_refresh(HeuShared* hs, const ull* cset) {
  const ull vs = *(volatile const ull*)&cset[0];
  hs->vstar_v = (vs == HEU_NONE) ? INT_MAX : (int)(vs >> 32);
  hs->best = *(volatile const ull*)&cset[1];
}

/* Tail of a greedy chain once at most HEU_LIST_MAX candidates survive: the survivors are written to shared
 * memory as a list in DESCENDING id order, so that "highest set bit of R" becomes "first live list entry"; each
 * pick then costs one adjacency word per live entry instead of a sweep over the whole bitset.  Same picks, same
 * order, same count as the bitset loop.  Returns the number of picks made (-1: a lower candidate already ended the
 * epoch, give up; -2: the chain cannot matter any more). */
template <int TH>
__device__ int heu_list_tail(const uint32_t* __restrict__ bits, int64_t stride32, const uint32_t* R, int top, int cnt,
                             int steps, int hi, int v, const ull* cset, HeuShared* hs, int32_t* my_picks, int32_t* L, int* sh) {
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
  ull best = hs->best;
  int m = cnt, made = 0;
  while (m > 0) {
    if (!heu_relevant(steps + made + m + 1, v, hi, best)) return -2;
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
    if (tid == 0) heu_refresh(hs, cset);
    __syncthreads();
    first = INT_MAX;
    live = 0;
#pragma unroll
    for (int i = 0; i < TH / 32; ++i) {
      first = min(first, sh[i]);
      live += sh[32 + i];
    }
    if (hs->vstar_v < v) return -1;
    best = hs->best;
    m = live;
    if (m == 0) break;
    if (!heu_relevant(steps + made + m + 1, v, hi, best)) return -2;
    const int pick = L[first];
    if (tid == 0) my_picks[steps + made] = pick;
    ++made;
    const uint32_t* prow = bits + (size_t)pick * stride32;
#pragma unroll
    for (int k = 0; k < LK; ++k) {
      if (alive[k]) {
        const int pos = k * TH + tid;
        if (pos == first) alive[k] = false;
        else alive[k] = (prow[u[k] >> 5] >> (u[k] & 31)) & 1u;
      }
    }
    m -= 1; /* refined at the top of the next iteration */
  }
  return made;
}

/* One epoch's candidate evaluation for this block: candidates first + blockIdx*vstep, stride gridDim*vstep (vstep > 1
 * when the candidates are partitioned over ranks).  cset = this epoch's control words {vstar, best, hi}.  The block keeps
 * two pick logs; *sel says which one holds the block's best completed chain (the other one is written). */
template <int TH>
__device__ __forceinline__ void heu_round_body(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                               const int32_t* __restrict__ deg, const uint32_t* degmask, int first, int vstep,
                                               int M, ull* cset, ull* stats, int32_t* picks_block, int32_t* dead_flags,
                                               int32_t* saved_sel) {
  extern __shared__ uint32_t R[];
  __shared__ int sh[64];
  __shared__ HeuShared hs;
  __shared__ int s_flag;
  __shared__ int32_t s_list[HEU_LIST_MAX];
  const int W = (n + 31) / 32;
  const int tid = threadIdx.x;
  int sel = saved_sel[blockIdx.x]; /* log `sel` is kept, log `1 - sel` is the one being written */
  const ull hi64 = cset[2];        /* written in the prepare phase, constant during the round */
  const int hi = hi64 > (ull)INT_MAX ? INT_MAX : (int)hi64;

  for (long long vv = first + (long long)blockIdx.x * vstep; vv < n; vv += (long long)gridDim.x * vstep) {
    const int v = (int)vv;
    int32_t* my_picks = picks_block + ((size_t)blockIdx.x * 2 + (size_t)(1 - sel)) * n;
    __syncthreads();
    if (tid == 0) heu_refresh(&hs, cset);
    __syncthreads();
    if (hs.vstar_v < v) return; /* the epoch ends before v: everything from here on belongs to the next one */
    ull best = hs.best;
    if (M > deg[v]) continue;      /* pruning 1 (skip test of the reference) */
    if (dead_flags[v]) continue;   /* icc(v) <= an earlier epoch's horizon under the same filtered neighbourhood */
    int cnt = 0, top = -1;
    for (int w = tid; w < W; w += blockDim.x) {
      const uint32_t r = bits[(size_t)v * stride32 + w] & degmask[w];
      R[w] = r;
      cnt += __popc(r);
      if (r) top = w;
    }
    block_sum_max(cnt, top, sh);
    int steps = 0;
    bool dead = !heu_relevant(cnt + 1, v, hi, best);
    if (tid == 0 && !dead) atomicAdd(&stats[CTL_CHAINS - CTL_ROW_ANDS], 1ULL);
    while (!dead && cnt > 0) {
      if (cnt <= HEU_LIST_MAX) {
        const int made = heu_list_tail<TH>(bits, stride32, R, top, cnt, steps, hi, v, cset, &hs, my_picks, s_list, sh);
        if (made == -1) return;
        if (made == -2) { dead = true; break; }
        steps += made;
        break;
      }
      if (tid == 0) heu_refresh(&hs, cset);
      const int t = top;
      const uint32_t T = R[t];
      const int lane = tid & 31;
      /* every warp resolves the window redundantly from the same 32x32 adjacency block (no block barrier
       * between resolution and sweep), and the sweep's row words are requested BEFORE the resolution so
       * that both dependent global accesses overlap: one memory round trip per window */
      uint32_t rw = 0;
      if ((T >> lane) & 1u) rw = bits[(size_t)(t * 32 + lane) * stride32 + t];
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
      __syncthreads(); /* all warps have read R[t]; the control snapshot is visible */
      if (hs.vstar_v < v) return;
      best = hs.best;
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
      } else if (t <= HEU_PRE * TH) {
        /* many picks in this window: AND all of their rows into every surviving word, four picks at a time.  All row words
         * of a group (4 picks x up to HEU_PRE words of this thread) are requested before any is used: one memory round trip
         * per group of picks instead of one per word */
        uint32_t r[HEU_PRE];
#pragma unroll
        for (int k = 0; k < HEU_PRE; ++k) {
          const int w = tid + k * TH;
          r[k] = w < t ? R[w] : 0u;
        }
        uint32_t q = P;
        while (q) {
          int b4[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            b4[c] = q ? 31 - __clz(q) : -1;
            if (b4[c] >= 0) q &= ~(1u << b4[c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < HEU_PRE; ++k)
              pre[c][k] = (b4[c] >= 0 && r[k]) ? bits[(size_t)(t * 32 + b4[c]) * stride32 + tid + k * TH] : 0xffffffffu;
          }
#pragma unroll
          for (int k = 0; k < HEU_PRE; ++k) r[k] &= (pre[0][k] & pre[1][k]) & (pre[2][k] & pre[3][k]);
        }
#pragma unroll
        for (int k = 0; k < HEU_PRE; ++k) {
          const int w = tid + k * TH;
          if (w < t) {
            R[w] = r[k];
            cnt += __popc(r[k]);
            if (r[k]) top = w;
          }
        }
      } else {
        /* rows wider than the register window (n > HEU_PRE * TH * 32): word by word, four picks in flight */
        for (int w = tid; w < t; w += blockDim.x) {
          uint32_t r = R[w];
          if (r) {
            uint32_t q = P;
            while (q) {
              uint32_t m4[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int b = q ? 31 - __clz(q) : -1;
                if (b >= 0) q &= ~(1u << b);
                m4[c] = b >= 0 ? bits[(size_t)(t * 32 + b) * stride32 + w] : 0xffffffffu;
              }
              r &= (m4[0] & m4[1]) & (m4[2] & m4[3]);
            }
            R[w] = r;
            cnt += __popc(r);
            if (r) top = w;
          }
        }
      }
      block_sum_max(cnt, top, sh);
      if (!heu_relevant(steps + cnt + 1, v, hi, best)) dead = true;
    }
    if (tid == 0) atomicAdd(&stats[0], (ull)(1 + steps));
    if (dead) {
      if (tid == 0) dead_flags[v] = 1; /* icc(v) <= hi: stays true for every later bound while the neighbourhood filter is unchanged */
      continue;
    }
    const int icc = steps + 1;
    if (icc > hi) {
      /* ends the epoch (unless a lower candidate does): the log stays in the buffer being written */
      if (tid == 0) atomicMin(&cset[0], ((ull)(unsigned)v << 32) | (unsigned)icc);
      return; /* later candidates of this block are > v: they belong to the next epoch */
    }
    __syncthreads();
    if (tid == 0) {
      const ull key = best_key(icc, v);
      const ull old = atomicMax(&cset[1], key);
      s_flag = old < key ? 1 : 0;
      dead_flags[v] = 1;
    }
    __syncthreads();
    if (s_flag) { /* best completed chain so far: keep its log, write the following ones into the other buffer */
      sel = 1 - sel;
      if (tid == 0) saved_sel[blockIdx.x] = sel;
    }
  }
}

/* Epoch preparation, executed by the whole grid: (1) cached verdicts of the neighbours of every vertex that left the
 * degree filter (M_prev <= deg < M) are dropped, (2) the filter mask of bound M, (3) hi = min{deg >= M}. */
__device__ __forceinline__ void heu_prepare_body(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                 const int32_t* __restrict__ deg, int M_prev, int M, bool invalidate,
                                                 uint32_t* degmask, ull* cset, int32_t* dead_flags) {
  const int W = (n + 31) / 32;
  const int tid = threadIdx.x;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gthreads = (long long)gridDim.x * blockDim.x;
  int lo = INT_MAX;
  for (long long w = gtid; w < W; w += gthreads) {
    uint32_t m = 0;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
      const int u = (int)w * 32 + b;
      if (u < n) {
        const int d = deg[u];
        if (d >= M) {
          m |= 1u << b;
          lo = min(lo, d);
        }
      }
    }
    degmask[w] = m;
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  if ((tid & 31) == 0 && lo != INT_MAX) atomicMin(&cset[2], (ull)lo);
  if (invalidate) {
    for (int u = blockIdx.x; u < n; u += gridDim.x) {
      const int d = deg[u];
      if (d >= M_prev && d < M) {
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
  }
}

__global__ void heu_prepare_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n, const int32_t* __restrict__ deg,
                                   int M_prev, int M, int invalidate, uint32_t* degmask, ull* cset, int32_t* dead_flags) {
  heu_prepare_body(bits, stride32, n, deg, M_prev, M, invalidate != 0, degmask, cset, dead_flags);
}

#ifndef RPGO_HEU_MINB /* resident blocks per SM the compiler has to leave room for (128-thread blocks) */
#define RPGO_HEU_MINB 5
#endif
template <int TH>
__global__ void __launch_bounds__(TH, TH == HEU_THREADS ? RPGO_HEU_MINB : 1) heu_round_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                                const int32_t* __restrict__ deg,
                                                                const uint32_t* __restrict__ degmask, int first, int vstep,
                                                                int M, ull* ctl, int32_t* picks_block, int32_t* dead_flags) {
  heu_round_body<TH>(bits, stride32, n, deg, degmask, first, vstep, M, ctl, ctl + CTL_ROW_ANDS, picks_block, dead_flags,
                     (int32_t*)((char*)ctl + CTL_SAVED_SEL_BYTES));
}

/* copy the pick log of candidate `v` (evaluated by block `blk`) to picks_out: which = 0 the log being written when the
 * block stopped (a v* chain), 1 the block's saved best chain */
__global__ void heu_copy_log_kernel(const int32_t* __restrict__ picks_block, const ull* ctl, int blk, int which, int n, int K,
                                    int32_t* picks_out) {
  const int32_t* saved_sel = (const int32_t*)((const char*)ctl + CTL_SAVED_SEL_BYTES);
  const int sel = saved_sel[blk];
  const int32_t* src = picks_block + ((size_t)blk * 2 + (size_t)(which ? sel : 1 - sel)) * n;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) picks_out[k] = src[k];
}

/* All epochs of one search in ONE cooperative launch (single-rank searches): preparation, candidate evaluation and the
 * epoch decision are separated by grid-wide barriers, so a search costs one launch and one host synchronisation. */
struct HeuResult {
  int M, winner, winner_M, winner_icc, epochs, pad;
};

template <int TH>
__global__ void __launch_bounds__(TH, TH == HEU_THREADS ? RPGO_HEU_MINB : 1) heu_persistent_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                                     const int32_t* __restrict__ deg, uint32_t* degmask,
                                                                     int first, int maxclq0, ull* ctl,
                                                                     int32_t* picks_block, int32_t* dead_flags,
                                                                     int32_t* picks_out) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int tid = threadIdx.x;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid;
  int32_t* saved_sel = (int32_t*)((char*)ctl + CTL_SAVED_SEL_BYTES);
  HeuResult* result = (HeuResult*)(ctl + CTL_RESULT);
  int M = maxclq0, M_prev = maxclq0, start = first < 0 ? 0 : first;
  int winner = -1, winner_M = 0, winner_icc = 0, epochs = 0;
  /* the host initialised set 0 completely and the horizon word of set 1 */
  while (start < n) {
    ull* cset = ctl + 4 * (epochs % 3);
    heu_prepare_body(bits, stride32, n, deg, M_prev, M, epochs > 0, degmask, cset, dead_flags);
    grid.sync();
    heu_round_body<TH>(bits, stride32, n, deg, degmask, start, 1, M, cset, ctl + CTL_ROW_ANDS, picks_block, dead_flags, saved_sel);
    grid.sync();
    const ull vs = *(volatile ull*)&cset[0];
    const ull bs = *(volatile ull*)&cset[1];
    const int e = epochs++;
    if (vs != HEU_NONE) {
      const int v = (int)(vs >> 32), icc = (int)(vs & 0xffffffffu);
      if ((int)blockIdx.x == (v - start) % (int)gridDim.x) {
        const int32_t* src = picks_block + ((size_t)blockIdx.x * 2 + (size_t)(1 - saved_sel[blockIdx.x])) * n;
        for (int k = tid; k < icc - 1; k += blockDim.x) picks_out[k] = src[k];
      }
      winner = v;
      winner_M = M;
      winner_icc = icc;
      M_prev = M;
      M = icc;
      start = v + 1;
      if (gtid == 0) {
        ull* nx = ctl + 4 * ((e + 1) % 3);
        nx[0] = HEU_NONE;
        nx[1] = best_key(M, 0); /* the largest key of length M: a completed chain has to be strictly longer to beat it */
        ctl[4 * ((e + 2) % 3) + 2] = HEU_NONE;
      }
      continue; /* the next prepare phase only touches set (e+1)%3's horizon word, reset one epoch earlier */
    }
    const int icc_b = (int)(bs >> 32) - 1, v_b = (int)(0xFFFFFFFFu - (unsigned)(bs & 0xffffffffu));
    if (icc_b > M) {
      if ((int)blockIdx.x == (v_b - start) % (int)gridDim.x) {
        const int32_t* src = picks_block + ((size_t)blockIdx.x * 2 + (size_t)saved_sel[blockIdx.x]) * n;
        for (int k = tid; k < icc_b - 1; k += blockDim.x) picks_out[k] = src[k];
      }
      winner = v_b;
      winner_M = M;
      winner_icc = icc_b;
      M = icc_b;
    }
    break;
  }
  if (gtid == 0) {
    result->M = M;
    result->winner = winner;
    result->winner_M = winner_M;
    result->winner_icc = winner_icc;
    result->epochs = epochs;
  }
}

/* (vstar, best) -> (vstar or LLONG_MAX, -best): one MIN all-reduce then yields the lowest v* and the best chain */
__global__ void heu_pack_keys_kernel(const ull* cset, long long* out) {
  const ull vs = cset[0];
  out[0] = vs == HEU_NONE ? LLONG_MAX : (long long)vs;
  out[1] = -(long long)cset[1];
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

/* out[p], p = 1..K: the (p-1)-th smallest u with elim[u] > kstar[p] (see DESIGN.md "scratch-buffer replay").
 * One warp per p scans elim in ascending u.  A lane reads four consecutive entries (one 128-bit load) and four such
 * loads are in flight per trip, so the n/32 dependent load -> ballot steps of the first version (0.5 ms at n = 50 000,
 * one L2 latency each) become n/512 trips; elim is allocated with 64 bytes of slack and entries >= n are masked. */
__global__ void heu_select_kernel(int n, int K, const int32_t* __restrict__ elim, const int32_t* __restrict__ kstar,
                                  int32_t* out) {
  const int p = 1 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p > K) return;
  const int ks = kstar[p];
  int need = p - 1; /* rank to find */
  int found = -1;
  const int4* e4 = reinterpret_cast<const int4*>(elim);
  const int n4 = (n + 3) >> 2; /* int4 words that hold entries < n (the tail word is padded by the allocation) */
  for (int base4 = 0; base4 < n4 && found < 0; base4 += 128) {
    int4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int w = base4 + q * 32 + lane;
      v[q] = w < n4 ? e4[w] : make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (found >= 0) break;
      const int u0 = (base4 + q * 32 + lane) * 4;
      const bool b0 = u0 < n && v[q].x > ks, b1 = u0 + 1 < n && v[q].y > ks, b2 = u0 + 2 < n && v[q].z > ks,
                 b3 = u0 + 3 < n && v[q].w > ks;
      const int c = (int)b0 + (int)b1 + (int)b2 + (int)b3;
      const int total = __reduce_add_sync(0xffffffffu, c);
      if (need < total) {
        /* inclusive prefix over the lanes, then the position inside the owning lane's four entries */
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        const int excl = incl - c;
        int mine = -1;
        if (need >= excl && need < incl) {
          int r = need - excl;
          if (b0) { if (r == 0) mine = u0; --r; }
          if (mine < 0 && b1) { if (r == 0) mine = u0 + 1; --r; }
          if (mine < 0 && b2) { if (r == 0) mine = u0 + 2; --r; }
          if (mine < 0 && b3) { if (r == 0) mine = u0 + 3; }
        }
        found = __reduce_max_sync(0xffffffffu, mine);
      } else {
        need -= total;
      }
    }
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
                     cudaStream_t st, CliqueShard cs, CliqueStats* stats_out) {
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
    cudaFuncSetAttribute(heu_persistent_kernel<HEU_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(heu_persistent_kernel<HEU_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  static const char* wide_env = getenv("RPGO_CLIQUE_WIDE"); /* A/B knob: 0 / 1 forces the block size */
  const bool wide = wide_env ? (wide_env[0] == '1') : (n > HEU_WIDE_N);
  const int threads = wide ? HEU_THREADS_WIDE : HEU_THREADS;
  static const bool trace = getenv("RPGO_CLIQUE_TRACE") != nullptr;
  static const bool host_loop = getenv("RPGO_CLIQUE_HOSTLOOP") != nullptr;
  auto now_ms = []() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  const double t_begin = now_ms();
  ull* ctl = (ull*)s.ctl;
  int start = first < 0 ? 0 : first;
  int M = maxclq0, M_prev = maxclq0;
  int winner = -1, winner_M = 0, winner_icc = 0, epochs = 0;
  if (start >= n) return M;

  /* grid: one block per candidate up to the scratch capacity; small groups take a slice of the machine so that
   * independent searches (batched entry point) run side by side */
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int mine = (n - start + world - 1) / world; /* candidates of this rank (upper bound) */
  int grid = (int)s.rwork_blocks;
  {
    const int want = std::max(64, (mine + 7) / 8);
    if (grid > want) grid = want;
    if (grid > mine) grid = std::max(1, mine);
  }

  /* control block: set 0 ready for epoch 0, horizon word of set 1, statistics, saved_sel */
  {
    ull init[16];
    for (int i = 0; i < 16; ++i) init[i] = 0;
    init[0] = HEU_NONE;
    init[1] = best_key(M, 0); /* the largest key of length M: only strictly longer chains beat it */
    init[2] = HEU_NONE;
    init[4 + 2] = HEU_NONE;
    CUCHECK(cudaMemcpyAsync(ctl, init, sizeof(init), cudaMemcpyHostToDevice, st));
    CUCHECK(cudaMemsetAsync((char*)ctl + CTL_SAVED_SEL_BYTES, 0, sizeof(int32_t) * (size_t)s.rwork_blocks, st));
    CUCHECK(cudaMemsetAsync(s.elim, 0, sizeof(int32_t) * n, st));
  }

  bool done = false;
  bool winner_local = true;
  if (!sharded && !host_loop) {
    /* single rank: all epochs in one cooperative launch */
    const void* pk = wide ? (const void*)heu_persistent_kernel<HEU_THREADS_WIDE> : (const void*)heu_persistent_kernel<HEU_THREADS>;
    int bps = 0;
    if (wide) CUCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, heu_persistent_kernel<HEU_THREADS_WIDE>, HEU_THREADS_WIDE, smem));
    else CUCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, heu_persistent_kernel<HEU_THREADS>, HEU_THREADS, smem));
    if (bps >= 1) {
      int cgrid = std::min(grid, bps * sms);
      const uint32_t* a_bits = bits;
      int64_t a_stride = stride32;
      int a_n = n, a_first = start, a_M = maxclq0;
      const int32_t* a_deg = deg;
      uint32_t* a_degmask = s.degmask;
      ull* a_ctl = ctl;
      int32_t* a_rwork = (int32_t*)s.rwork;
      int32_t* a_elim = s.elim;
      int32_t* a_picks = s.picks;
      void* args[] = {&a_bits, &a_stride, &a_n, &a_deg, &a_degmask, &a_first, &a_M, &a_ctl, &a_rwork, &a_elim, &a_picks};
      CUCHECK(cudaLaunchCooperativeKernel(pk, dim3(cgrid), dim3(threads), args, smem, st));
      *launches += 1;
      HeuResult hr;
      CUCHECK(cudaMemcpyAsync(&hr, ctl + CTL_RESULT, sizeof(hr), cudaMemcpyDeviceToHost, st));
      CUCHECK(cudaStreamSynchronize(st));
      M = hr.M;
      winner = hr.winner;
      winner_M = hr.winner_M;
      winner_icc = hr.winner_icc;
      epochs = hr.epochs;
      done = true;
    }
  }
  /* host-driven epochs: the multi-rank form (one all-reduce of the control words per epoch) and the fallback */
  int winner_owner = rank;
  while (!done && start < n) {
    const double t0 = now_ms();
    if (epochs > 0) {
      ull init[3] = {HEU_NONE, best_key(M, 0), HEU_NONE};
      CUCHECK(cudaMemcpyAsync(ctl, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    heu_prepare_kernel<<<std::min(4 * sms, std::max(1, (n + 127) / 128)), 128, 0, st>>>(bits, stride32, n, deg, M_prev, M, epochs > 0 ? 1 : 0,
                                                                                        s.degmask, ctl, s.elim);
    /* candidates of this rank: v = rank (mod world), starting at or after `start` */
    const int v0 = start + (((rank - start) % world) + world) % world;
    int egrid = (n - v0 + world - 1) / world;
    if (egrid > grid) egrid = grid;
    if (egrid > 0) {
      if (wide)
        heu_round_kernel<HEU_THREADS_WIDE><<<egrid, HEU_THREADS_WIDE, smem, st>>>(bits, stride32, n, deg, s.degmask, v0, world, M, ctl,
                                                                              (int32_t*)s.rwork, s.elim);
      else
        heu_round_kernel<HEU_THREADS><<<egrid, HEU_THREADS, smem, st>>>(bits, stride32, n, deg, s.degmask, v0, world, M, ctl,
                                                                      (int32_t*)s.rwork, s.elim);
      *launches += 1;
    }
    *launches += 1;
    /* epoch decision: lowest v* and best chain over all ranks */
    long long key[2];
    if (sharded && cs.comm) {
      /* on the device: -best so that one MIN all-reduce serves both words */
      heu_pack_keys_kernel<<<1, 1, 0, st>>>(ctl, (long long*)(ctl + 8));
      if (comm_allreduce_i64_device(cs.comm, (long long*)(ctl + 8), 2, false, st) != 0) return -3;
      CUCHECK(cudaMemcpyAsync(key, ctl + 8, sizeof(key), cudaMemcpyDeviceToHost, st));
      CUCHECK(cudaStreamSynchronize(st));
    } else {
      ull h[2];
      CUCHECK(cudaMemcpyAsync(h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
      CUCHECK(cudaStreamSynchronize(st));
      key[0] = h[0] == HEU_NONE ? LLONG_MAX : (long long)h[0];
      key[1] = -(long long)h[1];
      if (sharded && cs.xchg(RPGO_XCHG_MIN_I64, key, 2, 0) != 0) return -3;
    }
    ++epochs;
    if (trace) fprintf(stderr, "[clique] rank %d epoch %d start=%d M=%d grid=%d: %.3f ms\n", rank, epochs, start, M, egrid, now_ms() - t0);
    int which = -1;
    if (key[0] != LLONG_MAX) {
      winner = (int)((ull)key[0] >> 32);
      winner_icc = (int)((ull)key[0] & 0xffffffffu);
      which = 0;
    } else {
      const ull bs = (ull)(-key[1]);
      const int icc_b = (int)(bs >> 32) - 1;
      if (icc_b > M) {
        winner = (int)(0xFFFFFFFFu - (unsigned)(bs & 0xffffffffu));
        winner_icc = icc_b;
        which = 1;
      }
    }
    if (which >= 0) {
      winner_M = M;
      winner_owner = winner % world;
      winner_local = !sharded || winner_owner == rank;
      if (winner_local && winner_icc > 1) {
        const int blk = ((winner - v0) / world) % egrid;
        heu_copy_log_kernel<<<std::min(64, (winner_icc + 255) / 256), 256, 0, st>>>((const int32_t*)s.rwork, ctl, blk, which, n, winner_icc - 1,
                                                                                   s.picks);
        *launches += 1;
      }
    }
    if (which >= 0) {
      M_prev = M;
      M = winner_icc;
    }
    if (which != 0) break; /* no candidate crossed the horizon: the loop ends in this epoch */
    start = winner + 1;
  }
  if (stats_out) {
    ull st2[2] = {0, 0};
    CUCHECK(cudaMemcpyAsync(st2, ctl + CTL_ROW_ANDS, sizeof(st2), cudaMemcpyDeviceToHost, st));
    CUCHECK(cudaStreamSynchronize(st));
    stats_out->row_ands = (long long)st2[0];
    stats_out->chains = (long long)st2[1];
    stats_out->epochs = epochs;
  }
  if (winner < 0) return M; /* no candidate improved on maxclq0 (incremental mode) */
  const double t_replay = now_ms();
  if (trace) fprintf(stderr, "[clique] %d epochs done at %.3f ms\n", epochs, t_replay - t_begin);

  const int K = winner_icc - 1;
  std::vector<int32_t> picks_host(K > 0 ? K : 1);
  if (K > 0 && (!sharded || winner_local)) CUCHECK(cudaMemcpyAsync(picks_host.data(), s.picks, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
  CUCHECK(cudaStreamSynchronize(st));
  if (sharded && K > 0) {
    /* the winning chain ran on rank winner_owner: its pick log goes to everybody */
    if (cs.xchg(RPGO_XCHG_BCAST_I32, picks_host.data(), K, winner_owner) != 0) return -3;
    if (!winner_local) CUCHECK(cudaMemcpyAsync(s.picks, picks_host.data(), sizeof(int32_t) * K, cudaMemcpyHostToDevice, st));
  }
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
