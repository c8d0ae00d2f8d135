/* clique_exact.cu — K5: exact maximum clique by branch and bound over the bitset adjacency (sm_100a).
 *
 * Replaces KimeraRPGO::findMaxClique -> FMC::maxClique / maxCliqueHelper
 * (reference src/utils/GraphUtils.cpp:9-17, include/KimeraRPGO/max_clique_finder/findClique.cpp:31-145).
 * The reference visits root vertices i = n-1 .. 0, restricts candidates to neighbours j < i, always
 * branches on the largest remaining candidate first and replaces its incumbent only by a STRICTLY larger
 * clique.  Its answer is therefore the maximum clique that comes first in that DFS order — the one whose
 * descending id list is lexicographically greatest — returned in ascending id order (ids are pushed while
 * the recursion unwinds, findClique.cpp:71, :134).  Degree / size pruning never removes a maximum clique,
 * so any search that honours the same preference returns the same clique.
 *
 * GPU formulation: roots are partitioned over thread blocks (block b takes roots n-1-b, n-1-b-G, ...),
 * each block runs the DFS iteratively with its candidate sets R_l (bitsets) on a private stack in global
 * memory; set intersection, popcount and highest-bit search are block-wide.  The incumbent is one packed
 * 64-bit word (size << 32 | root) updated with atomicMax: "larger size, then larger root" is exactly the
 * reference's preference between roots, and inside a root the DFS order is the reference's.  A root with a
 * larger index than the incumbent's needs only size >= incumbent (it wins ties), otherwise size > incumbent.
 * Bound: size + |R| (the reference's "old pruning", findClique.cpp:51).
 */
#include <cstdlib>
#include <vector>

#include "kernels.cuh"

namespace rpgo {

static constexpr int EX_THREADS = 128;

__device__ __forceinline__ void ex_block_sum_max(int& s, int& m, int* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  s = __reduce_add_sync(0xffffffffu, s);
  m = __reduce_max_sync(0xffffffffu, m);
  __syncthreads();
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

/* stack layout per block: level l occupies W words at stack + l * W; cnt[l], path[l] in small arrays */
__global__ void __launch_bounds__(EX_THREADS) exact_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                           const int32_t* __restrict__ deg, unsigned long long* incumbent,
                                                           uint32_t* stacks, int depth_max, int32_t* paths,
                                                           int32_t* best_paths, unsigned long long* best_keys, int rank,
                                                           int world) {
  __shared__ int sh[64];
  __shared__ unsigned long long s_inc;
  const int W = (n + 31) / 32;
  const int tid = threadIdx.x;
  uint32_t* stack = stacks + (size_t)blockIdx.x * (size_t)(depth_max + 1) * W;
  int32_t* path = paths + (size_t)blockIdx.x * (depth_max + 2);     /* path[0] = root, path[l] = pick at level l */
  int32_t* best = best_paths + (size_t)blockIdx.x * (depth_max + 2);

  /* roots n-1-rank, n-1-rank-world, ...: this rank's share (world = 1: all of them) */
  for (long long rr = (long long)n - 1 - rank - (long long)blockIdx.x * world; rr >= 0; rr -= (long long)gridDim.x * world) {
    const int root = (int)rr;
    __syncthreads();
    if (tid == 0) s_inc = *(volatile unsigned long long*)incumbent;
    __syncthreads();
    unsigned long long inc = s_inc;
    int inc_size = (int)(inc >> 32), inc_root = (int)(inc & 0xffffffffu);
    int need = (inc == 0ULL) ? 1 : ((root > inc_root) ? inc_size : inc_size + 1);
    if (deg[root] + 1 < need) continue; /* pruning 1 */
    /* level 1 candidate set: neighbours j < root with deg(j) + 1 >= need */
    int cnt = 0, top = -1;
    for (int w = tid; w < W; w += blockDim.x) {
      uint32_t r = bits[(size_t)root * stride32 + w];
      const int base = w * 32;
      if (base + 31 >= root) r &= (base >= root) ? 0u : ((1u << (root - base)) - 1u);
      stack[(size_t)W + w] = r; /* level 1 */
      cnt += __popc(r);
      if (r) top = w;
    }
    ex_block_sum_max(cnt, top, sh);
    if (tid == 0) path[0] = root;
    if (cnt == 0) {
      /* the root alone: helper called with an empty U (findClique.cpp:42-48) */
      if (1 >= need) {
        __syncthreads();
        if (tid == 0) {
          const unsigned long long key = (1ULL << 32) | (unsigned)root;
          const unsigned long long old = atomicMax(incumbent, key);
          if (old < key) {
            best[0] = root;
            best_keys[blockIdx.x] = key;
          }
        }
        __syncthreads();
      }
      continue;
    }
    int level = 1; /* stack[level] = candidates to extend a clique of size `level` */
    /* per-level count / top are recomputed on return (cheap relative to the intersections) */
    while (level >= 1) {
      uint32_t* R = stack + (size_t)level * W;
      /* refresh the incumbent now and then */
      __syncthreads();
      if (tid == 0) s_inc = *(volatile unsigned long long*)incumbent;
      __syncthreads();
      inc = s_inc;
      inc_size = (int)(inc >> 32);
      inc_root = (int)(inc & 0xffffffffu);
      need = (inc == 0ULL) ? 1 : ((root > inc_root) ? inc_size : inc_size + 1);
      /* recompute cnt/top of R (state after pops) */
      cnt = 0;
      top = -1;
      for (int w = tid; w < W; w += blockDim.x) {
        const uint32_t r = R[w];
        cnt += __popc(r);
        if (r) top = w;
      }
      ex_block_sum_max(cnt, top, sh);
      if (cnt == 0 || level + cnt < need) { /* exhausted or bound: backtrack */
        --level;
        continue;
      }
      /* pick the largest candidate, remove it from this level */
      const uint32_t tw = R[top];
      const int b = 31 - __clz(tw);
      const int v = top * 32 + b;
      __syncthreads();
      if (tid == 0) {
        R[top] = tw & ~(1u << b);
        path[level] = v;
      }
      __syncthreads();
      /* child set = R (after removing v; all remaining are < v) ∧ N(v) */
      uint32_t* C = stack + (size_t)(level + 1) * W;
      int ccnt = 0;
      for (int w = tid; w <= top; w += blockDim.x) {
        const uint32_t r = R[w] & bits[(size_t)v * stride32 + w];
        C[w] = r;
        ccnt += __popc(r);
      }
      for (int w = top + 1 + tid; w < W; w += blockDim.x) C[w] = 0u;
      int dummy = -1;
      ex_block_sum_max(ccnt, dummy, sh);
      const int size = level + 1; /* root + picks at levels 1..level */
      if (ccnt == 0) {
        /* leaf: clique of `size` vertices path[0..level] */
        if (size >= need) {
          __syncthreads();
          if (tid == 0) {
            const unsigned long long key = ((unsigned long long)(unsigned)size << 32) | (unsigned)root;
            const unsigned long long old = atomicMax(incumbent, key);
            if (old < key) {
              for (int l = 0; l <= level; ++l) best[l] = path[l];
              best_keys[blockIdx.x] = key;
            }
          }
          __syncthreads();
        }
        /* stay on this level: next candidate */
      } else if (level + 1 <= depth_max - 1) {
        ++level; /* descend */
      }
    }
  }
}

/* ---- warp-level search with a colouring bound and a task frontier (n <= 1024) -----------------------------
 * One word of every candidate set per lane (32 lanes x 32 bits), the whole adjacency in shared memory, no block
 * barriers inside the search.  Work items are the EDGES (i > j) = the first two vertices of a clique, handed out
 * through one global counter in descending (i, j) order (the reference's DFS order), so a hard root is spread over
 * many warps.  The incumbent is one packed word (size << 42 | i << 21 | j): among cliques of equal size the
 * reference keeps the one its DFS meets first, i.e. the lexicographically greatest descending id list; tasks differ
 * in their first two vertices, inside a task the DFS order is the reference's.  A task ahead of the incumbent's in
 * that order only needs to tie.  Bounds: size + |R| (the reference's) and size + (greedy colour classes of R),
 * evaluated with early exit as soon as the colouring can no longer prune. */
static constexpr int EXW_WARPS = 8;

__device__ __forceinline__ int warp_highest(uint32_t w, int& word_lane) {
  const unsigned m = __ballot_sync(0xffffffffu, w != 0u);
  if (!m) return -1;
  word_lane = 31 - __clz(m);
  const uint32_t ww = __shfl_sync(0xffffffffu, w, word_lane);
  return word_lane * 32 + (31 - __clz(ww));
}

__global__ void __launch_bounds__(EXW_WARPS * 32) exact_warp_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n,
                                                                   const int32_t* __restrict__ task_pre, int n_tasks,
                                                                   int rank, int world, unsigned long long* incumbent,
                                                                   unsigned int* counter, uint32_t* stacks, int depth_max,
                                                                   int32_t* paths, int32_t* best_paths,
                                                                   unsigned long long* best_keys) {
  extern __shared__ uint32_t sadj[]; /* n rows x 32 words */
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * EXW_WARPS + wib;
  for (int idx = threadIdx.x; idx < n * 32; idx += blockDim.x) {
    const int v = idx >> 5, w = idx & 31;
    sadj[idx] = (w < (n + 31) / 32) ? bits[(size_t)v * stride32 + w] : 0u;
  }
  __syncthreads();
  uint32_t* stack = stacks + (size_t)gw * (size_t)(depth_max + 2) * 32;
  int32_t* path = paths + (size_t)gw * (depth_max + 2);
  int32_t* best = best_paths + (size_t)gw * (depth_max + 2);
  const unsigned full = 0xffffffffu;

  for (;;) {
    unsigned int k = 0;
    if (lane == 0) k = atomicAdd(counter, 1u);
    k = __shfl_sync(full, k, 0);
    const long long e = (long long)rank + (long long)k * world;
    if (e >= n_tasks) break;
    /* decode task e -> (i, j): task_pre[i] = number of tasks of roots > i (descending order) */
    int lo = 0, hi = n - 1; /* find the largest i with task_pre[i] <= e  (task_pre is non-increasing in i) */
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (task_pre[mid] <= e) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    int t = (int)(e - task_pre[i]); /* t-th highest neighbour below i */
    uint32_t rw = sadj[i * 32 + lane];
    {
      const int base = lane * 32;
      if (base + 31 >= i) rw &= (base >= i) ? 0u : ((1u << (i - base)) - 1u);
    }
    int j = -1;
    {
      /* suffix popcounts over lanes (number of set bits in higher lanes) */
      const int c = __popc(rw);
      int suf = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_down_sync(full, suf, d);
        if (lane + d < 32) suf += o;
      }
      const int above = suf - c; /* bits in lanes > this one */
      const bool here = (t >= above) && (t < above + c);
      const unsigned hm = __ballot_sync(full, here);
      const int hl = __ffs(hm) - 1;
      const uint32_t ww = __shfl_sync(full, rw, hl);
      int tt = t - __shfl_sync(full, above, hl);
      uint32_t q = ww;
      for (int x = 0; x < tt; ++x) q &= ~(1u << (31 - __clz(q)));
      j = hl * 32 + (31 - __clz(q));
    }
    unsigned long long inc = __shfl_sync(full, *(volatile unsigned long long*)incumbent, 0);
    int inc_size = (int)(inc >> 42);
    unsigned long long my_ij = ((unsigned long long)(unsigned)i << 21) | (unsigned)j;
    int need = (inc == 0ULL) ? 1 : ((my_ij > (inc & ((1ULL << 42) - 1))) ? inc_size : inc_size + 1);
    /* candidates: common neighbours below j */
    uint32_t R = rw & sadj[j * 32 + lane];
    {
      const int base = lane * 32;
      if (base + 31 >= j) R &= (base >= j) ? 0u : ((1u << (j - base)) - 1u);
    }
    if (lane == 0) { path[0] = i; path[1] = j; }
    int level = 2; /* current clique size; stack[level] = candidates to extend it */
    stack[(size_t)level * 32 + lane] = R;
    if (__ballot_sync(full, R != 0u) == 0u) {
      if (2 >= need && lane == 0) {
        const unsigned long long key = (2ULL << 42) | my_ij;
        const unsigned long long old = atomicMax(incumbent, key);
        if (old < key) { best[0] = i; best[1] = j; best_keys[gw] = key; }
      }
      __syncwarp();
      continue;
    }
    while (level >= 2) {
      R = stack[(size_t)level * 32 + lane];
      inc = __shfl_sync(full, *(volatile unsigned long long*)incumbent, 0);
      inc_size = (int)(inc >> 42);
      need = (inc == 0ULL) ? 1 : ((my_ij > (inc & ((1ULL << 42) - 1))) ? inc_size : inc_size + 1);
      const int cnt = __reduce_add_sync(full, __popc(R));
      if (cnt == 0 || level + cnt < need) { --level; continue; }
      /* colouring bound: can R still supply need - level mutually adjacent vertices? */
      {
        uint32_t Q = R;
        int colours = 0;
        bool can_prune = true;
        while (__ballot_sync(full, Q != 0u)) {
          ++colours;
          if (level + colours >= need) { can_prune = false; break; }
          uint32_t U = Q;
          int wl;
          int v;
          while ((v = warp_highest(U, wl)) >= 0) {
            if (lane == wl) { const uint32_t bit = 1u << (v & 31); Q &= ~bit; U &= ~bit; }
            U &= ~sadj[v * 32 + lane];
          }
        }
        if (can_prune) { --level; continue; }
      }
      int wl;
      const int v = warp_highest(R, wl);
      if (lane == wl) R &= ~(1u << (v & 31));
      stack[(size_t)level * 32 + lane] = R;
      if (lane == 0) path[level] = v;
      const uint32_t C = R & sadj[v * 32 + lane]; /* everything left in R is < v */
      const int size = level + 1;
      if (__ballot_sync(full, C != 0u) == 0u) {
        if (size >= need) {
          __syncwarp();
          if (lane == 0) {
            const unsigned long long key = ((unsigned long long)(unsigned)size << 42) | my_ij;
            const unsigned long long old = atomicMax(incumbent, key);
            if (old < key) {
              for (int l = 0; l <= level; ++l) best[l] = path[l];
              best_keys[gw] = key;
            }
          }
          __syncwarp();
        }
      } else if (level + 1 <= depth_max) {
        ++level;
        stack[(size_t)level * 32 + lane] = C;
      }
    }
  }
}

__global__ void lowdeg_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n, int32_t* lowdeg) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int W = (n + 31) / 32;
  int c = 0;
  for (int w = lane; w < W; w += 32) {
    uint32_t r = bits[(size_t)i * stride32 + w];
    const int base = w * 32;
    if (base + 31 >= i) r &= (base >= i) ? 0u : ((1u << (i - base)) - 1u);
    c += __popc(r);
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) lowdeg[i] = c;
}

#define EXCHECK(x)                                 \
  do {                                             \
    cudaError_t e_ = (x);                          \
    if (e_ != cudaSuccess) return -(int)e_ - 1000; \
  } while (0)

static int clique_exact_warp(const uint32_t* bits, int64_t stride32, int n, int maxdeg, int32_t* ids_out_host, int64_t* launches,
                             cudaStream_t st, bool sharded, int rank, int world, CliqueShard cs) {
  const int depth_max = maxdeg + 2;
  int32_t* lowdeg = nullptr;
  EXCHECK(cudaMalloc(&lowdeg, sizeof(int32_t) * (size_t)(n + 1)));
  lowdeg_kernel<<<(n + 7) / 8, 256, 0, st>>>(bits, stride32, n, lowdeg);
  std::vector<int32_t> hlow(n), pre(n + 1);
  EXCHECK(cudaMemcpyAsync(hlow.data(), lowdeg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  EXCHECK(cudaStreamSynchronize(st));
  long long acc = 0;
  for (int i = n - 1; i >= 0; --i) { pre[i] = (int32_t)acc; acc += hlow[i]; }
  const long long n_tasks = acc;
  if (n_tasks == 0) { /* no edge at all: the reference keeps its first root, n-1 */
    cudaFree(lowdeg);
    ids_out_host[0] = n - 1;
    return 1;
  }
  /* task_pre[i] = number of tasks of roots > i; the kernel takes the smallest i with task_pre[i] <= e, which is never a
   * root without tasks (for those task_pre[i-1] == task_pre[i]) */
  EXCHECK(cudaMemcpyAsync(lowdeg, pre.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)n * 32 * sizeof(uint32_t);
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(exact_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 32 * 4);
  int per_sm = (int)((200 * 1024) / (smem > 0 ? smem : 1));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  int grid = 148 * per_sm;
  const long long my_tasks = (n_tasks + world - 1) / world;
  if ((long long)grid * EXW_WARPS > my_tasks) grid = (int)((my_tasks + EXW_WARPS - 1) / EXW_WARPS);
  if (grid < 1) grid = 1;
  const size_t warps = (size_t)grid * EXW_WARPS;
  uint32_t* stacks = nullptr;
  int32_t *paths = nullptr, *best = nullptr;
  unsigned long long *inc = nullptr, *keys = nullptr;
  unsigned int* counter = nullptr;
  EXCHECK(cudaMalloc(&stacks, warps * (size_t)(depth_max + 2) * 32 * sizeof(uint32_t)));
  EXCHECK(cudaMalloc(&paths, warps * (size_t)(depth_max + 2) * sizeof(int32_t)));
  EXCHECK(cudaMalloc(&best, warps * (size_t)(depth_max + 2) * sizeof(int32_t)));
  EXCHECK(cudaMalloc(&inc, sizeof(unsigned long long)));
  EXCHECK(cudaMalloc(&counter, sizeof(unsigned int)));
  EXCHECK(cudaMalloc(&keys, warps * sizeof(unsigned long long)));
  EXCHECK(cudaMemsetAsync(inc, 0, sizeof(unsigned long long), st));
  EXCHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  EXCHECK(cudaMemsetAsync(keys, 0, warps * sizeof(unsigned long long), st));
  exact_warp_kernel<<<grid, EXW_WARPS * 32, smem, st>>>(bits, stride32, n, lowdeg, (int)n_tasks, rank, world, inc, counter, stacks,
                                                        depth_max, paths, best, keys);
  *launches += 2;
  unsigned long long hinc = 0;
  EXCHECK(cudaMemcpyAsync(&hinc, inc, sizeof(hinc), cudaMemcpyDeviceToHost, st));
  EXCHECK(cudaStreamSynchronize(st));
  EXCHECK(cudaGetLastError());
  const unsigned long long mine = hinc;
  int xrc = 0, owner = 0;
  if (sharded) {
    long long key = (long long)hinc;
    xrc = cs.xchg(RPGO_XCHG_MAX_I64, &key, 1, 0);
    hinc = (unsigned long long)key;
    long long who = (mine == hinc) ? rank : -1; /* the incumbent's task ran on exactly one rank */
    if (xrc == 0) xrc = cs.xchg(RPGO_XCHG_MAX_I64, &who, 1, 0);
    owner = (int)who;
  }
  const int size = (int)(hinc >> 42);
  int rc = xrc != 0 ? -3 : size;
  if (size > 0 && xrc == 0) {
    if (!sharded || owner == rank) {
      std::vector<unsigned long long> hk(warps);
      EXCHECK(cudaMemcpy(hk.data(), keys, warps * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      size_t b = 0;
      while (b < warps && hk[b] != hinc) ++b;
      if (b == warps) rc = -5;
      else {
        std::vector<int32_t> p(size);
        EXCHECK(cudaMemcpy(p.data(), best + b * (size_t)(depth_max + 2), sizeof(int32_t) * size, cudaMemcpyDeviceToHost));
        for (int l = 0; l < size; ++l) ids_out_host[l] = p[size - 1 - l]; /* ascending ids, as the reference returns them */
      }
    }
    if (sharded && rc >= 0 && cs.xchg(RPGO_XCHG_BCAST_I32, ids_out_host, size, owner) != 0) rc = -3;
  }
  cudaFree(lowdeg);
  cudaFree(stacks);
  cudaFree(paths);
  cudaFree(best);
  cudaFree(inc);
  cudaFree(counter);
  cudaFree(keys);
  return rc;
}

int clique_exact(const uint32_t* bits, int64_t stride32, int n, const int32_t* deg, CliqueScratch s, int32_t* ids_out_host,
                 int64_t* launches, cudaStream_t st, CliqueShard cs) {
  (void)s;
  const bool sharded = cs.active();
  const int world = sharded ? cs.world : 1, rank = sharded ? cs.rank : 0;
  if (n <= 0) return 0;
  const int W = (n + 31) / 32;
  std::vector<int32_t> hdeg(n);
  EXCHECK(cudaMemcpyAsync(hdeg.data(), deg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  EXCHECK(cudaStreamSynchronize(st));
  int maxdeg = 0;
  for (int v = 0; v < n; ++v) maxdeg = hdeg[v] > maxdeg ? hdeg[v] : maxdeg;
  static const bool force_block = getenv("RPGO_EXACT_BLOCK") != nullptr; /* A/B knob: the general block-per-root kernel */
  if (n <= 1024 && !force_block)
    return clique_exact_warp(bits, stride32, n, maxdeg, ids_out_host, launches, st, sharded, rank, world, cs);
  const int depth_max = maxdeg + 2;
  size_t per_block = (size_t)(depth_max + 1) * W * sizeof(uint32_t);
  int grid = 148 * 4;
  const size_t budget = (size_t)8 << 30;
  if ((size_t)grid * per_block > budget) grid = (int)(budget / per_block);
  if (grid < 1) return -4; /* graph too large for the exact search's stack budget */
  if (grid > (n + world - 1) / world) grid = (n + world - 1) / world;
  if (grid < 1) grid = 1;
  uint32_t* stacks = nullptr;
  int32_t *paths = nullptr, *best = nullptr;
  unsigned long long *inc = nullptr, *keys = nullptr;
  EXCHECK(cudaMalloc(&stacks, (size_t)grid * per_block));
  EXCHECK(cudaMalloc(&paths, (size_t)grid * (depth_max + 2) * sizeof(int32_t)));
  EXCHECK(cudaMalloc(&best, (size_t)grid * (depth_max + 2) * sizeof(int32_t)));
  EXCHECK(cudaMalloc(&inc, sizeof(unsigned long long)));
  EXCHECK(cudaMalloc(&keys, (size_t)grid * sizeof(unsigned long long)));
  EXCHECK(cudaMemsetAsync(inc, 0, sizeof(unsigned long long), st));
  EXCHECK(cudaMemsetAsync(keys, 0, (size_t)grid * sizeof(unsigned long long), st));
  exact_kernel<<<grid, EX_THREADS, 0, st>>>(bits, stride32, n, deg, inc, stacks, depth_max, paths, best, keys, rank, world);
  *launches += 1;
  unsigned long long hinc = 0;
  EXCHECK(cudaMemcpyAsync(&hinc, inc, sizeof(hinc), cudaMemcpyDeviceToHost, st));
  EXCHECK(cudaStreamSynchronize(st));
  EXCHECK(cudaGetLastError());
  int xrc = 0;
  const unsigned long long mine = hinc;
  if (sharded) {
    /* incumbent all-reduce: larger size, then larger root — the reference's preference between roots */
    long long key = (long long)hinc;
    xrc = cs.xchg(RPGO_XCHG_MAX_I64, &key, 1, 0);
    hinc = (unsigned long long)key;
  }
  const int size = (int)(hinc >> 32), root = (int)(hinc & 0xffffffffu);
  int rc = xrc != 0 ? -3 : size;
  if (size > 0 && xrc == 0) {
    const int owner = (n - 1 - root) % world;
    if (owner == rank && mine == hinc) {
      const int b = ((n - 1 - root - rank) / world) % grid;
      std::vector<int32_t> p(size);
      EXCHECK(cudaMemcpy(p.data(), best + (size_t)b * (depth_max + 2), sizeof(int32_t) * size, cudaMemcpyDeviceToHost));
      /* the reference returns the clique in ascending id order (ids pushed while unwinding) */
      for (int l = 0; l < size; ++l) ids_out_host[l] = p[size - 1 - l];
    }
    if (sharded && cs.xchg(RPGO_XCHG_BCAST_I32, ids_out_host, size, owner) != 0) rc = -3;
  }
  cudaFree(stacks);
  cudaFree(paths);
  cudaFree(best);
  cudaFree(inc);
  cudaFree(keys);
  return rc;
}

}  // namespace rpgo
