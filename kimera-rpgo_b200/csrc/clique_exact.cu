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
  int32_t* cnts = path + 0;                                         /* (cnt kept in registers per level via recompute) */
  (void)cnts;
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

#define EXCHECK(x)                                 \
  do {                                             \
    cudaError_t e_ = (x);                          \
    if (e_ != cudaSuccess) return -(int)e_ - 1000; \
  } while (0)

int clique_exact(const uint32_t* bits, int64_t stride32, int n, const int32_t* deg, CliqueScratch s, int32_t* ids_out_host,
                 int64_t* launches, cudaStream_t st, CliqueShard cs) {
  (void)s;
  const bool sharded = cs.world > 1 && cs.exchange != nullptr;
  const int world = sharded ? cs.world : 1, rank = sharded ? cs.rank : 0;
  if (n <= 0) return 0;
  const int W = (n + 31) / 32;
  std::vector<int32_t> hdeg(n);
  EXCHECK(cudaMemcpyAsync(hdeg.data(), deg, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  EXCHECK(cudaStreamSynchronize(st));
  int maxdeg = 0;
  for (int v = 0; v < n; ++v) maxdeg = hdeg[v] > maxdeg ? hdeg[v] : maxdeg;
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
    xrc = cs.exchange(cs.user, RPGO_XCHG_MAX_I64, &key, 1, 0);
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
    if (sharded && cs.exchange(cs.user, RPGO_XCHG_BCAST_I32, ids_out_host, size, owner) != 0) rc = -3;
  }
  cudaFree(stacks);
  cudaFree(paths);
  cudaFree(best);
  cudaFree(inc);
  cudaFree(keys);
  return rc;
}

}  // namespace rpgo
