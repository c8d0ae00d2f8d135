/* pcm_tiled.cu — K3, the pairwise consistency kernel, TMA-staged shared-memory variant (sm_100a).
 *
 * Replaces Pcm::incrementAdjMatrix + areLoopsConsistent (reference Pcm.h:670-768) for one
 * ObservationId group.  Work decomposition:
 *   - a gather pass packs, per closure, everything a pair check needs into one record
 *       [ T(key_from) | T(key_to) | T(closure) | prefix ]   (3*ENTRY+2 doubles)
 *     twice: array-of-structs (rows: the OLDER closure i, warp-uniform, read as broadcasts) and
 *     struct-of-arrays in slabs of 32 closures (columns: the NEWER closure j, one per lane, read
 *     conflict-free);
 *   - a thread block (12 warps, one block per SM) owns one slab of 32 columns and a segment of rows.  The column
 *     slab (38 KB) is brought in by ONE bulk TMA copy (cp.async.bulk ... mbarrier::complete_tx) and stays
 *     resident; the row records stream through 2-stage TMA/mbarrier pipelines, one row per warp;
 *   - one thread evaluates one pair with the straight-line pair function (rpgo_pair_v2.cuh: single running
 *     covariance in registers, b_odom_d parked in a per-thread shared-memory scratch column, rare reference
 *     branches deferred to pair_check_exact); the 32 decisions of a warp are packed by __ballot_sync into
 *     the adjacency word (i, j/32);
 *   - pairwise_grouped_kernel (default) splits the block into three phase-shifted groups of four warps so that
 *     FP64-saturated and latency-bound stages of different groups overlap; pairwise_tiled_kernel is the
 *     one-group form kept for A/B measurements.
 * FP64 CUDA-core bound by design (6x6 / 3x3 fp64 with data-dependent branches: no tensor cores).
 */
#include <atomic>
#include <cstdio>
#include <cstdlib>

#define RPGO_MATH_IMPL
#include "kernels.cuh"
#include "rpgo_pair_v2.cuh"

namespace rpgo {

/* PCM records carry full entries (pose, covariance, flags); PcmSimple needs poses and hop counts only, so its records
 * are compact (pose, rotation_info, node: rpgo_pair_v2.cuh::SimpleEntry) — 3.5x less shared memory and L2 traffic */
template <int D, int MODE = MODE_PCM>
struct Rec {
  static constexpr int E = MODE == MODE_PCM ? Dim<D>::ENTRY : SimpleEntry<D>::E;
  static constexpr int N = 3 * E + 2; /* doubles per record (16-byte multiple) */
  static constexpr int OFF_TF = 0, OFF_TB = E, OFF_LC = 2 * E, OFF_PFX = 3 * E;
  static constexpr int SCR = MODE == MODE_PCM ? Dim<D>::ENTRY : 0; /* per-thread scratch doubles (b_odom_d of the PCM chain) */
  /* Column records: one record per LANE, array-of-structs with an ODD pitch in doubles.  Every field then sits at a
   * compile-time offset from the lane's base (no per-load address arithmetic: 550 of 6020 instructions per pair in the
   * strided layout), and 16 lanes x 8 bytes at a pitch of CP doubles hit 16 different bank pairs (CP mod 16 is odd). */
  static constexpr int CP = (N % 2 == 0) ? N + 1 : N;
  static constexpr int SP = (SCR % 2 == 0) ? SCR + 1 : SCR; /* per-thread scratch pitch, odd for the same reason */
};
static_assert(Rec<3, MODE_SIMPLE>::N % 2 == 0 && Rec<2, MODE_SIMPLE>::N % 2 == 0 && Rec<3>::N % 2 == 0 && Rec<2>::N % 2 == 0,
              "bulk copies move 16-byte units");

int tiled_record_doubles(int dim, int mode) {
  if (mode == MODE_PCM) return dim == 3 ? Rec<3>::N : Rec<2>::N;
  return dim == 3 ? Rec<3, MODE_SIMPLE>::N : Rec<2, MODE_SIMPLE>::N;
}

/* ---- gather -------------------------------------------------------------------------------------- */
template <int D, int MODE>
__global__ void gather_records_kernel(GroupView g, const double* __restrict__ traj, int k0, double* aos, double* soa, double* col) {
  constexpr int E = Rec<D, MODE>::E, RN = Rec<D, MODE>::N, TE = Dim<D>::ENTRY, CP = Rec<D, MODE>::CP;
  const int k = k0 + blockIdx.x;
  if (k >= g.n) return;
  for (int f = threadIdx.x; f < RN; f += blockDim.x) {
    double v = 0.0;
    if (f < 3 * E) {
      const int which = f / E, e = f - which * E;
      /* source entry (trajectory / closure tables keep full entries) and field inside it */
      const double* src = which == 0 ? traj + (size_t)g.idx_front[k] * TE : which == 1 ? traj + (size_t)g.idx_back[k] * TE : g.lc + (size_t)k * TE;
      int sf = e;
      if (MODE != MODE_PCM) sf = e < Dim<D>::PS ? e : (e == SimpleEntry<D>::OFF_ROT ? Dim<D>::OFF_ROT : Dim<D>::OFF_NODE);
      v = src[sf];
    } else if (f == 3 * E) {
      v = (double)g.pfx_front[k];
    }
    aos[(size_t)k * RN + f] = v;
    soa[((size_t)(k >> 5) * RN + f) * 32 + (k & 31)] = v;
    col[(size_t)k * CP + f] = v;
  }
}

/* ---- PTX helpers: mbarrier + 1-D bulk TMA ------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ bool row_owned_t(const Shard& sh, int i) {
  if (sh.world <= 1) return true;
  const int64_t c = i / sh.chunk_rows;
  return c == sh.rank || c == 2 * (int64_t)sh.world - 1 - sh.rank;
}

/* shared memory of the grouped kernel: column slab as 32 per-lane records (pitch CP), row stages (pitch N), per-thread
 * scratch records (pitch SP) */
template <int D, int TILE_WARPS, int MODE>
struct GroupedSmem {
  typedef Rec<D, MODE> R;
  static constexpr size_t JT = (size_t)R::CP * 32 * 8;
  static constexpr size_t IT = (size_t)2 * TILE_WARPS * R::N * 8;
  static constexpr size_t SCR = (size_t)R::SP * (R::SCR ? 1 : 0) * TILE_WARPS * 32 * 8;
  static constexpr size_t BYTES = JT + IT + SCR + 256;
  static_assert(JT % 16 == 0 && IT % 16 == 0 && SCR % 8 == 0, "bulk-copy destinations are 16-byte aligned");
};

/* exact general paths for the rare lanes the straight-line code flags (kept out of line: cold code) */
template <int D>
__device__ __noinline__ bool pair_check_exact(const double* Ta, int sa, const double* Tb, int sb, const double* lci, int sli,
                                              const double* Tc, int sc, const double* Td, int sd, const double* lcj, int slj,
                                              double* scr, int ss, const Thresholds* th, double* dist, bool* near) {
  return pair_check_v1<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, scr, ss, *th, dist, near);
}
template <int D>
__device__ __noinline__ bool pair_simple_exact(const double* Ta, int sa, const double* Tb, int sb, const double* lci, int sli,
                                               const double* Tc, int sc, const double* Td, int sd, const double* lcj, int slj,
                                               const Thresholds* th, double* dist, bool* near) {
  return pair_check_simple_exact<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, *th, dist, near);
}

/* one pair through the mode's straight-line function; flagged lanes go through the exact general code */
template <int D, int MODE>
__device__ __forceinline__ bool tile_pair(const double* Ta, int sa, const double* Tb, int sb, const double* lci, int sli,
                                          const double* Tc, int sc, const double* Td, int sd, const double* lcj, int slj,
                                          double* scr, int ss, const Thresholds& th, bool* near) {
  double dist;
  bool bad, ok;
  if (MODE == MODE_PCM) {
    ok = pair_check_v2<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, scr, ss, th, &dist, near, &bad);
    if (bad) ok = pair_check_exact<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, scr, ss, &th, &dist, near);
  } else {
    ok = pair_check_simple_v2<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, th, &dist, near, &bad);
    if (bad) ok = pair_simple_exact<D>(Ta, sa, Tb, sb, lci, sli, Tc, sc, Td, sd, lcj, slj, &th, &dist, near);
  }
  return ok;
}

constexpr int TILE_SEG = 512;   /* rows per work item */

template <int D, int TILE_WARPS, int MODE = MODE_PCM>
struct TiledSmem {
  static constexpr int RN = Rec<D, MODE>::N;
  static constexpr size_t JT = (size_t)RN * 32 * 8;                 /* column slab */
  static constexpr size_t IT = (size_t)2 * TILE_WARPS * RN * 8;     /* 2 stages of row records */
  static constexpr size_t SCR = (size_t)Rec<D, MODE>::SCR * TILE_WARPS * 32 * 8; /* per-thread scratch entry */
  static constexpr size_t BYTES = JT + IT + SCR + 256;
};


template <int D, int TILE_WARPS, int MINB, int PAIRFN>
__global__ void __launch_bounds__(TILE_WARPS * 32, MINB)
    pairwise_tiled_kernel(GroupView g, const double* __restrict__ aos, const double* __restrict__ soa, int j_begin,
                          int cb_begin, Shard sh, Thresholds th, Flagged fl) {
  constexpr int RN = Rec<D>::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Jt = reinterpret_cast<double*>(smem_raw);
  double* It = reinterpret_cast<double*>(smem_raw + TiledSmem<D, TILE_WARPS>::JT);
  double* Scr = reinterpret_cast<double*>(smem_raw + TiledSmem<D, TILE_WARPS>::JT + TiledSmem<D, TILE_WARPS>::IT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TiledSmem<D, TILE_WARPS>::JT + TiledSmem<D, TILE_WARPS>::IT + TiledSmem<D, TILE_WARPS>::SCR);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int cb = cb_begin + blockIdx.y;          /* column slab */
  const int r0 = blockIdx.x * TILE_SEG;
  int r_end = min(g.n, cb * 32 + 31);            /* rows i with some j > i in this slab */
  r_end = min(r_end, r0 + TILE_SEG);
  if (r0 >= r_end) return;
  if (sh.world > 1) {
    /* whole segment inside one foreign chunk: nothing to do */
    const int64_t c0 = r0 / sh.chunk_rows, c1 = (r_end - 1) / sh.chunk_rows;
    if (c0 == c1 && !row_owned_t(sh, r0)) return;
  }

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bars[0], (uint32_t)TiledSmem<D, TILE_WARPS>::JT);
    tma_load_1d(Jt, soa + (size_t)cb * RN * 32, (uint32_t)TiledSmem<D, TILE_WARPS>::JT, &bars[0]);
    const int rows = min(TILE_WARPS, r_end - r0);
    mbar_expect_tx(&bars[1], (uint32_t)(rows * RN * 8));
    tma_load_1d(It, aos + (size_t)r0 * RN, (uint32_t)(rows * RN * 8), &bars[1]);
  }
  mbar_wait(&bars[0], 0);

  const int j = cb * 32 + lane;
  const double* Jl = Jt + lane; /* field f of column j: Jl[f * 32] */
  const uint8_t pc = (uint8_t)Jl[Rec<D>::OFF_PFX * 32];
  double* scr = Scr + tid;      /* field f: scr[f * 128] */

  int it = 0;
  for (int base = r0; base < r_end; base += TILE_WARPS, ++it) {
    const int stage = it & 1;
    if (tid == 0 && base + TILE_WARPS < r_end) {
      const int rows = min(TILE_WARPS, r_end - (base + TILE_WARPS));
      mbar_expect_tx(&bars[1 + (stage ^ 1)], (uint32_t)(rows * RN * 8));
      tma_load_1d(It + (size_t)(stage ^ 1) * TILE_WARPS * RN, aos + (size_t)(base + TILE_WARPS) * RN, (uint32_t)(rows * RN * 8),
                  &bars[1 + (stage ^ 1)]);
    }
    mbar_wait(&bars[1 + stage], (uint32_t)((it >> 1) & 1));
    const int i = base + w;
    if (i < r_end && row_owned_t(sh, i)) {
      const double* Ir = It + ((size_t)stage * TILE_WARPS + w) * RN;
      bool ok = false;
      if (j < g.n && j > i && j >= j_begin) {
        const uint8_t pa = (uint8_t)Ir[Rec<D>::OFF_PFX];
        /* Pcm.h:691-698: if the prefixes of a and c differ, c and d swap (measurement not inverted) */
        const double* Tc = (pa != pc) ? Jl + Rec<D>::OFF_TB * 32 : Jl + Rec<D>::OFF_TF * 32;
        const double* Td = (pa != pc) ? Jl + Rec<D>::OFF_TF * 32 : Jl + Rec<D>::OFF_TB * 32;
        double dist;
        bool near;
        if (PAIRFN == 2) {
          bool bad;
          ok = pair_check_v2<D>(Ir + Rec<D>::OFF_TF, 1, Ir + Rec<D>::OFF_TB, 1, Ir + Rec<D>::OFF_LC, 1, Tc, 32, Td, 32,
                                Jl + Rec<D>::OFF_LC * 32, 32, scr, TILE_WARPS * 32, th, &dist, &near, &bad);
          if (bad)
            ok = pair_check_exact<D>(Ir + Rec<D>::OFF_TF, 1, Ir + Rec<D>::OFF_TB, 1, Ir + Rec<D>::OFF_LC, 1, Tc, 32, Td, 32,
                                     Jl + Rec<D>::OFF_LC * 32, 32, scr, TILE_WARPS * 32, &th, &dist, &near);
        } else {
          ok = pair_check_v1<D>(Ir + Rec<D>::OFF_TF, 1, Ir + Rec<D>::OFF_TB, 1, Ir + Rec<D>::OFF_LC, 1, Tc, 32, Td, 32,
                                Jl + Rec<D>::OFF_LC * 32, 32, scr, TILE_WARPS * 32, th, &dist, &near);
        }
        if (near) {
          const unsigned long long slot = atomicAdd(fl.count, 1ULL);
          if ((int64_t)slot < fl.cap) {
            fl.pairs[2 * slot] = i;
            fl.pairs[2 * slot + 1] = j;
          }
        }
      }
      const unsigned word = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) {
        unsigned keep = 0;
        if (j_begin > cb * 32) keep = (j_begin >= cb * 32 + 32) ? 0xffffffffu : ((1u << (j_begin - cb * 32)) - 1u);
        uint32_t* p = g.bits + (size_t)i * g.stride32 + cb;
        *p = (*p & keep) | word;
      }
    }
    __syncthreads(); /* stage buffer is free for the copy issued at the top of the next iteration */
  }
}

/* ---- phase-staggered variant --------------------------------------------------------------------------
 * Same tiles, records and pair function; the TILE_WARPS warps of the block are split into G groups that each run their
 * own row stream (own 2-stage TMA pipeline, own named barrier) and start a fraction of an iteration apart.  A pair check
 * alternates between stages that saturate the FP64 pipe (H S H^T) and stages that wait on dependency chains (LLT, LU,
 * Logmap); with every warp of a scheduler in the same stage the pipe idles through the latter.  Groups are laid out so
 * that each scheduler (warp % 4) hosts warps of different groups. */
template <int TW, int G>
__device__ __forceinline__ void group_of(int w, int& grp, int& wi) {
  if (G == 1) { grp = 0; wi = w; }
  else if (G == 2) { grp = ((w >> 2) + (w & 3)) & 1; wi = w >> 1; }
  else if (G == TW / 4) { grp = w >> 2; wi = w & 3; }      /* one warp of each group per scheduler */
  else { grp = w % G; wi = w / G; }
}

#ifndef RPGO_K3_LB_THREADS /* measurement builds: declare a larger block than is launched = a lower register cap */
#define RPGO_K3_LB_THREADS(tw) ((tw) * 32)
#endif
template <int D, int MODE, int TILE_WARPS, int G, int SEG, int MINB>
__global__ void __launch_bounds__(RPGO_K3_LB_THREADS(TILE_WARPS), MINB)
    pairwise_grouped_kernel(GroupView g, const double* __restrict__ aos, const double* __restrict__ col, int j_begin,
                            int cb_begin, Shard sh, Thresholds th, Flagged fl) {
  typedef Rec<D, MODE> R;
  typedef GroupedSmem<D, TILE_WARPS, MODE> SM;
  constexpr int RN = R::N, CP = R::CP, SP = R::SP;
  constexpr int WG = TILE_WARPS / G;
  static_assert(WG * G == TILE_WARPS, "groups must divide the block");
  static_assert(1 + 2 * G <= 32, "mbarrier slots");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Jt = reinterpret_cast<double*>(smem_raw);
  double* It = reinterpret_cast<double*>(smem_raw + SM::JT);
  double* Scr = reinterpret_cast<double*>(smem_raw + SM::JT + SM::IT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SM::JT + SM::IT + SM::SCR);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int grp, wi;
  group_of<TILE_WARPS, G>(w, grp, wi);
  const int cb = cb_begin + blockIdx.y;
  int r0 = blockIdx.x * SEG, r_lim = g.n;
  if (sh.world > 1) {
    /* sharded rows: the grid's x dimension enumerates the segments of this rank's two row chunks only (launch_grouped),
     * so no block is launched for foreign rows and no segment straddles a chunk boundary (at 8 ranks the 7/8 of empty
     * blocks and the straddling ones cost 5 % of the kernel) */
    const int per = (int)((sh.chunk_rows + SEG - 1) / SEG);
    const int x = blockIdx.x;
    const int64_t base = (x < per ? (int64_t)sh.rank : 2 * (int64_t)sh.world - 1 - sh.rank) * sh.chunk_rows;
    r0 = (int)min((int64_t)g.n, base + (int64_t)(x < per ? x : x - per) * SEG);
    r_lim = (int)min((int64_t)g.n, base + sh.chunk_rows);
  }
  int r_end = min(r_lim, cb * 32 + 31);
  r_end = min(r_end, r0 + SEG);
  if (r0 >= r_end) return;
  if (tid == 0) {
    for (int b = 0; b < 1 + 2 * G; ++b) mbar_init(&bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bars[0], (uint32_t)SM::JT);
    tma_load_1d(Jt, col + (size_t)cb * CP * 32, (uint32_t)SM::JT, &bars[0]); /* the slab's 32 records are contiguous */
  }
  const bool leader = (wi == 0 && lane == 0);
  double* Ig = It + (size_t)grp * 2 * WG * RN;       /* this group's two stages */
  uint64_t* gb = bars + 1 + 2 * grp;
  const int first = r0 + grp * WG;
  if (leader && first < r_end) {
    const int rows = min(WG, r_end - first);
    mbar_expect_tx(&gb[0], (uint32_t)(rows * RN * 8));
    tma_load_1d(Ig, aos + (size_t)first * RN, (uint32_t)(rows * RN * 8), &gb[0]);
  }
  mbar_wait(&bars[0], 0);

  const int j = cb * 32 + lane;
  const double* Jl = Jt + (size_t)lane * CP;  /* this lane's column record: field f at Jl[f] */
  const uint8_t pc = (uint8_t)Jl[R::OFF_PFX];
  double* scr = Scr + (size_t)tid * SP;       /* this thread's scratch record */

  int it = 0;
  for (int base = first; base < r_end; base += TILE_WARPS, ++it) {
    const int stage = it & 1;
    if (leader && base + TILE_WARPS < r_end) {
      const int rows = min(WG, r_end - (base + TILE_WARPS));
      mbar_expect_tx(&gb[stage ^ 1], (uint32_t)(rows * RN * 8));
      tma_load_1d(Ig + (size_t)(stage ^ 1) * WG * RN, aos + (size_t)(base + TILE_WARPS) * RN, (uint32_t)(rows * RN * 8),
                  &gb[stage ^ 1]);
    }
    mbar_wait(&gb[stage], (uint32_t)((it >> 1) & 1));
    const int i = base + wi;
    if (i < r_end && row_owned_t(sh, i)) {
      const double* Ir = Ig + ((size_t)stage * WG + wi) * RN;
      bool ok = false;
      if (j < g.n && j > i && j >= j_begin) {
        const uint8_t pa = (uint8_t)Ir[R::OFF_PFX];
        /* Pcm.h:691-698: if the prefixes of a and c differ, c and d swap (measurement not inverted) */
        const double* Tc = (pa != pc) ? Jl + R::OFF_TB : Jl + R::OFF_TF;
        const double* Td = (pa != pc) ? Jl + R::OFF_TF : Jl + R::OFF_TB;
        bool near;
        /* the 6x6 chain is called directly: through the tile_pair wrapper ptxas spills 8 more bytes per thread and the 3D
         * kernel loses 2 % (measured, profiles/r2_k3_variants.md) */
        if constexpr (MODE == MODE_PCM) {
          double dist;
          bool bad;
          ok = pair_check_v2<D>(Ir + R::OFF_TF, 1, Ir + R::OFF_TB, 1, Ir + R::OFF_LC, 1, Tc, 1, Td, 1, Jl + R::OFF_LC, 1, scr, 1, th,
                                &dist, &near, &bad);
          if (bad)
            ok = pair_check_exact<D>(Ir + R::OFF_TF, 1, Ir + R::OFF_TB, 1, Ir + R::OFF_LC, 1, Tc, 1, Td, 1, Jl + R::OFF_LC, 1, scr, 1,
                                     &th, &dist, &near);
        } else
          ok = tile_pair<D, MODE>(Ir + R::OFF_TF, 1, Ir + R::OFF_TB, 1, Ir + R::OFF_LC, 1, Tc, 1, Td, 1, Jl + R::OFF_LC, 1, scr, 1, th,
                                  &near);
        if (near) {
          const unsigned long long slot = atomicAdd(fl.count, 1ULL);
          if ((int64_t)slot < fl.cap) {
            fl.pairs[2 * slot] = i;
            fl.pairs[2 * slot + 1] = j;
          }
        }
      }
      const unsigned word = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) {
        unsigned keep = 0;
        if (j_begin > cb * 32) keep = (j_begin >= cb * 32 + 32) ? 0xffffffffu : ((1u << (j_begin - cb * 32)) - 1u);
        uint32_t* p = g.bits + (size_t)i * g.stride32 + cb;
        *p = (*p & keep) | word;
      }
    }
    if (G == 1) __syncthreads();
    else if (WG == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(WG * 32) : "memory");
  }
}

/* ---- column mode: a handful of NEW closures against all older ones (the online case) -----------------------
 * With one new column the tiled kernels keep one lane per warp busy.  Here the roles of the operand layouts are swapped:
 * the lanes of a warp are 32 consecutive OLDER closures i, read straight from their struct-of-arrays slab in global memory
 * (every field load is one coalesced 256-byte line, L2-resident), and the NEW closure j is the warp-uniform operand (its
 * array-of-structs record, staged in shared memory once per block).  Same pair function, same argument roles (i older,
 * j newer), so the decisions are identical; bit (i, j) is set with atomicOr because a row word may receive several new
 * columns from different blocks. */
template <int D, int MODE, int TILE_WARPS>
__global__ void __launch_bounds__(TILE_WARPS * 32, 1)
    pairwise_column_kernel(GroupView g, const double* __restrict__ aos, const double* __restrict__ soa, int j_begin, Shard sh,
                           Thresholds th, Flagged fl) {
  typedef Rec<D, MODE> R;
  constexpr int RN = R::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Jrec = reinterpret_cast<double*>(smem_raw);                 /* the new closure's record (RN doubles) */
  double* Scr = reinterpret_cast<double*>(smem_raw + ((RN * 8 + 127) / 128) * 128);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int j = j_begin + blockIdx.y;
  if (j >= g.n) return;
  for (int f = tid; f < RN; f += blockDim.x) Jrec[f] = aos[(size_t)j * RN + f];
  __syncthreads();
  const int rs = blockIdx.x * TILE_WARPS + w; /* row slab of this warp */
  const int i = rs * 32 + lane;
  if (rs * 32 >= j) return;                   /* whole slab at or beyond the column: nothing below the diagonal here */
  if (i < j && i < g.n && row_owned_t(sh, i)) {
    const double* Il = soa + (size_t)rs * RN * 32 + lane; /* field f of closure i: Il[f * 32] */
    const uint8_t pa = (uint8_t)Il[R::OFF_PFX * 32];
    const uint8_t pc = (uint8_t)Jrec[R::OFF_PFX];
    /* Pcm.h:691-698: if the prefixes of a and c differ, c and d swap (measurement not inverted) */
    const double* Tc = (pa != pc) ? Jrec + R::OFF_TB : Jrec + R::OFF_TF;
    const double* Td = (pa != pc) ? Jrec + R::OFF_TF : Jrec + R::OFF_TB;
    double* scr = Scr + tid;
    bool near;
    const bool ok = tile_pair<D, MODE>(Il + R::OFF_TF * 32, 32, Il + R::OFF_TB * 32, 32, Il + R::OFF_LC * 32, 32, Tc, 1, Td, 1,
                                       Jrec + R::OFF_LC, 1, scr, TILE_WARPS * 32, th, &near);
    if (near) {
      const unsigned long long slot = atomicAdd(fl.count, 1ULL);
      if ((int64_t)slot < fl.cap) {
        fl.pairs[2 * slot] = i;
        fl.pairs[2 * slot + 1] = j;
      }
    }
    /* set or clear: the word may hold a stale bit when a column is recomputed */
    if (ok) atomicOr(g.bits + (size_t)i * g.stride32 + (j >> 5), 1u << (j & 31));
    else atomicAnd(g.bits + (size_t)i * g.stride32 + (j >> 5), ~(1u << (j & 31)));
  }
}

template <int D, int MODE>
static void launch_column(GroupView g, const double* aos, const double* soa, int j_begin, Shard sh, Thresholds th, Flagged fl,
                          cudaStream_t st) {
  constexpr int TW = 12;
  const size_t smem = ((Rec<D, MODE>::N * 8 + 127) / 128) * 128 + (size_t)Rec<D, MODE>::SCR * TW * 32 * 8;
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(pairwise_column_kernel<D, MODE, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int slabs = (g.n + 31) / 32;
  dim3 grid((slabs + TW - 1) / TW, g.n - j_begin);
  pairwise_column_kernel<D, MODE, TW><<<grid, TW * 32, smem, st>>>(g, aos, soa, j_begin, sh, th, fl);
}

int tiled_column_pitch(int dim, int mode) {
  if (mode == MODE_PCM) return dim == 3 ? Rec<3>::CP : Rec<2>::CP;
  return dim == 3 ? Rec<3, MODE_SIMPLE>::CP : Rec<2, MODE_SIMPLE>::CP;
}

void launch_gather_records(int dim, int mode, GroupView g, const double* traj, int k0, double* aos, double* soa, double* col,
                           cudaStream_t st) {
  if (k0 >= g.n) return;
  if (mode == MODE_PCM) {
    if (dim == 3) gather_records_kernel<3, MODE_PCM><<<g.n - k0, 160, 0, st>>>(g, traj, k0, aos, soa, col);
    else gather_records_kernel<2, MODE_PCM><<<g.n - k0, 64, 0, st>>>(g, traj, k0, aos, soa, col);
  } else {
    if (dim == 3) gather_records_kernel<3, MODE_SIMPLE><<<g.n - k0, 64, 0, st>>>(g, traj, k0, aos, soa, col);
    else gather_records_kernel<2, MODE_SIMPLE><<<g.n - k0, 32, 0, st>>>(g, traj, k0, aos, soa, col);
  }
}

/* ---- validation of the straight-line operations against the built-in IEEE ones (debug hook) -------------- */
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__global__ void fastmath_check_kernel(unsigned long long seed, long long per_thread, unsigned long long* mismatches,
                                      unsigned long long* checked) {
  const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long bad_cnt = 0, ok_cnt = 0;
  for (long long it = 0; it < per_thread; ++it) {
    const unsigned long long r1 = mix64(seed + gid * 0x100000001b3ULL + (unsigned long long)it * 2ULL);
    const unsigned long long r2 = mix64(r1);
    /* random sign/mantissa, exponent spread over 2^-300 .. 2^300 (mostly) with occasional extremes */
    const int e1 = (int)((r1 >> 52) % 640) - 320, e2 = (int)((r2 >> 52) % 640) - 320;
    double a = __longlong_as_double((long long)((r1 & 0x800fffffffffffffULL) | ((unsigned long long)(1023 + e1) << 52)));
    double x = __longlong_as_double((long long)((r2 & 0x800fffffffffffffULL) | ((unsigned long long)(1023 + e2) << 52)));
    if ((it & 15) == 0) a = __longlong_as_double((long long)((r1 & 0x8000000000000fffULL) | 0x3ff0000000000000ULL)); /* near 1 */
    if ((it & 31) == 1) x = __longlong_as_double((long long)((r2 | 0x000ffffffffff000ULL) & 0x800fffffffffffffULL | 0x3ff0000000000000ULL)); /* near 2 */
    bool b1 = false, b2 = false, b3 = false;
    const double r = opt_rcp(x, b1);
    const double q = opt_div_by(a, x, r, b2);
    b2 = b2 || b1;
    const double ax = fabs(a);
    const double sq = opt_sqrt(ax, b3);
    if (!b1) { if (r != 1.0 / x) { ++bad_cnt; atomicAdd(mismatches + 2, 1ULL); } else ++ok_cnt; }
    if (!b2) { if (q != a / x) { ++bad_cnt; atomicAdd(mismatches + 3, 1ULL); } else ++ok_cnt; }
    if (!b3) { if (sq != sqrt(ax)) { ++bad_cnt; if (atomicAdd(mismatches + 4, 1ULL) == 0) { mismatches[5] = (unsigned long long)__double_as_longlong(ax); } } else ++ok_cnt; }
    /* fused sqrt + reciprocal (LLT pivots, Logmap): s, 1/s and x/s through the Markstein correction */
    bool b4 = false, b5 = false;
    double rs;
    const double s2 = opt_sqrt_rcp(ax, rs, b4);
    const double q2 = opt_div_by(x, s2, rs, b5);
    if (!b4) {
      const double s_ref = sqrt(ax);
      if (s2 != s_ref || rs != 1.0 / s_ref) { ++bad_cnt; atomicAdd(mismatches + 6, 1ULL); } else ++ok_cnt;
      if (!b5) { if (q2 != x / s_ref) { ++bad_cnt; atomicAdd(mismatches + 7, 1ULL); } else ++ok_cnt; }
    }
  }
  atomicAdd(mismatches, bad_cnt);
  atomicAdd(checked, ok_cnt + bad_cnt);
}

int fastmath_check(long long n, unsigned long long seed, unsigned long long* mismatches, unsigned long long* checked,
                   cudaStream_t st) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return -1;
  cudaMemsetAsync(d, 0, 64, st);
  const int threads = 256, blocks = 148 * 8;
  const long long per_thread = (n + (long long)threads * blocks - 1) / ((long long)threads * blocks);
  fastmath_check_kernel<<<blocks, threads, 0, st>>>(seed, per_thread, d, d + 1);
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaFree(d);
  *mismatches = h[0];
  *checked = h[1];
  if (h[0]) fprintf(stderr, "[fastmath] mismatches: rcp %llu div %llu sqrt %llu (first sqrt operand bits 0x%016llx) sqrt_rcp %llu div-by-sqrt %llu\n", h[2], h[3], h[4], h[5], h[6], h[7]);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <int D, int TW, int MINB, int PAIRFN>
static void launch_variant(GroupView g, const double* aos, const double* soa, int j_begin, int cb_begin, dim3 grid, Shard sh,
                           Thresholds th, Flagged fl, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first())
    cudaFuncSetAttribute(pairwise_tiled_kernel<D, TW, MINB, PAIRFN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)TiledSmem<D, TW>::BYTES);
  pairwise_tiled_kernel<D, TW, MINB, PAIRFN><<<grid, TW * 32, TiledSmem<D, TW>::BYTES, st>>>(g, aos, soa, j_begin, cb_begin, sh, th, fl);
}

template <int D, int MODE, int TW, int G, int SEG, int MINB>
static void launch_grouped(GroupView g, const double* aos, const double* col, int j_begin, int cb_begin, int cb_end, Shard sh,
                           Thresholds th, Flagged fl, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(pairwise_grouped_kernel<D, MODE, TW, G, SEG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)GroupedSmem<D, TW, MODE>::BYTES);
    /* MINB blocks per SM only materialise if the shared-memory carve-out is large enough for all of them */
    cudaFuncSetAttribute(pairwise_grouped_kernel<D, MODE, TW, G, SEG, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  }
  dim3 grid((g.n + SEG - 1) / SEG, cb_end - cb_begin);
  if (sh.world > 1) grid.x = 2 * (unsigned)((sh.chunk_rows + SEG - 1) / SEG); /* segments of the two owned row chunks */
  pairwise_grouped_kernel<D, MODE, TW, G, SEG, MINB><<<grid, TW * 32, GroupedSmem<D, TW, MODE>::BYTES, st>>>(g, aos, col, j_begin,
                                                                                                         cb_begin, sh, th, fl);
}

/* Block shapes per pair function (measured, profiles/r2_*): the 6x6 chain needs 168 registers and 50 doubles of scratch per
 * thread, so one 12-warp block fills an SM; the 3x3 chain and the pose-only PcmSimple chains are small enough for 2-3
 * blocks of 8 warps per SM, which is what hides their dependency latency. */
#ifndef RPGO_2D_MINB
#define RPGO_2D_MINB 4
#endif
#ifndef RPGO_S3_MINB
#define RPGO_S3_MINB 3
#endif
#ifndef RPGO_S2_MINB
#define RPGO_S2_MINB 3
#endif
template <int D, int MODE>
struct TileShape;
#ifndef RPGO_3D_G
#define RPGO_3D_G 3
#define RPGO_3D_SEG 504
#endif
template <> struct TileShape<3, MODE_PCM> { static constexpr int TW = 12, G = RPGO_3D_G, SEG = RPGO_3D_SEG, SEG_SMALL = 48, MINB = 1; };
#ifndef RPGO_2D_TW
#define RPGO_2D_TW 8
#define RPGO_2D_G 2
#define RPGO_2D_SEG 512
#endif
#ifndef RPGO_S3_TW
#define RPGO_S3_TW 8
#define RPGO_S3_G 2
#define RPGO_S3_SEG 512
#endif
template <> struct TileShape<2, MODE_PCM> { static constexpr int TW = RPGO_2D_TW, G = RPGO_2D_G, SEG = RPGO_2D_SEG, SEG_SMALL = 64, MINB = RPGO_2D_MINB; };
template <> struct TileShape<3, MODE_SIMPLE> { static constexpr int TW = RPGO_S3_TW, G = RPGO_S3_G, SEG = RPGO_S3_SEG, SEG_SMALL = 64, MINB = RPGO_S3_MINB; };
template <> struct TileShape<2, MODE_SIMPLE> { static constexpr int TW = 8, G = 2, SEG = 512, SEG_SMALL = 64, MINB = RPGO_S2_MINB; };

template <int D, int MODE>
static void launch_mode(GroupView g, const double* aos, const double* soa, const double* col, int j_begin, Shard sh, Thresholds th,
                        Flagged fl, cudaStream_t st) {
  typedef TileShape<D, MODE> S;
  const int cb_begin = j_begin / 32;
  const int cb_end = (g.n + 31) / 32;
  /* online case: a handful of new closures against a large group -> lanes over the older closures */
  if (g.n - j_begin <= 32 && j_begin >= 32) {
    launch_column<D, MODE>(g, aos, soa, j_begin, sh, th, fl, st);
    return;
  }
  /* few work items (small groups, or a handful of new columns in the online case): short row segments, so that the
   * launch fills the SMs and a block is a few iterations long (latency of a single-closure update) */
  const long long items = (long long)((g.n + S::SEG - 1) / S::SEG) * (cb_end - cb_begin);
  if (items < 4LL * 148 * S::MINB)
    launch_grouped<D, MODE, S::TW, S::G, S::SEG_SMALL, S::MINB>(g, aos, col, j_begin, cb_begin, cb_end, sh, th, fl, st);
  else
    launch_grouped<D, MODE, S::TW, S::G, S::SEG, S::MINB>(g, aos, col, j_begin, cb_begin, cb_end, sh, th, fl, st);
}

void launch_pairwise_tiled(int dim, int mode, GroupView g, const double* aos, const double* soa, const double* col, int j_begin,
                           Shard sh, Thresholds th, Flagged fl, int variant, cudaStream_t st) {
  if (g.n < 2 || j_begin >= g.n) return;
  if (variant != 0 && mode == MODE_PCM) {
    /* cross-check forms for the parity tests: one warp group, straight-line (1) or plain branchy (2) pair function */
    const int cb_begin = j_begin / 32;
    const int cb_end = (g.n + 31) / 32;
    dim3 grid((g.n + TILE_SEG - 1) / TILE_SEG, cb_end - cb_begin);
    if (dim == 3) {
      if (variant == 1) launch_variant<3, 12, 1, 2>(g, aos, soa, j_begin, cb_begin, grid, sh, th, fl, st);
      else launch_variant<3, 12, 1, 1>(g, aos, soa, j_begin, cb_begin, grid, sh, th, fl, st);
    } else {
      if (variant == 1) launch_variant<2, 12, 1, 2>(g, aos, soa, j_begin, cb_begin, grid, sh, th, fl, st);
      else launch_variant<2, 12, 1, 1>(g, aos, soa, j_begin, cb_begin, grid, sh, th, fl, st);
    }
    return;
  }
  if (mode == MODE_PCM) {
    if (dim == 3) launch_mode<3, MODE_PCM>(g, aos, soa, col, j_begin, sh, th, fl, st);
    else launch_mode<2, MODE_PCM>(g, aos, soa, col, j_begin, sh, th, fl, st);
  } else {
    if (dim == 3) launch_mode<3, MODE_SIMPLE>(g, aos, soa, col, j_begin, sh, th, fl, st);
    else launch_mode<2, MODE_SIMPLE>(g, aos, soa, col, j_begin, sh, th, fl, st);
  }
}

}  // namespace rpgo
