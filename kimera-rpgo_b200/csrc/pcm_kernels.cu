/* pcm_kernels.cu — sm_100a kernels for the PCM consistency stage.
 *
 *   K1  traj_fold / traj_scan   cumulative pose + covariance along each robot's odometry chain
 *                               (replaces Pcm::updateOdom's left fold, reference Pcm.h:545-556)
 *   K2  lc_prepare              PoseWithCovariance/PoseWithNode ctor from the factor + odometry
 *                               consistency check (Pcm.h:604-629, GeometryUtils.h:91-115)
 *   K3  pairwise_direct         one thread per (older i, newer j) closure pair: areLoopsConsistent
 *                               (Pcm.h:670-718) -> warp-ballot packed adjacency words
 *       mirror / degree / clear_last   bitset maintenance (symmetric fill, popcount degrees,
 *                               removeLastLoopClosure's shrink Pcm.h:320-323)
 * The tiled TMA variant of K3 lives in pcm_tiled.cu.
 * Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (the numerical contract forbids
 * implicit contraction; all fused operations are explicit fma() calls in rpgo_math.cuh).
 */
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace rpgo {

#define RPGO_DISPATCH(dim, mode, CALL)                         \
  do {                                                         \
    if ((dim) == 3 && (mode) == MODE_PCM) { CALL(3, MODE_PCM); }       \
    else if ((dim) == 3) { CALL(3, MODE_SIMPLE); }             \
    else if ((mode) == MODE_PCM) { CALL(2, MODE_PCM); }        \
    else { CALL(2, MODE_SIMPLE); }                             \
  } while (0)

/* ------------------------------------------------------------------------------------------------
 * K1: exact left fold, one thread per chain.  Sequential by definition (the reference's rounding is
 * that of a strict left fold); chains (robots) run in parallel.
 * ---------------------------------------------------------------------------------------------- */
template <int D, int MODE>
__global__ void traj_fold_kernel(int n_chains, const FoldChain* __restrict__ chains, const int32_t* __restrict__ out_idx,
                                 const double* __restrict__ dpose, const double* __restrict__ dcov, double* entries) {
  constexpr int E = Dim<D>::ENTRY, PS = Dim<D>::PS, NN = Dim<D>::N * Dim<D>::N;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const FoldChain ch = chains[c];
  PoseT<D, MODE> cur, delta, nxt;
  load_entry<D, MODE>(entries + (size_t)ch.start_idx * E, 1, cur);
  for (int s = 0; s < ch.n_steps; ++s) {
    const int k = ch.first_step + s;
    from_factor<D, MODE>(dpose + (size_t)k * PS, dcov + (size_t)k * NN, delta);
    pt_compose<D, MODE>(cur, delta, nxt);
    store_entry<D, MODE>(entries + (size_t)out_idx[k] * E, 1, nxt);
    cur = nxt;
  }
}


/* K1, batched exact fold (MODE_PCM; round 1's kernel, kept as the A/B reference: RPGO_FOLD_V2=1).  Same arithmetic, same
 * order as the one-thread fold above (bit-identical), but
 * everything that does not depend on the running value is taken off the sequential path:
 *   - the factors of the next 32 steps are copied global -> shared asynchronously (cp.async) while the current 32
 *     steps are folded, so the chain never waits on HBM;
 *   - a parallel "prep" phase (one thread per step) builds Ad(delta^-1), the NaN-masked delta covariance and the
 *     rotation_info flag of each step;
 *   - warp 0 folds the covariance with one or two elements per lane (18 lanes in 3D, 9 in 2D: the sequential path
 *     is ~25 FP64 operations and two shuffle hops per step), warp 1 folds the pose: two independent chains. */
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int D>
__global__ void __launch_bounds__(64) traj_fold_batched_kernel(int n_chains, const FoldChain* __restrict__ chains,
                                                               const int32_t* __restrict__ out_idx,
                                                               const double* __restrict__ dpose,
                                                               const double* __restrict__ dcov, double* entries) {
  constexpr int E = Dim<D>::ENTRY, PS = Dim<D>::PS, N = Dim<D>::N, NN = N * N, OC = Dim<D>::OFF_COV, RD = Dim<D>::RD,
                TD = Dim<D>::TD, HS = (D == 3 ? 18 : 9), B = 32;
  __shared__ __align__(16) double raw_pose[B * PS];
  __shared__ __align__(16) double raw_cov[B * NN];
  __shared__ __align__(16) double cH[2][B * HS];
  __shared__ __align__(16) double cDc[2][B * NN];
  __shared__ __align__(16) double cDl[2][B * PS];
  __shared__ int cOut[2][B];
  __shared__ int cRot[2][B];
  const int c = blockIdx.x;
  if (c >= n_chains) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const FoldChain ch = chains[c];
  if (ch.n_steps <= 0) return;
  const int nb = (ch.n_steps + B - 1) / B;
  /* warp 0: lane (r, cg) owns the covariance elements (r, cg) and, in 3D, (r, cg + 3): 3N lanes are active */
  constexpr int CPL = (D == 3 ? 2 : 1);
  const int ll = lane < 3 * N ? lane : 0;
  const int r = ll / 3, cg = ll % 3;

  auto issue_raw = [&](int b) {
    const size_t k0 = (size_t)ch.first_step + (size_t)b * B;
    const int cnt = min(B, ch.n_steps - b * B);
    for (int i = tid; i < cnt * PS; i += 64) cp_async8(&raw_pose[i], dpose + k0 * PS + i);
    for (int i = tid; i < cnt * NN; i += 64) cp_async8(&raw_cov[i], dcov + k0 * NN + i);
  };
  auto prep = [&](int b) {
    const int cnt = min(B, ch.n_steps - b * B);
    const int st = b & 1;
    if (tid < cnt) {
      Pose<D> Dl;
#pragma unroll
      for (int i = 0; i < PS; ++i) Dl.m[i] = raw_pose[tid * PS + i];
      /* from_factor: NaN rotation covariance => keep only the translation block (GeometryUtils.h:98-113) */
      double tr = raw_cov[tid * NN];
#pragma unroll
      for (int i = 1; i < RD; ++i) tr = tr + raw_cov[tid * NN + i * N + i];
      const bool drot = !(tr != tr);
#pragma unroll
      for (int rr = 0; rr < N; ++rr)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double v = raw_cov[tid * NN + rr * N + j];
          if (!drot) v = (rr >= RD && j >= RD && rr < RD + TD && j < RD + TD) ? v : 0.0;
          cDc[st][tid * NN + rr * N + j] = v;
        }
      const Adj<D> H = adjoint<D>(inverse<D>(Dl));
#pragma unroll
      for (int i = 0; i < HS; ++i) cH[st][tid * HS + i] = H.h[i];
#pragma unroll
      for (int i = 0; i < PS; ++i) cDl[st][tid * PS + i] = Dl.m[i];
      cRot[st][tid] = drot ? 1 : 0;
      cOut[st][tid] = out_idx[(size_t)ch.first_step + (size_t)b * B + tid];
    }
  };

  /* running state */
  Pose<D> P;
  double S[CPL];
  bool rot = true;
  if (warp == 0) {
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) S[cc] = entries[(size_t)ch.start_idx * E + OC + r * N + cg + 3 * cc];
  } else {
    load_pose<D>(entries + (size_t)ch.start_idx * E, 1, P);
    rot = entries[(size_t)ch.start_idx * E + Dim<D>::OFF_ROT] != 0.0;
  }

  issue_raw(0);
  cp_async_wait_all();
  __syncthreads();
  prep(0);
  __syncthreads();
  for (int b = 0; b < nb; ++b) {
    const int st = b & 1;
    const int cnt = min(B, ch.n_steps - b * B);
    if (b + 1 < nb) issue_raw(b + 1); /* in flight during the chain phase */
    if (warp == 0) {
      for (int i = 0; i < cnt; ++i) {
        const double* Hh = cH[st] + i * HS;
        const double* Dc = cDc[st] + i * NN + r * N + cg;
        /* row r of H: columns 0..2 (hr) and, for 3D rows >= 3, columns 3..5 (ar) */
        const double* hr = (D == 3) ? Hh + (r < 3 ? r * 3 : 9 + (r - 3) * 3) : Hh + r * 3;
        const double h0 = hr[0], h1 = hr[1], h2 = hr[2];
        /* t[cc] = (H S)(r, cg + 3 cc), k-order; S(k, j) lives in lane (k, cg) */
        double t[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const double s0 = __shfl_sync(0xffffffffu, S[cc], 0 * 3 + cg);
          const double s1 = __shfl_sync(0xffffffffu, S[cc], 1 * 3 + cg);
          const double s2 = __shfl_sync(0xffffffffu, S[cc], 2 * 3 + cg);
          t[cc] = dot3(h0, h1, h2, s0, s1, s2);
        }
        if (D == 3) {
          const double* ar = Hh + (r < 3 ? r : r - 3) * 3;
          const double a0 = ar[0], a1 = ar[1], a2 = ar[2];
#pragma unroll
          for (int cc = 0; cc < CPL; ++cc) {
            const double s3 = __shfl_sync(0xffffffffu, S[cc], 3 * 3 + cg);
            const double s4 = __shfl_sync(0xffffffffu, S[cc], 4 * 3 + cg);
            const double s5 = __shfl_sync(0xffffffffu, S[cc], 5 * 3 + cg);
            if (r >= 3) {
              t[cc] = fma(a0, s3, t[cc]);
              t[cc] = fma(a1, s4, t[cc]);
              t[cc] = fma(a2, s5, t[cc]);
            }
          }
        }
        /* the full row r of t: element j = g + 3 cc sits in lane (r, g) */
        double tf[N];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
          for (int g = 0; g < 3; ++g) tf[g + 3 * cc] = __shfl_sync(0xffffffffu, t[cc], r * 3 + g);
        /* out(r, j) = t(r, :) . H(j, :) + Dc(r, j) */
        const double* Aj = Hh + cg * 3;
        S[0] = dot3(tf[0], tf[1], tf[2], Aj[0], Aj[1], Aj[2]) + Dc[0];
        if (D == 3) {
          const double* Bj = Hh + 9 + cg * 3;
          double acc = dot3(tf[0], tf[1], tf[2], Bj[0], Bj[1], Bj[2]);
          acc = fma(tf[3 % N], Aj[0], acc);
          acc = fma(tf[4 % N], Aj[1], acc);
          acc = fma(tf[5 % N], Aj[2], acc);
          S[CPL - 1] = acc + Dc[3 % N];
        }
        if (lane < 3 * N) {
          double* o = entries + (size_t)cOut[st][i] * E + OC + r * N + cg;
#pragma unroll
          for (int cc = 0; cc < CPL; ++cc) o[3 * cc] = S[cc];
        }
      }
    } else {
      for (int i = 0; i < cnt; ++i) {
        Pose<D> Dl;
#pragma unroll
        for (int q = 0; q < PS; ++q) Dl.m[q] = cDl[st][i * PS + q];
        P = compose<D>(P, Dl);
        rot = rot && (cRot[st][i] != 0);
        if (lane == 0) {
          double* o = entries + (size_t)cOut[st][i] * E;
#pragma unroll
          for (int q = 0; q < PS; ++q) o[q] = P.m[q];
          o[Dim<D>::OFF_ROT] = rot ? 1.0 : 0.0;
          o[Dim<D>::OFF_NODE] = 0.0;
        }
      }
    }
    if (b + 1 < nb) {
      cp_async_wait_all();
      __syncthreads();
      prep(b + 1);
    }
    __syncthreads();
  }
}


/* K1, pipelined exact fold (MODE_PCM).  Same arithmetic, same order as traj_fold_batched_kernel (bit-identical); the
 * sequential path of a step is reduced to the dependent FP64 chain plus the exchange of the running covariance:
 *   - three warps: warp 0 folds the covariance, warp 1 the pose, warp 2 is the producer: it copies the factors of block
 *     b+2 global -> shared (cp.async, three raw slots) and builds Ad(delta^-1) + the rotation_info flag of block b+1 while
 *     the chain warps fold block b (the batched kernel prepared a block between two chain phases);
 *   - the chain warps read the delta covariance / pose / output slot straight from the raw slot (the NaN mask of
 *     from_factor is one select per element) and fetch the operands of step i+1 before the dependent part of step i;
 *   - the two products of H S H^T exchange the covariance through two shared-memory hops: S column-major (a lane reads
 *     its two columns with 6 LDS.128), H S row-major.  Measured alternatives (profiles/r2_k3_variants.md): shuffles
 *     (18 + 18 SHFL per step) and a single hop where every lane reads the whole S and recomputes its row of H S -- both
 *     within 2 % of this form. */
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait_le() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) { /* a value the compiler may not re-derive inside a loop */
  uint32_t o;
  asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(v));
  return o;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds_f64x2(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void stg_f64x2(double* p, double x, double y) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double sel_f64(bool c, double x, double y) {
  double o;
  asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(o) : "d"(x), "d"(y), "r"((int)c));
  return o;
}

template <int D>
__global__ void __launch_bounds__(96) traj_fold_pipelined_kernel(int n_chains, const FoldChain* __restrict__ chains,
                                                                 const int32_t* __restrict__ out_idx,
                                                                 const double* __restrict__ dpose,
                                                                 const double* __restrict__ dcov, double* entries) {
  constexpr int E = Dim<D>::ENTRY, PS = Dim<D>::PS, N = Dim<D>::N, NN = N * N, OC = Dim<D>::OFF_COV, RD = Dim<D>::RD,
                TD = Dim<D>::TD, HS = (D == 3 ? 18 : 9), B = 32, CPL = (D == 3 ? 2 : 1);
  __shared__ __align__(16) double raw_pose[3][B * PS];
  __shared__ __align__(16) double raw_cov[3][B * NN];
  __shared__ __align__(16) double cH[2][B * HS];
  __shared__ __align__(16) double xs[NN];
  __shared__ __align__(16) double xt[NN];
  __shared__ int raw_out[3][B];
  __shared__ int cRot[2][B];
  const int c = blockIdx.x;
  if (c >= n_chains) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const FoldChain ch = chains[c];
  if (ch.n_steps <= 0) return;
  const int nb = (ch.n_steps + B - 1) / B;

  auto issue_raw = [&](int b) { /* producer warp */
    if (b < nb) {
      const int slot = b % 3;
      const size_t k0 = (size_t)ch.first_step + (size_t)b * B;
      const int cnt = min(B, ch.n_steps - b * B);
      for (int i = lane; i < cnt * PS; i += 32) cp_async8(&raw_pose[slot][i], dpose + k0 * PS + i);
      for (int i = lane; i < cnt * NN; i += 32) cp_async8(&raw_cov[slot][i], dcov + k0 * NN + i);
      if (lane < cnt) cp_async4(&raw_out[slot][lane], out_idx + k0 + lane);
    }
    cp_async_commit(); /* one group per block index, empty past the end: keeps the wait counts uniform */
  };
  auto prep = [&](int b) { /* producer warp: everything of block b that does not depend on the running value */
    if (b >= nb) return;
    const int slot = b % 3, st = b & 1;
    const int cnt = min(B, ch.n_steps - b * B);
    if (lane < cnt) {
      Pose<D> Dl;
#pragma unroll
      for (int i = 0; i < PS; ++i) Dl.m[i] = raw_pose[slot][lane * PS + i];
      double tr = raw_cov[slot][lane * NN];
#pragma unroll
      for (int i = 1; i < RD; ++i) tr = tr + raw_cov[slot][lane * NN + i * N + i];
      cRot[st][lane] = (tr != tr) ? 0 : 1;
      const Adj<D> H = adjoint<D>(inverse<D>(Dl));
#pragma unroll
      for (int i = 0; i < HS; ++i) cH[st][lane * HS + i] = H.h[i];
    }
  };

  if (warp == 2) {
    issue_raw(0);
    issue_raw(1);
    cp_async_wait_le<1>();
    __syncwarp();
    prep(0);
  }
  __syncthreads();

  if (warp == 2) {
    for (int b = 0; b < nb; ++b) {
      issue_raw(b + 2);      /* slot (b+2)%3 was read by the chain warps in iteration b-1 */
      cp_async_wait_le<1>(); /* block b+1 has landed */
      __syncwarp();
      prep(b + 1);
      __syncthreads();
    }
    return;
  }
  /* The chain warps address shared memory through explicit 32-bit window addresses held in registers (the compiler
   * otherwise re-derives them from SR_CgaCtaId inside the loop, a 25-cycle S2UR on the sequential path), keep their loop
   * bodies free of branches (operand fetches are clamped instead of guarded; lanes without an element of their own shadow
   * lane 0 and store the same values to the same addresses), and the loops are unrolled by two so that the operands
   * fetched for step i+1 need no register moves. */
  const uint32_t a_pose = opaque_u32((uint32_t)__cvta_generic_to_shared(&raw_pose[0][0]));
  const uint32_t a_cov = opaque_u32((uint32_t)__cvta_generic_to_shared(&raw_cov[0][0]));
  const uint32_t a_H = opaque_u32((uint32_t)__cvta_generic_to_shared(&cH[0][0]));
  const uint32_t a_out = opaque_u32((uint32_t)__cvta_generic_to_shared(&raw_out[0][0]));
  const uint32_t a_rot = opaque_u32((uint32_t)__cvta_generic_to_shared(&cRot[0][0]));
  if (warp == 1) {
    /* pose chain: every lane composes the same pose and stores it (identical values to identical addresses) */
    Pose<D> P;
    load_pose<D>(entries + (size_t)ch.start_idx * E, 1, P);
    int rot = entries[(size_t)ch.start_idx * E + Dim<D>::OFF_ROT] != 0.0 ? 1 : 0;
    for (int b = 0; b < nb; ++b) {
      const int slot = b % 3, st = b & 1;
      const int cnt = min(B, ch.n_steps - b * B);
      const uint32_t bp = a_pose + slot * (B * PS * 8), bo = a_out + slot * (B * 4), br = a_rot + st * (B * 4);
      Pose<D> Dn;
      int rn, on;
      auto fetch = [&](int i) {
#pragma unroll
        for (int q = 0; q < PS; q += 2) lds_f64x2(bp + (i * PS + q) * 8, Dn.m[q], Dn.m[q + 1]);
        rn = lds_s32(br + i * 4);
        on = lds_s32(bo + i * 4);
      };
      fetch(0);
#pragma unroll 2
      for (int i = 0; i < cnt; ++i) {
        const Pose<D> Dl = Dn;
        const int rc = rn;
        double* o = entries + (size_t)on * E;
        fetch(min(i + 1, cnt - 1));
        P = compose<D>(P, Dl);
        rot &= rc;
#pragma unroll
        for (int q = 0; q < PS; q += 2) stg_f64x2(o + q, P.m[q], P.m[q + 1]);
        if ((Dim<D>::OFF_ROT & 1) == 0 && Dim<D>::OFF_NODE == Dim<D>::OFF_ROT + 1) {
          stg_f64x2(o + Dim<D>::OFF_ROT, rot ? 1.0 : 0.0, 0.0);
        } else {
          o[Dim<D>::OFF_ROT] = rot ? 1.0 : 0.0;
          o[Dim<D>::OFF_NODE] = 0.0;
        }
      }
      __syncthreads();
    }
    return;
  }
  {
    /* covariance chain: lane (r, cg) owns the elements (r, cg) and, in 3D, (r, cg + 3) */
    const int ll = opaque_u32(lane < 3 * N ? lane : 0);
    const int r = ll / 3, cg = ll % 3;
    double S[CPL];
    int keep[CPL]; /* element survives a NaN rotation covariance (translation block only) */
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int j = cg + 3 * cc;
      S[cc] = entries[(size_t)ch.start_idx * E + OC + r * N + j];
      keep[cc] = (r >= RD && j >= RD && r < RD + TD && j < RD + TD) ? 1 : 0;
    }
    const int hro = (D == 3) ? (r < 3 ? r * 3 : 9 + (r - 3) * 3) : r * 3; /* row r of H, columns 0..2 */
    const int aro = (r < 3 ? r : r - 3) * 3;                              /* 3D rows >= 3: columns 3..5 */
    const bool lower = (D == 3) && r >= 3;
    const uint32_t a_xs = opaque_u32((uint32_t)__cvta_generic_to_shared(&xs[0]));
    const uint32_t a_xt = opaque_u32((uint32_t)__cvta_generic_to_shared(&xt[0]));
    const uint32_t w_xs = a_xs + (cg * N + r) * 8; /* column-major S: element (r, cg); (r, cg + 3) is 3 N doubles further */
    const uint32_t r_xs = a_xs + cg * N * 8;
    const uint32_t w_xt = a_xt + (r * N + cg) * 8; /* row-major H S */
    const uint32_t r_xt = a_xt + r * N * 8;
    double* const obase = entries + OC + r * N + cg;
    for (int b = 0; b < nb; ++b) {
      const int slot = b % 3, st = b & 1;
      const int cnt = min(B, ch.n_steps - b * B);
      const uint32_t bH = a_H + st * (B * HS * 8), bc = a_cov + slot * (B * NN * 8) + (r * N + cg) * 8, bo = a_out + slot * (B * 4),
                     br = a_rot + st * (B * 4);
      double nh[3], na[3], nA[3], nB[3], ndc[CPL];
      double* no;
      auto fetch = [&](int i) {
        const uint32_t Hh = bH + i * (HS * 8);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          nh[q] = lds_f64(Hh + (hro + q) * 8);
          nA[q] = lds_f64(Hh + (cg * 3 + q) * 8);
          if (D == 3) { na[q] = lds_f64(Hh + (aro + q) * 8); nB[q] = lds_f64(Hh + (9 + cg * 3 + q) * 8); }
        }
        const int drot = lds_s32(br + i * 4);
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const double v = lds_f64(bc + (i * NN + 3 * cc) * 8);
          ndc[cc] = sel_f64((drot | keep[cc]) != 0, v, 0.0);
        }
        no = obase + (size_t)lds_s32(bo + i * 4) * E;
      };
      fetch(0);
#pragma unroll 2
      for (int i = 0; i < cnt; ++i) {
        double h[3], a[3], A[3], Bq[3], dc[CPL];
#pragma unroll
        for (int q = 0; q < 3; ++q) { h[q] = nh[q]; A[q] = nA[q]; if (D == 3) { a[q] = na[q]; Bq[q] = nB[q]; } }
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) dc[cc] = ndc[cc];
        double* o = no;
        double tf[N];
        double t[CPL];
        /* hop 1: S column-major; a lane reads its own columns */
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) sts_f64(w_xs + 3 * cc * N * 8, S[cc]);
        __syncwarp();
        double col[CPL][N + 1];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          if (D == 3) {
#pragma unroll
            for (int q = 0; q < N; q += 2) lds_f64x2(r_xs + (3 * cc * N + q) * 8, col[cc][q], col[cc][q + 1]);
          } else {
#pragma unroll
            for (int q = 0; q < N; ++q) col[cc][q] = lds_f64(r_xs + q * 8);
          }
        }
        fetch(min(i + 1, cnt - 1));
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          double tt = dot3(h[0], h[1], h[2], col[cc][0], col[cc][1], col[cc][2]);
          if (D == 3) {
            double t6 = fma(a[0], col[cc][3 % N], tt);
            t6 = fma(a[1], col[cc][4 % N], t6);
            t6 = fma(a[2], col[cc][5 % N], t6);
            tt = sel_f64(lower, t6, tt);
          }
          t[cc] = tt;
        }
        /* hop 2: H S row-major; a lane reads its own row */
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) sts_f64(w_xt + 3 * cc * 8, t[cc]);
        __syncwarp();
        if (D == 3) {
#pragma unroll
          for (int j = 0; j < N; j += 2) lds_f64x2(r_xt + j * 8, tf[j], tf[(j + 1) % N]);
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) tf[j] = lds_f64(r_xt + j * 8);
        }
        /* out(r, j) = (H S)(r, :) . H(j, :) + Dc(r, j) */
        S[0] = dot3(tf[0], tf[1], tf[2], A[0], A[1], A[2]) + dc[0];
        if (D == 3) {
          double acc = dot3(tf[0], tf[1], tf[2], Bq[0], Bq[1], Bq[2]);
          acc = fma(tf[3 % N], A[0], acc);
          acc = fma(tf[4 % N], A[1], acc);
          acc = fma(tf[5 % N], A[2], acc);
          S[CPL - 1] = acc + dc[CPL - 1];
        }
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) o[3 * cc] = S[cc];
      }
      __syncthreads();
    }
  }
}

void launch_traj_fold(int dim, int mode, int n_chains, const FoldChain* chains, const int32_t* out_idx,
                      const double* delta_pose, const double* delta_cov, double* entries, cudaStream_t st) {
  if (n_chains <= 0) return;
  /* one chain per block so that independent robots land on different SMs */
  const int blocks = n_chains;
  if (mode == MODE_PCM) {
    {
      static const bool v2 = getenv("RPGO_FOLD_V2") != nullptr; /* A/B knob: the two-warp batched kernel */
      if (v2) {
        if (dim == 3) traj_fold_batched_kernel<3><<<blocks, 64, 0, st>>>(n_chains, chains, out_idx, delta_pose, delta_cov, entries);
        else traj_fold_batched_kernel<2><<<blocks, 64, 0, st>>>(n_chains, chains, out_idx, delta_pose, delta_cov, entries);
      } else {
        if (dim == 3) traj_fold_pipelined_kernel<3><<<blocks, 96, 0, st>>>(n_chains, chains, out_idx, delta_pose, delta_cov, entries);
        else traj_fold_pipelined_kernel<2><<<blocks, 96, 0, st>>>(n_chains, chains, out_idx, delta_pose, delta_cov, entries);
      }
    }
    return;
  }
#define CALL(D, M) traj_fold_kernel<D, M><<<blocks, 1, 0, st>>>(n_chains, chains, out_idx, delta_pose, delta_cov, entries)
  RPGO_DISPATCH(dim, mode, CALL);
#undef CALL
}

/* ------------------------------------------------------------------------------------------------
 * K1 (scan): three-phase chunked prefix scan over the associative operator
 *   (Ta,Sa) o (Tb,Sb) = (Ta Tb, Ad(Tb^-1) Sa Ad(Tb^-1)^T + Sb).
 * phase 1: each chunk folds its steps from the identity (parallel over chunks)
 * phase 2: chunk totals are folded sequentially per chain (n_steps/chunk long)
 * phase 3: every element is prefixed by its chunk's carry (parallel over elements)
 * Deterministic, but re-associated w.r.t. the reference's strict left fold.
 * scratch: per chunk one ENTRY (carry-in), laid out after the per-step local prefixes in `entries`'
 * output slots (local prefixes are written to the final slots first, then overwritten in phase 3).
 * ---------------------------------------------------------------------------------------------- */
template <int D, int MODE>
__global__ void traj_scan_phase1(int total_chunks, int chunk, const FoldChain* __restrict__ chains, int n_chains,
                                 const int32_t* __restrict__ chunk_chain, const int32_t* __restrict__ chunk_first,
                                 const int32_t* __restrict__ out_idx, const double* __restrict__ dpose,
                                 const double* __restrict__ dcov, double* entries) {
  constexpr int E = Dim<D>::ENTRY, PS = Dim<D>::PS, NN = Dim<D>::N * Dim<D>::N;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total_chunks) return;
  const FoldChain ch = chains[chunk_chain[q]];
  const int s0 = chunk_first[q];
  const int s1 = min(s0 + chunk, ch.first_step + ch.n_steps);
  PoseT<D, MODE> cur, delta, nxt;
  pose_identity<D>(cur.pose);
  if (MODE == MODE_PCM) {
#pragma unroll
    for (int i = 0; i < NN; ++i) cur.cov[i] = 0.0;
  }
  cur.node = 0;
  cur.rot = true;
  for (int k = s0; k < s1; ++k) {
    from_factor<D, MODE>(dpose + (size_t)k * PS, dcov + (size_t)k * NN, delta);
    pt_compose<D, MODE>(cur, delta, nxt);
    store_entry<D, MODE>(entries + (size_t)out_idx[k] * E, 1, nxt);
    cur = nxt;
  }
}

template <int D, int MODE>
__global__ void traj_scan_phase2(int n_chains, int chunk, const FoldChain* __restrict__ chains,
                                 const int32_t* __restrict__ chain_first_chunk, const int32_t* __restrict__ out_idx,
                                 const double* entries, double* carry) {
  constexpr int E = Dim<D>::ENTRY;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const FoldChain ch = chains[c];
  const int nchunks = (ch.n_steps + chunk - 1) / chunk;
  PoseT<D, MODE> cur, tot, nxt;
  load_entry<D, MODE>(entries + (size_t)ch.start_idx * E, 1, cur);
  for (int q = 0; q < nchunks; ++q) {
    store_entry<D, MODE>(carry + (size_t)(chain_first_chunk[c] + q) * E, 1, cur);
    const int last = min(ch.first_step + (q + 1) * chunk, ch.first_step + ch.n_steps) - 1;
    load_entry<D, MODE>(entries + (size_t)out_idx[last] * E, 1, tot);
    pt_compose<D, MODE>(cur, tot, nxt);
    cur = nxt;
  }
}

template <int D, int MODE>
__global__ void traj_scan_phase3(int total_steps, int chunk, const int32_t* __restrict__ step_chunk,
                                 const int32_t* __restrict__ out_idx, const double* carry, double* entries) {
  constexpr int E = Dim<D>::ENTRY;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total_steps) return;
  PoseT<D, MODE> g, l, o;
  load_entry<D, MODE>(carry + (size_t)step_chunk[k] * E, 1, g);
  load_entry<D, MODE>(entries + (size_t)out_idx[k] * E, 1, l);
  pt_compose<D, MODE>(g, l, o);
  store_entry<D, MODE>(entries + (size_t)out_idx[k] * E, 1, o);
}

void launch_traj_scan_phases(int dim, int mode, int n_chains, const FoldChain* chains, const int32_t* out_idx,
                             const double* delta_pose, const double* delta_cov, double* entries, int chunk,
                             int total_steps, int total_chunks, const int32_t* chunk_chain, const int32_t* chunk_first,
                             const int32_t* chain_first_chunk, const int32_t* step_chunk, double* carry,
                             cudaStream_t st) {
  if (total_steps <= 0) return;
  const int T = 64;
#define CALL(D, M)                                                                                                  \
  traj_scan_phase1<D, M><<<(total_chunks + T - 1) / T, T, 0, st>>>(total_chunks, chunk, chains, n_chains, chunk_chain, \
                                                                  chunk_first, out_idx, delta_pose, delta_cov, entries); \
  traj_scan_phase2<D, M><<<n_chains, 1, 0, st>>>(n_chains, chunk, chains, chain_first_chunk, out_idx, entries, carry); \
  traj_scan_phase3<D, M><<<(total_steps + T - 1) / T, T, 0, st>>>(total_steps, chunk, step_chunk, out_idx, carry, entries)
  RPGO_DISPATCH(dim, mode, CALL);
#undef CALL
}

/* ------------------------------------------------------------------------------------------------
 * K2: closure constructor + odometry consistency check, one thread per new closure.
 *   result = getBetween(i, j).compose(T(lc).inverse());  ok = dist < threshold   (Pcm.h:604-629)
 * ---------------------------------------------------------------------------------------------- */
template <int D, int MODE>
__global__ void lc_prepare_kernel(int n, const double* __restrict__ pose, const double* __restrict__ cov,
                                  const int32_t* __restrict__ idx_front, const int32_t* __restrict__ idx_back,
                                  const uint8_t* __restrict__ check, const double* __restrict__ traj, Thresholds th,
                                  double* entries_out, uint8_t* ok_out, double* dist_out) {
  constexpr int E = Dim<D>::ENTRY, PS = Dim<D>::PS, NN = Dim<D>::N * Dim<D>::N;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  PoseT<D, MODE> lc;
  from_factor<D, MODE>(pose + (size_t)k * PS, cov + (size_t)k * NN, lc);
  store_entry<D, MODE>(entries_out + (size_t)k * E, 1, lc);
  bool ok = true;
  double dist = nan("");
  if (check[k]) {
    PoseT<D, MODE> a, b, pij, res;
    load_entry<D, MODE>(traj + (size_t)idx_front[k] * E, 1, a);
    load_entry<D, MODE>(traj + (size_t)idx_back[k] * E, 1, b);
    pt_between<D, MODE>(a, b, pij);
    pt_inverse_inplace<D, MODE>(lc);
    pt_compose<D, MODE>(pij, lc, res);
    bool near;
    ok = check_consistent<D, MODE>(res, th, true, &dist, &near);
  }
  ok_out[k] = ok ? 1 : 0;
  dist_out[k] = dist;
}

void launch_lc_prepare(int dim, int mode, int n, const double* pose, const double* cov, const int32_t* idx_front,
                       const int32_t* idx_back, const uint8_t* check, const double* traj, Thresholds th,
                       double* entries_out, uint8_t* ok_out, double* dist_out, cudaStream_t st) {
  if (n <= 0) return;
  const int T = 64;
#define CALL(D, M) \
  lc_prepare_kernel<D, M><<<(n + T - 1) / T, T, 0, st>>>(n, pose, cov, idx_front, idx_back, check, traj, th, entries_out, ok_out, dist_out)
  RPGO_DISPATCH(dim, mode, CALL);
#undef CALL
}

__global__ void scatter_entries_kernel(int n, int E, const double* __restrict__ entries, const uint64_t* __restrict__ dst) {
  const int k = blockIdx.x;
  if (k >= n) return;
  const uint64_t d = dst[k];
  if (d == 0) return;
  double* out = reinterpret_cast<double*>(d);
  for (int i = threadIdx.x; i < E; i += blockDim.x) out[i] = entries[(size_t)k * E + i];
}
void launch_scatter_entries(int dim, int n, const double* entries, const uint64_t* dst_ptrs, cudaStream_t st) {
  if (n <= 0) return;
  const int E = dim == 3 ? Dim<3>::ENTRY : Dim<2>::ENTRY;
  scatter_entries_kernel<<<n, 64, 0, st>>>(n, E, entries, dst_ptrs);
}

/* ------------------------------------------------------------------------------------------------
 * K3 (direct): warp = one older closure i x 32 consecutive newer closures j; lane = j.
 * The 32 decisions are packed with one __ballot_sync into the adjacency word (i, j/32).
 * Operands are gathered straight from the global tables (L2-resident); this is the reference
 * kernel the tiled variant is validated against.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ bool row_owned(const Shard& sh, int i) {
  if (sh.world <= 1) return true;
  const int64_t c = i / sh.chunk_rows;
  return c == sh.rank || c == 2 * (int64_t)sh.world - 1 - sh.rank;
}

template <int D, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) pairwise_direct_kernel(GroupView g, const double* __restrict__ traj, int j_begin, int w_begin,
                                                              Shard sh, Thresholds th, Flagged fl, double* dist_out) {
  constexpr int E = Dim<D>::ENTRY;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int w = w_begin + blockIdx.x;                 /* adjacency word (32 columns) */
  const int i = blockIdx.y * (blockDim.x >> 5) + wib; /* older closure */
  const int j = w * 32 + lane;                        /* newer closure */
  if (i >= g.n || i >= w * 32 + 31) return;           /* whole word on/below the diagonal */
  if (!row_owned(sh, i)) return;
  bool ok = false;
  if (j < g.n && j > i && j >= j_begin) {
    const uint8_t pa = g.pfx_front[i], pc = g.pfx_front[j];
    const int ia = g.idx_front[i], ib = g.idx_back[i];
    int ic = g.idx_front[j], id = g.idx_back[j];
    if (pa != pc) { const int t = ic; ic = id; id = t; } /* Pcm.h:691-698: keys swapped, measurement not inverted */
    double dist;
    bool near;
    ok = pair_check<D, MODE>(traj + (size_t)ia * E, 1, traj + (size_t)ib * E, 1, g.lc + (size_t)i * E, 1,
                             traj + (size_t)ic * E, 1, traj + (size_t)id * E, 1, g.lc + (size_t)j * E, 1, th, &dist,
                             &near);
    if (near) {
      const unsigned long long slot = atomicAdd(fl.count, 1ULL);
      if ((int64_t)slot < fl.cap) {
        fl.pairs[2 * slot] = i;
        fl.pairs[2 * slot + 1] = j;
      }
    }
    if (dist_out) {
      dist_out[(size_t)i * g.n + j] = dist;
      dist_out[(size_t)j * g.n + i] = dist;
    }
  }
  const unsigned word = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) {
    /* keep bits of columns < j_begin (computed earlier); columns <= i belong to the mirror pass */
    unsigned keep = 0;
    if (j_begin > w * 32) keep = (j_begin >= w * 32 + 32) ? 0xffffffffu : ((1u << (j_begin - w * 32)) - 1u);
    uint32_t* p = g.bits + (size_t)i * g.stride32 + w;
    *p = (*p & keep) | word;
  }
}

void launch_pairwise_direct(int dim, int mode, GroupView g, const double* traj, int j_begin, Shard sh, Thresholds th,
                            Flagged fl, double* dist_out, cudaStream_t st) {
  if (g.n < 2 || j_begin >= g.n) return;
  const int w_begin = j_begin / 32;
  const int w_end = (g.n + 31) / 32;
  const int warps = 4;
  dim3 grid(w_end - w_begin, (g.n + warps - 1) / warps);
#define CALL(D, M) pairwise_direct_kernel<D, M, 2><<<grid, warps * 32, 0, st>>>(g, traj, j_begin, w_begin, sh, th, fl, dist_out)
  RPGO_DISPATCH(dim, mode, CALL);
#undef CALL
}

/* ------------------------------------------------------------------------------------------------
 * N3: landmark re-observations.  Same tiling as the direct pairwise kernel: warp = older observation i x 32
 * newer observations j, decisions ballot-packed into word (i, j/32).  Groups are tiny (a handful of
 * observations per landmark), so this is latency- not throughput-relevant.
 * ---------------------------------------------------------------------------------------------- */
template <int D, int MODE>
__global__ void __launch_bounds__(128) landmark_direct_kernel(GroupView g, const double* __restrict__ traj, int j_begin,
                                                              int w_begin, Thresholds th, Flagged fl, double* dist_out) {
  constexpr int E = Dim<D>::ENTRY;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int w = w_begin + blockIdx.x;
  const int i = blockIdx.y * (blockDim.x >> 5) + wib;
  const int j = w * 32 + lane;
  if (i >= g.n || i >= w * 32 + 31) return;
  bool ok = false;
  if (j < g.n && j > i && j >= j_begin) {
    const bool cross = g.pfx_front[i] != g.pfx_front[j];
    double dist;
    bool near;
    ok = landmark_pair_check<D, MODE>(traj + (size_t)g.idx_front[i] * E, traj + (size_t)g.idx_front[j] * E,
                                      traj + (size_t)g.idx_a0[i] * E, traj, cross, g.lc + (size_t)i * E, g.lc + (size_t)j * E, th,
                                      &dist, &near);
    if (near) {
      const unsigned long long slot = atomicAdd(fl.count, 1ULL);
      if ((int64_t)slot < fl.cap) {
        fl.pairs[2 * slot] = i;
        fl.pairs[2 * slot + 1] = j;
      }
    }
    if (dist_out) {
      dist_out[(size_t)i * g.n + j] = dist;
      dist_out[(size_t)j * g.n + i] = dist;
    }
  }
  const unsigned word = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) {
    unsigned keep = 0;
    if (j_begin > w * 32) keep = (j_begin >= w * 32 + 32) ? 0xffffffffu : ((1u << (j_begin - w * 32)) - 1u);
    uint32_t* p = g.bits + (size_t)i * g.stride32 + w;
    *p = (*p & keep) | word;
  }
}

void launch_landmark_direct(int dim, int mode, GroupView g, const double* traj, int j_begin, Thresholds th, Flagged fl,
                            double* dist_out, cudaStream_t st) {
  if (g.n < 2 || j_begin >= g.n) return;
  const int w_begin = j_begin / 32;
  const int w_end = (g.n + 31) / 32;
  const int warps = 4;
  dim3 grid(w_end - w_begin, (g.n + warps - 1) / warps);
#define CALL(D, M) landmark_direct_kernel<D, M><<<grid, warps * 32, 0, st>>>(g, traj, j_begin, w_begin, th, fl, dist_out)
  RPGO_DISPATCH(dim, mode, CALL);
#undef CALL
}

/* ------------------------------------------------------------------------------------------------
 * N4: multi-robot frame alignment, front half (Pcm.h:1024-1055) and getRobotOdomValues (Pcm.h:1074-1082).
 * One thread per closure / trajectory entry; pose arithmetic only.
 * ---------------------------------------------------------------------------------------------- */
template <int D>
__global__ void frame_align_kernel(GroupView g, const double* __restrict__ traj, int E, uint8_t r0, int m,
                                   const int32_t* __restrict__ closure_idx, double* out) {
  constexpr int PS = Dim<D>::PS;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int c = closure_idx[t];
  Pose<D> Tfb, Tf, Tb;
  load_pose<D>(g.lc + (size_t)c * E, 1, Tfb);
  int ifr = g.idx_front[c], ibk = g.idx_back[c];
  if (g.pfx_front[c] != r0) { /* closure stated ri -> r0: swap the keys and invert the measurement */
    const int tmp = ifr; ifr = ibk; ibk = tmp;
    Tfb = inverse<D>(Tfb);
  }
  load_pose<D>(traj + (size_t)ifr * E, 1, Tf);
  load_pose<D>(traj + (size_t)ibk * E, 1, Tb);
  const Pose<D> r = compose<D>(compose<D>(Tf, Tfb), inverse<D>(Tb));
#pragma unroll
  for (int i = 0; i < PS; ++i) out[(size_t)t * PS + i] = r.m[i];
}
void launch_frame_align(int dim, GroupView g, const double* traj, int entry, uint8_t r0, int m, const int32_t* closure_idx,
                        double* out, cudaStream_t st) {
  if (m <= 0) return;
  const int T = 128;
  if (dim == 3) frame_align_kernel<3><<<(m + T - 1) / T, T, 0, st>>>(g, traj, entry, r0, m, closure_idx, out);
  else frame_align_kernel<2><<<(m + T - 1) / T, T, 0, st>>>(g, traj, entry, r0, m, closure_idx, out);
}

template <int D>
__global__ void transform_poses_kernel(const double* __restrict__ traj, int E, int m, const int32_t* __restrict__ entry_idx,
                                       const double* __restrict__ transform, double* out) {
  constexpr int PS = Dim<D>::PS;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  Pose<D> T, P;
  load_pose<D>(transform, 1, T);
  load_pose<D>(traj + (size_t)entry_idx[t] * E, 1, P);
  const Pose<D> r = compose<D>(T, P);
#pragma unroll
  for (int i = 0; i < PS; ++i) out[(size_t)t * PS + i] = r.m[i];
}
void launch_transform_poses(int dim, const double* traj, int entry, int m, const int32_t* entry_idx, const double* transform,
                            double* out, cudaStream_t st) {
  if (m <= 0) return;
  const int T = 128;
  if (dim == 3) transform_poses_kernel<3><<<(m + T - 1) / T, T, 0, st>>>(traj, entry, m, entry_idx, transform, out);
  else transform_poses_kernel<2><<<(m + T - 1) / T, T, 0, st>>>(traj, entry, m, entry_idx, transform, out);
}

/* ------------------------------------------------------------------------------------------------
 * mirror: the pairwise kernel writes the strictly-upper triangle (bit (i, j), i < j); this pass fills the lower one
 * (Pcm.h:756-763 sets both adj(i, j) and adj(j, i)).  HBM-bound bit-matrix transpose: n^2/8 bytes.
 * One block moves a 512 x 512-bit tile: rows are read as 64-byte segments into shared memory (row pitch 17 words: the 32
 * rows of a 32x32 block fall into 32 different banks), every 32x32 block is transposed inside a warp with the 5-step
 * shuffle butterfly and swapped with its mirror block, and the tile is written back as 64-byte row segments.
 * Diagonal tiles keep their own upper part.  Only tile columns >= j_begin / 512 hold new bits.
 * ---------------------------------------------------------------------------------------------- */
constexpr int MIRROR_T = 512, MIRROR_W = MIRROR_T / 32, MIRROR_PITCH = MIRROR_W + 1;

/* 32x32 bit transpose inside a warp: lane l holds row l, the result in lane c is column c (bit r = row r's bit c).
 * Five exchange steps with the partner lane l ^ j (j = 16, 8, 4, 2, 1); each step keeps one half of the own word and takes
 * the other half from the partner, moved by j bits.  The two coarse steps are byte permutations (one PRMT), the fine ones a
 * rotate plus one three-input logic op; the per-lane constants live in Tr32 (computed once per thread). */
struct Tr32 {
  uint32_t sel16, sel8;      /* PRMT selectors of the 16- and 8-bit steps */
  uint32_t keep4, keep2, keep1; /* bits kept from the own word in the 4 / 2 / 1-bit steps */
  uint32_t rot4, rot2, rot1; /* left-rotation of the partner's word */
  __device__ __forceinline__ explicit Tr32(int lane) {
    sel16 = (lane & 16) ? 0x3276u : 0x5410u;
    sel8 = (lane & 8) ? 0x3715u : 0x6240u;
    keep4 = (lane & 4) ? 0xF0F0F0F0u : 0x0F0F0F0Fu;
    keep2 = (lane & 2) ? 0xCCCCCCCCu : 0x33333333u;
    keep1 = (lane & 1) ? 0xAAAAAAAAu : 0x55555555u;
    rot4 = (lane & 4) ? 28u : 4u;
    rot2 = (lane & 2) ? 30u : 2u;
    rot1 = (lane & 1) ? 31u : 1u;
  }
  __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
    x = __byte_perm(x, __shfl_xor_sync(0xffffffffu, x, 16), sel16);
    x = __byte_perm(x, __shfl_xor_sync(0xffffffffu, x, 8), sel8);
    uint32_t y = __shfl_xor_sync(0xffffffffu, x, 4);
    x = (x & keep4) | (__funnelshift_l(y, y, rot4) & ~keep4);
    y = __shfl_xor_sync(0xffffffffu, x, 2);
    x = (x & keep2) | (__funnelshift_l(y, y, rot2) & ~keep2);
    y = __shfl_xor_sync(0xffffffffu, x, 1);
    x = (x & keep1) | (__funnelshift_l(y, y, rot1) & ~keep1);
    return x;
  }
};

/* bits of word `w` of row `row` that lie strictly above the diagonal (column > row) */
__device__ __forceinline__ uint32_t upper_mask(int row, int w) {
  const int c0 = w * 32;
  if (c0 > row) return 0xffffffffu;
  if (c0 + 31 <= row) return 0u;
  return ~((2u << (row - c0)) - 1u);
}

constexpr int MIRROR_PATCH = 16; /* concurrently resident blocks cover compact 16 x 16-tile patches: neighbouring tiles touch
                                    neighbouring 64-byte pieces of the same DRAM pages at about the same time, on the read side
                                    (same rows, adjacent columns) and on the write side (same mirrored rows) alike */

__global__ void __launch_bounds__(256) mirror_tile_kernel(uint32_t* bits, int64_t stride32, int n, int tc_begin, int tiles,
                                                          int patches_x) {
  __shared__ uint32_t S[MIRROR_T * MIRROR_PITCH];
  const int patch = blockIdx.x / (MIRROR_PATCH * MIRROR_PATCH), inner = blockIdx.x % (MIRROR_PATCH * MIRROR_PATCH);
  const int tc = tc_begin + (patch % patches_x) * MIRROR_PATCH + inner % MIRROR_PATCH;
  const int tr = (patch / patches_x) * MIRROR_PATCH + inner / MIRROR_PATCH;
  if (tc >= tiles || tr > tc) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rsub = lane >> 2, q4 = (lane & 3) * 4; /* a warp moves 8 rows x 64 bytes per instruction, 16 bytes per lane */
  const bool diag = tr == tc;
  const bool edge = diag || (tc + 1) * MIRROR_T > n;
  /* load the source tile: rows tr*512.., words tc*16.. (all eight 16-byte loads of a thread are in flight together) */
  {
    uint4 v[MIRROR_T / 64];
#pragma unroll
    for (int it = 0; it < MIRROR_T / 64; ++it) {
      const int gr = tr * MIRROR_T + it * 64 + warp * 8 + rsub, gw = tc * MIRROR_W + q4;
      v[it] = make_uint4(0u, 0u, 0u, 0u);
      if (gr < n && gw < stride32 && gw * 32 < n) v[it] = *reinterpret_cast<const uint4*>(bits + (size_t)gr * stride32 + gw);
    }
#pragma unroll
    for (int it = 0; it < MIRROR_T / 64; ++it) {
      const int r = it * 64 + warp * 8 + rsub;
      const int gr = tr * MIRROR_T + r, gw = tc * MIRROR_W + q4;
      uint32_t x[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
      if (edge) { /* tiles on the diagonal or at the right border: mask what is not strictly-upper / beyond column n */
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c0 = (gw + k) * 32;
          if (c0 >= n) x[k] = 0u;
          else if (c0 + 32 > n) x[k] &= (1u << (n - c0)) - 1u;
          if (diag) x[k] &= upper_mask(gr, gw + k);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) S[r * MIRROR_PITCH + q4 + k] = x[k];
    }
  }
  __syncthreads();
  /* transpose: block (a, b) <-> block (b, a); the 136 pairs a <= b are spread evenly over the warps */
  {
    const Tr32 tr32(lane);
    for (int item = warp; item < MIRROR_W * (MIRROR_W + 1) / 2; item += 8) {
      int b = (int)((sqrtf(8.0f * (float)item + 1.0f) - 1.0f) * 0.5f); /* triangular root; exact for item < 2^20 after the fix-up */
      b += ((b + 1) * (b + 2) / 2 <= item) ? 1 : 0;
      b -= (b * (b + 1) / 2 > item) ? 1 : 0;
      const int a = item - b * (b + 1) / 2;
      const uint32_t x = S[(32 * a + lane) * MIRROR_PITCH + b];
      const uint32_t y = S[(32 * b + lane) * MIRROR_PITCH + a];
      const uint32_t xt = tr32(x), yt = tr32(y);
      __syncwarp();
      S[(32 * b + lane) * MIRROR_PITCH + a] = xt;
      if (a != b) S[(32 * a + lane) * MIRROR_PITCH + b] = yt;
    }
  }
  __syncthreads();
  /* store the mirrored tile: rows tc*512.., words tr*16.. */
#pragma unroll
  for (int it = 0; it < MIRROR_T / 64; ++it) {
    const int r = it * 64 + warp * 8 + rsub;
    const int gr = tc * MIRROR_T + r, gw = tr * MIRROR_W + q4;
    if (gr < n && gw < stride32 && gw * 32 < n) {
      uint4* p = reinterpret_cast<uint4*>(bits + (size_t)gr * stride32 + gw);
      uint4 v = make_uint4(S[r * MIRROR_PITCH + q4], S[r * MIRROR_PITCH + q4 + 1], S[r * MIRROR_PITCH + q4 + 2], S[r * MIRROR_PITCH + q4 + 3]);
      if (diag) { /* keep this row's own upper part */
        const uint4 o = *p;
        v.x |= o.x & upper_mask(gr, gw);
        v.y |= o.y & upper_mask(gr, gw + 1);
        v.z |= o.z & upper_mask(gr, gw + 2);
        v.w |= o.w & upper_mask(gr, gw + 3);
      }
      *p = v;
    }
  }
}

void launch_mirror(uint32_t* bits, int64_t stride32, int n, int j_begin, cudaStream_t st) {
  if (n < 2 || j_begin >= n) return;
  const int tiles = (n + MIRROR_T - 1) / MIRROR_T;
  const int tc_begin = j_begin / MIRROR_T;
  const int patches_x = (tiles - tc_begin + MIRROR_PATCH - 1) / MIRROR_PATCH, patches_y = (tiles + MIRROR_PATCH - 1) / MIRROR_PATCH;
#ifdef RPGO_MIRROR_CARVEOUT
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(mirror_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, RPGO_MIRROR_CARVEOUT);
#endif
  mirror_tile_kernel<<<patches_x * patches_y * MIRROR_PATCH * MIRROR_PATCH, 256, 0, st>>>(bits, stride32, n, tc_begin, tiles, patches_x);
}

/* degree: popcount of each row, one warp per row */
__global__ void degree_kernel(const uint32_t* __restrict__ bits, int64_t stride32, int n, int32_t* deg) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int words = (n + 31) / 32;
  int c = 0;
  for (int w = lane; w < words; w += 32) c += __popc(bits[(size_t)row * stride32 + w]);
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) deg[row] = c;
}
void launch_degree(const uint32_t* bits, int64_t stride32, int n, int32_t* deg, cudaStream_t st) {
  if (n <= 0) return;
  const int warps = 8;
  degree_kernel<<<(n + warps - 1) / warps, warps * 32, 0, st>>>(bits, stride32, n, deg);
}

/* removeLastLoopClosure: clear row n_after and column n_after (Pcm.h:320-323 keeps the leading block) */
__global__ void clear_last_kernel(uint32_t* bits, int64_t stride32, int n_after) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int words = (n_after + 32) / 32;
  if (t < n_after) bits[(size_t)t * stride32 + n_after / 32] &= ~(1u << (n_after & 31));
  if (t < words) bits[(size_t)n_after * stride32 + t] = 0u;
}
void launch_clear_last(uint32_t* bits, int64_t stride32, int n_after, cudaStream_t st) {
  const int T = 128;
  clear_last_kernel<<<(n_after + 1 + T - 1) / T + 1, T, 0, st>>>(bits, stride32, n_after);
}

/* ------------------------------------------------------------------------------------------------
 * FP64 peak micro-benchmark (roofline denominator for K3): 8 independent DFMA chains per thread.
 * ---------------------------------------------------------------------------------------------- */
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

double fp64_peak_tflops(cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 512, blocks = sms * 4, iters = 1 << 15;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace rpgo
