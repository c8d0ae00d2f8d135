// Dependent-issue latency of FP64 / shuffle / shared-memory operations on one warp (sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[64];
  double x = a + threadIdx.x;
  long long t0, t1;
  const int N = 4096;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  double y = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = y * b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  double z = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = z + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  double s = z;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // two independent chains interleaved: throughput view
  double p = s, q = s + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { p = fma(p, b, a); q = fma(q, b, a); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  double p4[8];
  for (int k = 0; k < 8; ++k) p4[k] = p + k;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p4[k] = fma(p4[k], b, a);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  sm[threadIdx.x] = p4[0];
  __syncwarp();
  double w = q;
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { w = sm[idx]; idx = ((int)w) & 31; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  double acc = 0;
  for (int k = 0; k < 8; ++k) acc += p4[k];
  out[threadIdx.x] = x + y + z + s + p + q + acc + w;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
  lat<<<1, 32>>>(out, cyc, 1e-3, 0.999);
  lat<<<1, 32>>>(out, cyc, 1e-3, 0.999);
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const double N = 4096;
  printf("DFMA dependent latency  %.1f cycles\nDMUL %.1f\nDADD %.1f\nSHFL.64 (2x32) %.1f\n2 interleaved DFMA chains: %.1f cycles per pair\n8 chains: %.1f cycles per 8\nLDS.64 + cvt dependent %.1f\n",
         h[0] / N, h[1] / N, h[2] / N, h[3] / N, h[4] / N, h[5] / N, h[6] / N);
  return 0;
}
