// Development probe: FP64 tensor-core (DMMA m8n8k4) throughput on sm_100a, alone and mixed with DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC, int NFMA>
__global__ void k(double* out, int iters, double a, double b) {
  double c[NACC][2];
  double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NFMA; ++i) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NFMA; ++i) s += f[i];
  if (s == 1.2345) out[0] = s;
}
template <int NACC, int NFMA>
void run(int warps, int ctas) {
  double* out; cudaMalloc(&out, 8);
  int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC, NFMA><<<148 * ctas, warps * 32>>>(out, 10, 0.999, 1e-9);
  cudaEventRecord(e0);
  k<NACC, NFMA><<<148 * ctas, warps * 32>>>(out, iters, 0.999, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mma_flops = 2.0 * 256 * NACC * 4.0 * iters * warps * ctas * 148;
  double fma_flops = 2.0 * 32 * NFMA * 4.0 * iters * warps * ctas * 148;
  printf("warps/SM=%2d NACC=%d NFMA=%d: DMMA %.1f TF + DFMA %.1f TF = %.1f TF  (%.2f ms)\n", warps * ctas, NACC, NFMA,
         mma_flops / ms / 1e9, fma_flops / ms / 1e9, (mma_flops + fma_flops) / ms / 1e9, ms);
  cudaFree(out);
}
int main() {
  run<1, 0>(4, 1); run<2, 0>(4, 1); run<4, 0>(4, 1); run<8, 0>(4, 1);
  run<4, 0>(8, 1); run<8, 0>(8, 2); run<4, 0>(16, 2);
  run<0 + 1, 8>(8, 2); run<4, 8>(8, 2); run<4, 16>(8, 2); run<2, 16>(8, 2); run<8, 8>(8, 2);
  return 0;
}
