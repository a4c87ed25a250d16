// Development probe: DFMA latency and throughput vs warps/ILP on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  int iters = 2000;
  k<ILP><<<148, warps * 32>>>(out, iters, 0.9999, 1e-9, cyc);
  cudaDeviceSynchronize();
  k<ILP><<<148, warps * 32>>>(out, iters, 0.9999, 1e-9, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_instr = (double)h / (iters * 8.0 * ILP);   // cycles per warp-DFMA issued by this warp
  double per_sm = warps * 32.0 / per_instr;             // FMA lanes per clk per SM
  printf("warps/SM=%2d ILP=%2d: %.2f clk per DFMA per warp -> %.1f FMA/clk/SM\n", warps, ILP, per_instr, per_sm);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
