// Gaussian-LOSVD path (ParametricComponent.evaluate_ybar,
// /root/reference/popkinmocks/components/parametric.py:355-373):
//
//   Y-hat[k,x] = sum_t exp(-i mu_v[t,x] w_k - (sig_v[t,x] w_k)^2 / 2) * sum_z P[t,x,z] * FXw[k,z,t]
//
// with w_k = linspace(0, pi, nfo)[k] / dv handed over by the host (the reference's own array; it
// is NOT the true DFT frequency for odd n_fft).  Same thread mapping as the
// density contraction: warp = 4 frequency groups x 8 pixels, thread = KT frequencies of one pixel,
// FXw streamed from L2 in [t][z][k] order, accumulation over (t,z) in registers.  The inverse
// FFT is shared with the density path (fft_kernels.cu).
#include "pkm_internal.h"

constexpr int GA_KT = 4;   // frequencies per thread
constexpr int GA_GPW = 4;  // frequency groups per warp (x 8 pixels = 32 lanes)

__global__ void __launch_bounds__(256, 2)
k_gaussian(const double2* __restrict__ SA, const double* __restrict__ omega,
           const double* __restrict__ P, const double* __restrict__ mu, long long mu_s0,
           long long mu_s1, long long mu_s2, const double* __restrict__ sig, long long sg_s0,
           long long sg_s1, long long sg_s2, long long nx2, long long npix_total, long long pix0,
           int npix_tile, int nt, int nz, int kpad, int nunit, int nfo, int xpad,
           double2* __restrict__ yhat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 8 + warp;
  if (unit >= nunit) return;
  const int px = lane & 7, gsub = lane >> 3;
  const int x = blockIdx.y * 8 + px;
  const bool live = x < npix_tile;
  const long long xg = pix0 + (live ? x : 0);
  const long long x1 = xg / nx2, x2 = xg - x1 * nx2;
  const int k0 = (unit * GA_GPW + gsub) * GA_KT;
  double om[GA_KT];
#pragma unroll
  for (int r = 0; r < GA_KT; ++r) om[r] = omega[k0 + r];  // padded with zeros
  double2 acc[GA_KT];
#pragma unroll
  for (int r = 0; r < GA_KT; ++r) acc[r] = make_double2(0.0, 0.0);

  for (int t = 0; t < nt; ++t) {
    const double m = mu[t * mu_s0 + x1 * mu_s1 + x2 * mu_s2];
    const double sg = sig[t * sg_s0 + x1 * sg_s1 + x2 * sg_s2];
    double2 q[GA_KT];
#pragma unroll
    for (int r = 0; r < GA_KT; ++r) q[r] = make_double2(0.0, 0.0);
    const double* pp = P + ((long long)t * npix_total + xg) * nz;
    const double2* sp = SA + (size_t)t * nz * kpad + k0;
    for (int z = 0; z < nz; ++z) {
      const double pz = __ldg(pp + z);
#pragma unroll
      for (int r = 0; r < GA_KT; ++r) {
        double2 s = __ldg(sp + r);
        q[r].x = fma(pz, s.x, q[r].x);
        q[r].y = fma(pz, s.y, q[r].y);
      }
      sp += kpad;
    }
#pragma unroll
    for (int r = 0; r < GA_KT; ++r) {
      const double so = sg * om[r];
      const double g = exp(-0.5 * (so * so));
      double sn, cs;
      sincos(m * om[r], &sn, &cs);
      const double er = g * cs, ei = -g * sn;
      acc[r].x = fma(er, q[r].x, acc[r].x);
      acc[r].x = fma(-ei, q[r].y, acc[r].x);
      acc[r].y = fma(er, q[r].y, acc[r].y);
      acc[r].y = fma(ei, q[r].x, acc[r].y);
    }
  }
  if (!live) {
#pragma unroll
    for (int r = 0; r < GA_KT; ++r) acc[r] = make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int r = 0; r < GA_KT; ++r)
    if (k0 + r < nfo) yhat[(size_t)x * nfo + k0 + r] = acc[r];
}

void launch_gaussian(pkm_plan* pl, const double* P_txz, const double* mu, const int64_t* mu_st,
                     const double* sig, const int64_t* sig_st, int64_t nx2, int64_t npix_total,
                     int64_t pix0, int npix_tile, int xpad, double2* yhat, cudaStream_t st) {
  const int nunit = pl->kpadA / (GA_GPW * GA_KT);
  dim3 grid((nunit + 7) / 8, (npix_tile + 7) / 8);
  KernelTimer tm(pl, PKM_K_GAUSSIAN, st);
  k_gaussian<<<grid, 256, 0, st>>>(pl->d_SA, pl->d_omega, P_txz, mu, mu_st[0], mu_st[1], mu_st[2], sig,
                                   sig_st[0], sig_st[1], sig_st[2], nx2, npix_total, pix0, npix_tile,
                                   pl->d.nt, pl->d.nz, pl->kpadA, nunit, pl->nfo, xpad, yhat);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}
