// Gaussian-LOSVD path (ParametricComponent.evaluate_ybar,
// /root/reference/popkinmocks/components/parametric.py:355-373):
//
//   Y-hat[k,x] = sum_t E[k,t,x] * Q[k,t,x],   Q[k,t,x] = sum_z P[t,x,z] * FXw[k,z,t],
//   E[k,t,x]   = exp(-i mu_v[t,x] w_k - (sig_v[t,x] w_k)^2 / 2),
//
// with w_k = linspace(0, pi, nfo)[k] / dv handed over by the host (the reference's own array; it
// is NOT the true DFT frequency for odd n_fft).
//
//   * Q is a small-K GEMM per age t ([pixels x nz] times [nz x (k, re|im)]): it runs on the FP64
//     tensor cores (DMMA m8n8k4).  Warp = 8 pixels x 64 frequencies; a lane ends up with
//     Q(re, im) for pixel lane/4 and frequencies lane%4 + 4j.  FXw is pre-arranged on the host in
//     DMMA fragment order and staged per (frequency tile, age) by one TMA bulk copy, double
//     buffered, shared by the 8 warps (8 pixel blocks) of the CTA.
//   * E is evaluated exactly (sincos + 2 exp) once per (t, x, lane) and then advanced along the
//     lane's frequencies (stride D = w_4) by recurrence: e_{j+1} = e_j r_j, r_{j+1} = r_j beta with
//     r_0 = exp(-sig^2 D (w_k0 + D/2)) exp(-i mu D), beta = exp(-sig^2 D^2): 6 FP64 ops per
//     frequency instead of ~80 for sincos + exp; 16 steps, so the round-off growth is ~1e-15.
// The inverse FFT is shared with the density path (fft_kernels.cu).
#include <cstdio>

#include "pkm_internal.h"

constexpr int GA_NB = 16;      // frequency blocks of 4 per warp (64 frequencies per CTA tile)
constexpr int GA_WARPS = 8;    // pixel blocks of 8 per CTA
constexpr int GA_KTILE = 4 * GA_NB;

namespace {

__device__ __forceinline__ unsigned ga_smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ga_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct GaGeom {
  long long mu_s0, mu_s1, mu_s2, sg_s0, sg_s1, sg_s2, nx2, npix_total, pix0;
  int npix_tile, nt, nz, nks, nfo;
  double D;  // omega[4] - omega[0]: frequency stride of one lane
};

template <int NKS>
__global__ void __launch_bounds__(GA_WARPS * 32, 2)
k_gaussian(const double* __restrict__ Sg, const double* __restrict__ omega,
           const double* __restrict__ P, const double* __restrict__ mu,
           const double* __restrict__ sig, GaGeom g, double2* __restrict__ yhat) {
  extern __shared__ __align__(128) double ga_smem[];
  constexpr int STAGE = GA_NB * NKS * 32;                      // doubles per (tile, age) slab
  double* sS = ga_smem;                                        // [2][STAGE]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sS + 2 * STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int kt = blockIdx.x;
  const int x = (blockIdx.y * GA_WARPS + warp) * 8 + fr;       // pixel of this lane within the tile
  const long long xg = g.pix0 + min(x, g.npix_tile - 1);       // lanes past the tile recompute the last pixel
  const long long x1 = xg / g.nx2, x2 = xg - x1 * g.nx2;
  const int k0 = kt * GA_KTILE + fk;
  const double om0 = __ldg(omega + min(k0, g.nfo - 1));
  const double D = g.D;
  const unsigned bar0 = ga_smem_u32(bars);
  constexpr unsigned BYTES = STAGE * sizeof(double);
  const double* Sk = Sg + (size_t)kt * g.nt * STAGE;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 2 && s < g.nt; ++s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ga_smem_u32(sS + s * STAGE)), "l"(Sk + (size_t)s * STAGE), "r"(BYTES), "r"(bar0 + 8u * s) : "memory");
    }
  }
  __syncthreads();

  double2 acc[GA_NB];
#pragma unroll
  for (int nb = 0; nb < GA_NB; ++nb) acc[nb] = make_double2(0.0, 0.0);

  for (int t = 0; t < g.nt; ++t) {
    const int stage = t & 1;
    const unsigned parity = (t >> 1) & 1;
    // this lane's A fragments: P[t, x, z = ks*4 + fk]
    double pf[NKS];
    const double* pp = P + ((size_t)t * g.npix_total + xg) * g.nz;
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int z = ks * 4 + fk;
      pf[ks] = z < g.nz ? __ldg(pp + z) : 0.0;
    }
    const double m = __ldg(mu + t * g.mu_s0 + x1 * g.mu_s1 + x2 * g.mu_s2);
    const double sg = __ldg(sig + t * g.sg_s0 + x1 * g.sg_s1 + x2 * g.sg_s2);
    // exact E at the lane's first frequency, then the recurrence factors
    double sn, cs;
    sincos(m * om0, &sn, &cs);
    const double so = sg * om0;
    const double g0 = exp(-0.5 * (so * so));
    double2 e = make_double2(g0 * cs, -g0 * sn);
    sincos(m * D, &sn, &cs);
    const double a0 = exp(-(sg * sg) * D * (om0 + 0.5 * D));
    double2 r = make_double2(a0 * cs, -a0 * sn);
    const double beta = exp(-(sg * sg) * (D * D));

    {  // wait for this age's S slab
      asm volatile(
          "{\n.reg .pred p;\nGA_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra GA_DONE;\nbra GA_WAIT;\nGA_DONE:\n}\n" ::"r"(bar0 + 8u * stage), "r"(parity) : "memory");
    }
    const double* sb = sS + stage * STAGE + lane;
#pragma unroll
    for (int nb = 0; nb < GA_NB; ++nb) {
      double qr = 0.0, qi = 0.0;
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) ga_dmma(qr, qi, pf[ks], sb[(nb * NKS + ks) * 32]);
      acc[nb].x = fma(e.x, qr, acc[nb].x);
      acc[nb].x = fma(-e.y, qi, acc[nb].x);
      acc[nb].y = fma(e.x, qi, acc[nb].y);
      acc[nb].y = fma(e.y, qr, acc[nb].y);
      const double ex = e.x * r.x - e.y * r.y;
      e.y = e.x * r.y + e.y * r.x;
      e.x = ex;
      r.x *= beta;
      r.y *= beta;
    }
    __syncthreads();  // every warp is done with this stage
    if (threadIdx.x == 0 && t + 2 < g.nt) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * stage), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ga_smem_u32(sS + stage * STAGE)), "l"(Sk + (size_t)(t + 2) * STAGE), "r"(BYTES), "r"(bar0 + 8u * stage) : "memory");
    }
  }
  if (x < g.npix_tile) {
    double2* yo = yhat + (size_t)x * g.nfo;
#pragma unroll
    for (int nb = 0; nb < GA_NB; ++nb) {
      const int k = k0 + 4 * nb;
      if (k < g.nfo) yo[k] = acc[nb];
    }
  }
}

// ----------------------------------------------------------------------------
// float32-arithmetic variant (plan precision = 1; BASELINE.json's optional fp32 path, tolerance 1e-5).
// Warp = 32 pixels (one per lane) x 16 frequencies k0 + 4 j of a 64-frequency tile, 4 warps per CTA.
// FXw is staged per (tile, age) as float2 [64 k][nz] by one TMA bulk copy, double buffered, and
// read as warp-uniform broadcasts; P(t,x,:) sits in registers.  The metallicity contraction and
// the E recurrence run on the FP32 pipe; E is still evaluated exactly in FP64 once per (t, x, warp)
// (mu w reaches ~100 rad: sincosf would cost 1e-5 by itself), and the per-age float32 partial
// sums are folded into FP64 accumulators every 8 ages.
// ----------------------------------------------------------------------------
constexpr int GF_WARPS = 4;
constexpr int GF_FOLD = 8;

// phi - 2 pi rint(phi / 2 pi) in FP64 (two-term 2 pi), returned as float in [-pi, pi]
__device__ __forceinline__ float reduce_phase(double phi) {
  const double n = rint(phi * 0.15915494309189533577);
  double r = fma(n, -6.28318530717958623200, phi);
  r = fma(n, -2.44929359829470635445e-16, r);
  return (float)r;
}

template <int NZ4>
__global__ void __launch_bounds__(GF_WARPS * 32, 3)
k_gaussian_f32(const float2* __restrict__ Sg, const double* __restrict__ omega, const double* __restrict__ P,
               const double* __restrict__ mu, const double* __restrict__ sig, GaGeom g,
               double2* __restrict__ yhat) {
  extern __shared__ __align__(128) unsigned char gf_smem[];
  const int nz = g.nz;
  constexpr int NZP = 4 * NZ4;                                   // metallicities padded to a multiple of 4 (zeros)
  constexpr int STAGE = GA_KTILE * NZP;                          // float2 per (tile, age) slab
  float2* sS = reinterpret_cast<float2*>(gf_smem);               // [2][STAGE]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sS + 2 * STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x;
  const int x = blockIdx.y * 32 + lane;
  const long long xg = g.pix0 + min(x, g.npix_tile - 1);
  const long long x1 = xg / g.nx2, x2 = xg - x1 * g.nx2;
  const int k0 = kt * GA_KTILE + warp;                           // this warp's frequencies: k0 + 4 j
  const double om0 = __ldg(omega + min(k0, g.nfo - 1));
  const double D = g.D;                                          // omega[4] - omega[0]
  const unsigned bar0 = ga_smem_u32(bars);
  const unsigned BYTES = (unsigned)STAGE * sizeof(float2);
  const float2* Sk = Sg + (size_t)kt * g.nt * STAGE;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 2 && s < g.nt; ++s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ga_smem_u32(sS + s * STAGE)), "l"(Sk + (size_t)s * STAGE), "r"(BYTES), "r"(bar0 + 8u * s) : "memory");
    }
  }
  __syncthreads();

  double2 acc[GA_NB];
  float2 part[GA_NB];
#pragma unroll
  for (int j = 0; j < GA_NB; ++j) {
    acc[j] = make_double2(0.0, 0.0);
    part[j] = make_float2(0.f, 0.f);
  }
  for (int t = 0; t < g.nt; ++t) {
    const int stage = t & 1;
    const unsigned parity = (t >> 1) & 1;
    float pf[NZP];                                               // P(t, x, z)
    const double* pp = P + ((size_t)t * g.npix_total + xg) * nz;
#pragma unroll
    for (int z = 0; z < NZP; ++z) pf[z] = z < nz ? (float)__ldg(pp + z) : 0.f;
    const double m = __ldg(mu + t * g.mu_s0 + x1 * g.mu_s1 + x2 * g.mu_s2);
    const double sg = __ldg(sig + t * g.sg_s0 + x1 * g.sg_s1 + x2 * g.sg_s2);
    // E at the warp's first frequency and the recurrence factors: the PHASES are reduced to
    // [-pi, pi] in FP64 (mu w reaches ~100 rad), everything else is float32 (sincosf / expf on
    // small arguments; a large damping exponent only matters where the result is ~0 anyway)
    float sn, cs;
    sincosf(reduce_phase(m * om0), &sn, &cs);
    const float so = (float)(sg * om0);
    const float g0 = expf(-0.5f * (so * so));
    float2 e = make_float2(g0 * cs, -g0 * sn);
    sincosf(reduce_phase(m * D), &sn, &cs);
    const float a0 = expf((float)(-(sg * sg) * D * (om0 + 0.5 * D)));
    float2 r = make_float2(a0 * cs, -a0 * sn);
    const float beta = expf((float)(-(sg * sg) * (D * D)));
    {
      asm volatile(
          "{\n.reg .pred p;\nGF_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra GF_DONE;\nbra GF_WAIT;\nGF_DONE:\n}\n" ::"r"(bar0 + 8u * stage), "r"(parity) : "memory");
    }
    const float2* sb = sS + stage * STAGE;
#pragma unroll
    for (int j = 0; j < GA_NB; ++j) {
      // warp-uniform broadcast reads, two metallicities per 16-byte load
      const float4* srow = reinterpret_cast<const float4*>(sb + (warp + 4 * j) * NZP);
      float qr = 0.f, qi = 0.f;
#pragma unroll
      for (int z2 = 0; z2 < NZP / 2; ++z2) {
        const float4 s = srow[z2];
        qr = fmaf(pf[2 * z2], s.x, qr);
        qi = fmaf(pf[2 * z2], s.y, qi);
        qr = fmaf(pf[2 * z2 + 1], s.z, qr);
        qi = fmaf(pf[2 * z2 + 1], s.w, qi);
      }
      part[j].x = fmaf(e.x, qr, part[j].x);
      part[j].x = fmaf(-e.y, qi, part[j].x);
      part[j].y = fmaf(e.x, qi, part[j].y);
      part[j].y = fmaf(e.y, qr, part[j].y);
      const float ex = e.x * r.x - e.y * r.y;
      e.y = e.x * r.y + e.y * r.x;
      e.x = ex;
      r.x *= beta;
      r.y *= beta;
    }
    if ((t % GF_FOLD) == GF_FOLD - 1 || t == g.nt - 1) {
#pragma unroll
      for (int j = 0; j < GA_NB; ++j) {
        acc[j].x += (double)part[j].x;
        acc[j].y += (double)part[j].y;
        part[j] = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (threadIdx.x == 0 && t + 2 < g.nt) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * stage), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ga_smem_u32(sS + stage * STAGE)), "l"(Sk + (size_t)(t + 2) * STAGE), "r"(BYTES), "r"(bar0 + 8u * stage) : "memory");
    }
  }
  if (x < g.npix_tile) {
    double2* yo = yhat + (size_t)x * g.nfo;
#pragma unroll
    for (int j = 0; j < GA_NB; ++j) {
      const int k = k0 + 4 * j;
      if (k < g.nfo) yo[k] = acc[j];
    }
  }
}

}  // namespace

// FXw in DMMA B-fragment order: [ktile][t][nb][ks][lane], lane holds S_ri[k][z] with
// z = ks*4 + lane%4, column c = lane/4 -> k = ktile*64 + nb*4 + c/2, ri = c%2 (zero padded).
void gaussian_tables_upload(pkm_plan* pl, const double* FXw) {
  const int nt = pl->d.nt, nz = pl->d.nz, nfo = pl->nfo;
  const int nks = (nz + 3) / 4;
  PKM_REQUIRE(nks <= 8, "nz=%d > 32 unsupported by the Gaussian-path kernel", nz);
  const int nkt = (nfo + GA_KTILE - 1) / GA_KTILE;
  pl->ga_nks = nks;
  pl->ga_nkt = nkt;
  std::vector<double> h((size_t)nkt * nt * GA_NB * nks * 32, 0.0);
  for (int kt = 0; kt < nkt; ++kt)
    for (int t = 0; t < nt; ++t)
      for (int nb = 0; nb < GA_NB; ++nb)
        for (int ks = 0; ks < nks; ++ks)
          for (int lane = 0; lane < 32; ++lane) {
            const int z = ks * 4 + lane % 4, c = lane / 4;
            const int k = kt * GA_KTILE + nb * 4 + c / 2, ri = c % 2;
            if (k < nfo && z < nz)
              h[((((size_t)kt * nt + t) * GA_NB + nb) * nks + ks) * 32 + lane] =
                  FXw[size_t(2) * ((size_t(k) * nz + z) * nt + t) + ri];
          }
  PKM_CUDA(cudaMalloc(&pl->d_SA, h.size() * sizeof(double)));
  PKM_CUDA(cudaMemcpy(pl->d_SA, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (pl->d.precision == 1) {
    // float32 path: FXw as float2 [ktile][t][64 k][nzp] (zero padded beyond nfo and nz; nzp = nz
    // rounded up to a multiple of 4)
    const int nzp = 4 * nks;
    std::vector<float> f((size_t)2 * nkt * nt * GA_KTILE * nzp, 0.f);
    for (int kt = 0; kt < nkt; ++kt)
      for (int t = 0; t < nt; ++t)
        for (int kk = 0; kk < GA_KTILE; ++kk) {
          const int k = kt * GA_KTILE + kk;
          if (k >= nfo) continue;
          for (int z = 0; z < nz; ++z) {
            const size_t o = 2 * ((((size_t)kt * nt + t) * GA_KTILE + kk) * nzp + z);
            f[o] = (float)FXw[size_t(2) * ((size_t(k) * nz + z) * nt + t)];
            f[o + 1] = (float)FXw[size_t(2) * ((size_t(k) * nz + z) * nt + t) + 1];
          }
        }
    PKM_CUDA(cudaMalloc(&pl->d_SA32, f.size() * sizeof(float)));
    PKM_CUDA(cudaMemcpy(pl->d_SA32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
}

template <int NKS>
static void launch_gaussian_t(pkm_plan* pl, const GaGeom& g, const double* P, const double* mu,
                              const double* sig, double2* yhat, cudaStream_t st) {
  const size_t smem = sizeof(double) * 2 * GA_NB * NKS * 32 + 2 * sizeof(unsigned long long);
  PKM_CUDA(cudaFuncSetAttribute(k_gaussian<NKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(pl->ga_nkt, (g.npix_tile + GA_WARPS * 8 - 1) / (GA_WARPS * 8));
  KernelTimer tm(pl, PKM_K_GAUSSIAN, st);
  k_gaussian<NKS><<<grid, GA_WARPS * 32, smem, st>>>(reinterpret_cast<const double*>(pl->d_SA), pl->d_omega, P, mu,
                                                     sig, g, yhat);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}

void launch_gaussian(pkm_plan* pl, const double* P_txz, const double* mu, const int64_t* mu_st,
                     const double* sig, const int64_t* sig_st, int64_t nx2, int64_t npix_total,
                     int64_t pix0, int npix_tile, int xpad, double omega_step4, double2* yhat,
                     cudaStream_t st) {
  (void)xpad;
  GaGeom g;
  g.mu_s0 = mu_st[0]; g.mu_s1 = mu_st[1]; g.mu_s2 = mu_st[2];
  g.sg_s0 = sig_st[0]; g.sg_s1 = sig_st[1]; g.sg_s2 = sig_st[2];
  g.nx2 = nx2; g.npix_total = npix_total; g.pix0 = pix0; g.npix_tile = npix_tile;
  g.nt = pl->d.nt; g.nz = pl->d.nz; g.nks = pl->ga_nks; g.nfo = pl->nfo; g.D = omega_step4;
  if (pl->d.precision == 1) {
    const size_t smem = sizeof(float2) * 2 * GA_KTILE * 4 * pl->ga_nks + 2 * sizeof(unsigned long long);
    dim3 grid(pl->ga_nkt, (g.npix_tile + 31) / 32);
    KernelTimer tm(pl, PKM_K_GAUSSIAN, st);
    switch (pl->ga_nks) {
#define CASE(N)                                                                                                     \
  case N:                                                                                                           \
    PKM_CUDA(cudaFuncSetAttribute(k_gaussian_f32<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    k_gaussian_f32<N><<<grid, GF_WARPS * 32, smem, st>>>(pl->d_SA32, pl->d_omega, P_txz, mu, sig, g, yhat);         \
    break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
      default: PKM_THROW(PKM_ERR_INVALID, "unsupported nz=%d", pl->d.nz);
    }
    PKM_CUDA(cudaGetLastError());
    pl->launches++;
    return;
  }
  switch (pl->ga_nks) {
#define CASE(N) case N: launch_gaussian_t<N>(pl, g, P_txz, mu, sig, yhat, st); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: PKM_THROW(PKM_ERR_INVALID, "unsupported nz=%d", pl->d.nz);
  }
}
