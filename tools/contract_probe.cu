// Development harness for the spline-contraction kernel (k_contract, density_kernels.cu):
// times instruction-order / staging variants of
//     Y-hat[k,x] = sum_tz S'[k,tz] * (c3 + sum_{q<3} w[k][q] (c_q - c3))
// on synthetic data of the headline shape (nfo = 2454, ntz = 636, nfi = 51, one 7168-pixel tile)
// and checks every variant against variant 0 (a copy of the shipped kernel).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/contract_probe tools/contract_probe.cu
//   tools/contract_probe [npix] [variant-mask]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int TZC = 12;

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  for (unsigned spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double2 ldg_nc_f64x2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// one (t,z) step for KT frequencies: G frequencies interleaved, spline FMAs ordered so that the
// d_q component (slot B) is shared by G consecutive DFMAs, MACs so that 3 of 4 share an operand
template <int KT, int G>
__device__ __forceinline__ void tz_step(const double (&w)[KT][3], const double2 (&d)[3], const double2 c3,
                                        const double2* __restrict__ srow, double2 (&acc)[KT]) {
#pragma unroll
  for (int h = 0; h < KT; h += G) {
    double pr[G], pi[G];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (h + g < KT) pr[g] = fma(w[h + g][q], d[q].x, q == 0 ? c3.x : pr[g]);
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (h + g < KT) pi[g] = fma(w[h + g][q], d[q].y, q == 0 ? c3.y : pi[g]);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (h + g >= KT) continue;
      const double2 s = srow[h + g];
      acc[h + g].x = fma(s.x, pr[g], acc[h + g].x);
      acc[h + g].y = fma(s.x, pi[g], acc[h + g].y);
      acc[h + g].x = fma(-s.y, pi[g], acc[h + g].x);
      acc[h + g].y = fma(s.y, pr[g], acc[h + g].y);
    }
  }
}

// ---------------------------------------------------------------- variant A: the shipped structure
// warp = (segment of KT frequencies) x 32 pixels, private S' ring (TMA), c rows by LDG one step ahead.
// G = 0 reproduces the shipped instruction order (2 frequencies interleaved, per-frequency spline).
template <int KT, int G, int MINB, int PF>
__global__ void __launch_bounds__(128, MINB)
k_va(const double2* __restrict__ Sw, const double* __restrict__ Ww, const int4* __restrict__ wsegs,
     const double2* __restrict__ cT, double2* __restrict__ yhat, int nwseg, int ntz, int nchunk, int nfi,
     int np, int nfo, long long nitems) {
  constexpr int WARPS = 4, STAGES = 3, CH = TZC * KT;
  constexpr unsigned CHB = CH * sizeof(double2);
  extern __shared__ __align__(128) unsigned char smem[];
  double2 (*sS)[STAGES][CH] = reinterpret_cast<double2 (*)[STAGES][CH]>(smem);
  unsigned long long (*sBar)[STAGES] =
      reinterpret_cast<unsigned long long (*)[STAGES]>(smem + sizeof(double2) * WARPS * STAGES * CH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * WARPS + warp;
  if (item >= nitems) return;
  const int pg = (int)(item / nwseg), ws = (int)(item - (long long)pg * nwseg);
  const int4 sg = __ldg(wsegs + ws);
  const int x = pg * 32 + lane;
  const int xl = min(x, np - 1);
  double w[KT][3];
#pragma unroll
  for (int r = 0; r < KT; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) w[r][q] = __ldg(Ww + ((size_t)ws * KT + r) * 4 + q);
  double2 acc[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = make_double2(0.0, 0.0);
  const double2* Sg = Sw + (size_t)ws * nchunk * CH;
  const unsigned bar0 = smem_u32(&sBar[warp][0]);
  const unsigned buf0 = smem_u32(&sS[warp][0][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      if (s < nchunk) {
        mbar_expect_tx(bar0 + 8u * s, CHB);
        bulk_g2s(buf0 + CHB * s, Sg + (size_t)s * CH, CHB, bar0 + 8u * s);
      }
  }
  __syncwarp();
  const size_t c_step = (size_t)nfi * 32;
  const double2* cp = cT + (((size_t)pg * ntz) * nfi + sg.x) * 32 + (xl & 31);
  double2 c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q] = ldg_nc_f64x2(cp + q * 32);
  int stage = 0;
  unsigned parity = 0;
  int tz = 0;
  for (int ci = 0; ci < nchunk; ++ci) {
    mbar_wait(bar0 + 8u * stage, parity);
    const double2* sbuf = &sS[warp][stage][0];
    const int nrow = min(TZC, ntz - ci * TZC);
    for (int j = 0; j < nrow; ++j, ++tz) {
      double2 d[3];
      const double2 c3 = c[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) d[q] = make_double2(c[q].x - c3.x, c[q].y - c3.y);
      cp += (tz + 1 < ntz) ? c_step : 0;
      if (tz + PF < ntz) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)(PF - 1) * c_step + q * 32));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) c[q] = ldg_nc_f64x2(cp + q * 32);
      if (G > 0) {
        tz_step<KT, (G > 0 ? G : 1)>(w, d, c3, sbuf + j * KT, acc);
      } else {
#pragma unroll
        for (int h = 0; h < KT; h += 2) {
          double pr[2], pi[2];
          double2 sv[2];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            sv[g] = sbuf[j * KT + h + g];
            pr[g] = c3.x;
            pi[g] = c3.y;
          }
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              pr[g] = fma(w[h + g][q], d[q].x, pr[g]);
              pi[g] = fma(w[h + g][q], d[q].y, pi[g]);
            }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            acc[h + g].x = fma(sv[g].x, pr[g], acc[h + g].x);
            acc[h + g].y = fma(sv[g].x, pi[g], acc[h + g].y);
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            acc[h + g].x = fma(-sv[g].y, pi[g], acc[h + g].x);
            acc[h + g].y = fma(sv[g].y, pr[g], acc[h + g].y);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0 && ci + STAGES < nchunk) {
      mbar_expect_tx(bar0 + 8u * stage, CHB);
      bulk_g2s(buf0 + CHB * stage, Sg + (size_t)(ci + STAGES) * CH, CHB, bar0 + 8u * stage);
    }
    if (++stage == STAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (x < np) {
    double2* yo = yhat + (size_t)x * nfo + sg.y;
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < sg.z) yo[r] = acc[r];
  }
}

// ---------------------------------------------------------------- variant B: span CTAs
// CTA = (32 pixels) x (up to W consecutive warp segments of ONE spline span).  The 4 coefficient
// rows of the span (2 KB per (t,z)) and the S' rows of the CTA's segments are staged in shared
// memory by TMA bulk copies, TZS (t,z) steps per stage, NST stages, full/empty mbarriers; the warps
// only read shared memory.  Coefficient traffic from L2 drops W-fold and nothing in the inner loop
// waits on global memory, so few warps with long reuse-friendly DFMA runs can feed the FP64 pipe.
template <int KT, int G, int W, int TZS, int NST, int MINB>
__global__ void __launch_bounds__(W * 32, MINB)
k_vb(const double2* __restrict__ Sw, const double* __restrict__ Ww, const int4* __restrict__ wsegs,
     const int4* __restrict__ citems, const double2* __restrict__ cT, double2* __restrict__ yhat,
     int ncitem, int ntz, int srows, int nfi, int np, int nfo) {
  constexpr unsigned C_BYTES = TZS * 4 * 32 * sizeof(double2);      // TZS x 2 KB
  constexpr unsigned S_BYTES = TZS * KT * sizeof(double2);          // per warp segment
  constexpr unsigned STAGE_BYTES = C_BYTES + W * S_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned sm0 = smem_u32(smem);
  const unsigned full0 = sm0 + NST * STAGE_BYTES, empty0 = full0 + 8u * NST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci = blockIdx.x % ncitem, pg = blockIdx.x / ncitem;
  const int4 it = __ldg(citems + ci);          // {first warp segment, number of segments, first coefficient row, -}
  const int nseg = it.y;
  const int nstep = (ntz + TZS - 1) / TZS;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, nseg);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= nseg) return;
  const int ws = it.x + warp;
  const int4 sg = __ldg(wsegs + ws);
  const int x = pg * 32 + lane;

  auto issue = [&](int step) {   // thread 0 only
    const int s = step % NST;
    const unsigned dst = sm0 + s * STAGE_BYTES, bar = full0 + 8u * s;
    mbar_expect_tx(bar, C_BYTES + nseg * S_BYTES);
#pragma unroll
    for (int i = 0; i < TZS; ++i) {
      const int tz = min(step * TZS + i, ntz - 1);
      bulk_g2s(dst + i * 2048u, cT + (((size_t)pg * ntz + tz) * nfi + it.z) * 32, 2048u, bar);
    }
    for (int wi = 0; wi < nseg; ++wi)
      bulk_g2s(dst + C_BYTES + wi * S_BYTES, Sw + ((size_t)(it.x + wi) * srows + (size_t)step * TZS) * KT, S_BYTES, bar);
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < NST && s < nstep; ++s) issue(s);

  double w[KT][3];
#pragma unroll
  for (int r = 0; r < KT; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) w[r][q] = __ldg(Ww + ((size_t)ws * KT + r) * 4 + q);
  double2 acc[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = make_double2(0.0, 0.0);

  for (int step = 0; step < nstep; ++step) {
    const int s = step % NST;
    const unsigned par = (step / NST) & 1u;
    // the producer refills the stage every warp finished one step ago
    if (threadIdx.x == 0 && step >= 1 && step - 1 + NST < nstep) {
      const int sp = (step - 1) % NST;
      mbar_wait(empty0 + 8u * sp, ((step - 1) / NST) & 1u);
      issue(step - 1 + NST);
    }
    __syncwarp();
    mbar_wait(full0 + 8u * s, par);
    const double2* cst = reinterpret_cast<const double2*>(smem + s * STAGE_BYTES);
    const double2* sst = reinterpret_cast<const double2*>(smem + s * STAGE_BYTES + C_BYTES + warp * S_BYTES);
#pragma unroll
    for (int i = 0; i < TZS; ++i) {
      const double2 c3 = cst[(i * 4 + 3) * 32 + lane];
      double2 d[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double2 cq = cst[(i * 4 + q) * 32 + lane];
        d[q] = make_double2(cq.x - c3.x, cq.y - c3.y);
      }
      tz_step<KT, G>(w, d, c3, sst + i * KT, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8u * s);
  }
  if (x < np) {
    double2* yo = yhat + (size_t)x * nfo + sg.y;
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < sg.z) yo[r] = acc[r];
  }
}

// ---------------------------------------------------------------- variant C: weights in uniform registers
// As variant A, but the 4 warps of a CTA work on the SAME warp segment (4 pixel groups), so the
// segment index depends on blockIdx only and the 3 x KT spline weights can be read from the
// kernel-parameter bank with a uniform index: ptxas keeps them in UNIFORM registers (LDCU.64 +
// DFMA with a UR operand).  The spline DFMAs then read two vector registers instead of three
// (tools/dfma_rf.cu: 33 vs 24.5 TFLOP/s) and 60 vector registers are freed.
template <int KT>
constexpr int wp_segs() { return 30000 / (KT * 24); }
template <int KT>
struct WParams {
  double w[wp_segs<KT>()][KT][3];
};

template <int KT, int G, int MINB, int PF>
__global__ void __launch_bounds__(128, MINB)
k_vc(const __grid_constant__ WParams<KT> wp, int ws0, int nws, const double2* __restrict__ Sw,
     const int4* __restrict__ wsegs, const double2* __restrict__ cT, double2* __restrict__ yhat, int ntz,
     int nchunk, int nfi, int np, int nfo, int npg) {
  constexpr int WARPS = 4, STAGES = 3, CH = TZC * KT;
  constexpr unsigned CHB = CH * sizeof(double2);
  extern __shared__ __align__(128) unsigned char smem[];
  double2 (*sS)[STAGES][CH] = reinterpret_cast<double2 (*)[STAGES][CH]>(smem);
  unsigned long long (*sBar)[STAGES] =
      reinterpret_cast<unsigned long long (*)[STAGES]>(smem + sizeof(double2) * WARPS * STAGES * CH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wsl = blockIdx.x % nws;                 // uniform: segment within this launch
  const int pg = (blockIdx.x / nws) * WARPS + warp;
  if (pg >= npg) return;
  const int ws = ws0 + wsl;
  const int4 sg = __ldg(wsegs + ws);
  const int x = pg * 32 + lane;
  const int xl = min(x, np - 1);
  const double (&w)[KT][3] = wp.w[wsl];
  double2 acc[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = make_double2(0.0, 0.0);
  const double2* Sg = Sw + (size_t)ws * nchunk * CH;
  const unsigned bar0 = smem_u32(&sBar[warp][0]);
  const unsigned buf0 = smem_u32(&sS[warp][0][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      if (s < nchunk) {
        mbar_expect_tx(bar0 + 8u * s, CHB);
        bulk_g2s(buf0 + CHB * s, Sg + (size_t)s * CH, CHB, bar0 + 8u * s);
      }
  }
  __syncwarp();
  const size_t c_step = (size_t)nfi * 32;
  const double2* cp = cT + (((size_t)pg * ntz) * nfi + sg.x) * 32 + (xl & 31);
  double2 c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q] = ldg_nc_f64x2(cp + q * 32);
  int stage = 0;
  unsigned parity = 0;
  int tz = 0;
  for (int ci = 0; ci < nchunk; ++ci) {
    mbar_wait(bar0 + 8u * stage, parity);
    const double2* sbuf = &sS[warp][stage][0];
    const int nrow = min(TZC, ntz - ci * TZC);
    for (int j = 0; j < nrow; ++j, ++tz) {
      double2 d[3];
      const double2 c3 = c[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) d[q] = make_double2(c[q].x - c3.x, c[q].y - c3.y);
      cp += (tz + 1 < ntz) ? c_step : 0;
      if (tz + PF < ntz) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)(PF - 1) * c_step + q * 32));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) c[q] = ldg_nc_f64x2(cp + q * 32);
      tz_step<KT, G>(w, d, c3, sbuf + j * KT, acc);
    }
    __syncwarp();
    if (lane == 0 && ci + STAGES < nchunk) {
      mbar_expect_tx(bar0 + 8u * stage, CHB);
      bulk_g2s(buf0 + CHB * stage, Sg + (size_t)(ci + STAGES) * CH, CHB, bar0 + 8u * stage);
    }
    if (++stage == STAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (x < np) {
    double2* yo = yhat + (size_t)x * nfo + sg.y;
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < sg.z) yo[r] = acc[r];
  }
}

// ---------------------------------------------------------------- host side
struct Tables {
  int KT, nwseg, ncitem;
  std::vector<int> seg;      // [nwseg][4] {row, k0, cnt, 0}
  std::vector<double> w;     // [nwseg][KT][4]
  std::vector<int> citem;    // [ncitem][4] {first wseg, nseg, row, 0}
};

static Tables build_tables(int nfi, int nfo, int KT, int W) {
  Tables t;
  t.KT = KT;
  std::vector<int> span(nfo);
  std::vector<double> tw((size_t)nfo * 4);
  srand(7);
  for (int k = 0; k < nfo; ++k) {
    int i = (int)std::floor((double)k * (nfi - 1) / (nfo - 1)) - 1;
    span[k] = std::min(std::max(i, 0), nfi - 4);
    double s = 0, v[4];
    for (int q = 0; q < 4; ++q) { v[q] = 0.1 + rand() / (double)RAND_MAX; s += v[q]; }
    for (int q = 0; q < 4; ++q) tw[(size_t)k * 4 + q] = v[q] / s;
  }
  int k = 0;
  while (k < nfo) {
    int cnt = 1;
    while (cnt < KT && k + cnt < nfo && span[k + cnt] == span[k]) ++cnt;
    t.seg.insert(t.seg.end(), {span[k], k, cnt, 0});
    for (int r = 0; r < KT; ++r)
      for (int q = 0; q < 4; ++q) t.w.push_back(r < cnt ? tw[(size_t)(k + r) * 4 + q] : 0.0);
    k += cnt;
  }
  t.nwseg = (int)t.seg.size() / 4;
  int s = 0;
  while (s < t.nwseg) {
    int n = 1;
    while (n < W && s + n < t.nwseg && t.seg[4 * (s + n)] == t.seg[4 * s]) ++n;
    t.citem.insert(t.citem.end(), {s, n, t.seg[4 * s], 0});
    s += n;
  }
  t.ncitem = (int)t.citem.size() / 4;
  return t;
}

struct Problem {
  int nfi = 51, nfo = 2454, ntz = 636, np = 7168;
  int srows;                 // padded (t,z) rows of S'
  double2* d_cT = nullptr;
  double2* d_y = nullptr;
  std::vector<double> ref;   // variant 0 result
};

struct DevTables {
  int KT, nwseg, ncitem;
  double2* Sw;
  double* Ww;
  int4* wsegs;
  int4* citems;
};

static DevTables upload(const Problem& P, const Tables& t, const std::vector<double>& Sfull /* [nfo][ntz] complex */) {
  DevTables d;
  d.KT = t.KT; d.nwseg = t.nwseg; d.ncitem = t.ncitem;
  std::vector<double> Sw((size_t)2 * t.nwseg * P.srows * t.KT, 0.0);
  for (int ws = 0; ws < t.nwseg; ++ws) {
    const int k0 = t.seg[4 * ws + 1], cnt = t.seg[4 * ws + 2];
    for (int tz = 0; tz < P.ntz; ++tz)
      for (int r = 0; r < cnt; ++r) {
        const size_t o = 2 * (((size_t)ws * P.srows + tz) * t.KT + r);
        Sw[o] = Sfull[2 * ((size_t)(k0 + r) * P.ntz + tz)];
        Sw[o + 1] = Sfull[2 * ((size_t)(k0 + r) * P.ntz + tz) + 1];
      }
  }
  CK(cudaMalloc(&d.Sw, Sw.size() * 8));
  CK(cudaMemcpy(d.Sw, Sw.data(), Sw.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.Ww, t.w.size() * 8));
  CK(cudaMemcpy(d.Ww, t.w.data(), t.w.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.wsegs, t.seg.size() * 4));
  CK(cudaMemcpy(d.wsegs, t.seg.data(), t.seg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.citems, t.citem.size() * 4));
  CK(cudaMemcpy(d.citems, t.citem.data(), t.citem.size() * 4, cudaMemcpyHostToDevice));
  return d;
}

template <class F>
static void run_variant(const char* name, Problem& P, F&& launch, bool is_ref = false) {
  CK(cudaMemset(P.d_y, 0xff, sizeof(double2) * (size_t)P.np * P.nfo));
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e));
    exit(2);   // context is gone after a trap
  }
  std::vector<double> y((size_t)2 * P.np * P.nfo);
  CK(cudaMemcpy(y.data(), P.d_y, y.size() * 8, cudaMemcpyDeviceToHost));
  double err = 0, mx = 0;
  if (is_ref) P.ref = y;
  for (size_t i = 0; i < y.size(); ++i) {
    mx = std::max(mx, std::fabs(P.ref[i]));
    const double dd = std::fabs(y[i] - P.ref[i]);
    if (!(dd <= err)) err = dd;   // NaN propagates
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f, tot = 0.f;
  const int reps = 4;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
    tot += ms;
  }
  const double flops = 20.6 * P.nfo * (double)P.ntz * P.np;
  printf("%-44s %8.3f ms (mean %8.3f)  %6.2f TFLOP/s  rel.err %.2e\n", name, best, tot / reps,
         flops / (best * 1e-3) / 1e12, err / mx);
  fflush(stdout);
}

template <int KT, int G, int MINB, int PF>
static void launch_va(const Problem& P, const DevTables& d) {
  constexpr int CH = TZC * KT;
  const long long npg = (P.np + 31) / 32, nitems = npg * d.nwseg;
  const size_t smem = sizeof(double2) * 4 * 3 * CH + 8 * 4 * 3;
  CK(cudaFuncSetAttribute(k_va<KT, G, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_va<KT, G, MINB, PF><<<(unsigned)((nitems + 3) / 4), 128, smem>>>(d.Sw, d.Ww, d.wsegs, P.d_cT, P.d_y, d.nwseg, P.ntz,
                                                                   P.srows / TZC, P.nfi, P.np, P.nfo, nitems);
  CK(cudaGetLastError());
}

template <int KT, int G, int W, int TZS, int NST, int MINB>
static void launch_vb(const Problem& P, const DevTables& d) {
  const size_t smem = (size_t)NST * (TZS * 2048 + W * TZS * KT * 16) + 16 * NST;
  CK(cudaFuncSetAttribute(k_vb<KT, G, W, TZS, NST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long npg = (P.np + 31) / 32;
  k_vb<KT, G, W, TZS, NST, MINB><<<(unsigned)(npg * d.ncitem), W * 32, smem>>>(
      d.Sw, d.Ww, d.wsegs, d.citems, P.d_cT, P.d_y, d.ncitem, P.ntz, P.srows, P.nfi, P.np, P.nfo);
  CK(cudaGetLastError());
}

template <int KT, int G, int MINB, int PF>
static void launch_vc(const Problem& P, const DevTables& d, const Tables& t) {
  constexpr int CH = TZC * KT;
  const int npg = (P.np + 31) / 32;
  const size_t smem = sizeof(double2) * 4 * 3 * CH + 8 * 4 * 3;
  CK(cudaFuncSetAttribute(k_vc<KT, G, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static WParams<KT> wp;
  constexpr int WP_SEGS = wp_segs<KT>();
  for (int ws0 = 0; ws0 < d.nwseg; ws0 += WP_SEGS) {
    const int nws = std::min(WP_SEGS, d.nwseg - ws0);
    for (int i = 0; i < nws; ++i)
      for (int r = 0; r < KT; ++r)
        for (int q = 0; q < 3; ++q) wp.w[i][r][q] = t.w[((size_t)(ws0 + i) * KT + r) * 4 + q];
    const unsigned grid = (unsigned)(((npg + 3) / 4) * nws);
    k_vc<KT, G, MINB, PF><<<grid, 128, smem>>>(wp, ws0, nws, d.Sw, d.wsegs, P.d_cT, P.d_y, P.ntz, P.srows / TZC, P.nfi,
                                             P.np, P.nfo, npg);
    CK(cudaGetLastError());
  }
}

__global__ void k_fill_rand(double* p, size_t n, unsigned long long seed, double scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = seed + i;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    p[i] = ((double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5) * scale;
  }
}

int main(int argc, char** argv) {
  Problem P;
  if (argc > 1) P.np = atoi(argv[1]);
  const unsigned mask = argc > 2 ? (unsigned)strtoul(argv[2], nullptr, 0) : 0xffffffffu;
  P.srows = ((P.ntz + TZC - 1) / TZC) * TZC;
  const size_t ncoef = (size_t)((P.np + 31) / 32) * P.ntz * P.nfi * 32;
  CK(cudaMalloc(&P.d_cT, ncoef * sizeof(double2)));
  CK(cudaMalloc(&P.d_y, sizeof(double2) * (size_t)P.np * P.nfo));
  const double cscale = argc > 3 ? atof(argv[3]) : 1.0;
  k_fill_rand<<<148 * 8, 256>>>(reinterpret_cast<double*>(P.d_cT), ncoef * 2, 11, cscale);
  CK(cudaDeviceSynchronize());
  std::vector<double> Sfull((size_t)2 * P.nfo * P.ntz);
  srand(3);
  const double sscale = argc > 4 ? atof(argv[4]) : 1.0;
  for (auto& v : Sfull) v = (rand() / (double)RAND_MAX - 0.5) * sscale;

  const Tables t10 = build_tables(P.nfi, P.nfo, 10, 5);
  const DevTables d10 = upload(P, t10, Sfull);
  const Tables t5 = build_tables(P.nfi, P.nfo, 5, 10);
  const DevTables d5 = upload(P, t5, Sfull);
  const Tables t13 = build_tables(P.nfi, P.nfo, 13, 4);
  const DevTables d13 = upload(P, t13, Sfull);
  const Tables t17 = build_tables(P.nfi, P.nfo, 17, 3);
  const DevTables d17 = upload(P, t17, Sfull);
  const Tables t25 = build_tables(P.nfi, P.nfo, 25, 2);
  const DevTables d25 = upload(P, t25, Sfull);
  printf("nwseg13=%d nwseg17=%d nwseg25=%d\n", d13.nwseg, d17.nwseg, d25.nwseg);
  printf("np=%d nwseg10=%d ncitem10=%d nwseg5=%d ncitem5=%d\n", P.np, d10.nwseg, d10.ncitem, d5.nwseg, d5.ncitem);

  int v = 0;
#define RUN(name, call)                                         \
  do {                                                          \
    if (v == 0 || (mask >> v) & 1u) run_variant(name, P, [&] { call; }, v == 0); \
    ++v;                                                        \
  } while (0)
  RUN("A  shipped order  KT10 G2  3cta PF12", (launch_va<10, 0, 3, 12>(P, d10)));
  RUN("A  reuse order    KT10 G2  3cta PF12", (launch_va<10, 2, 3, 12>(P, d10)));
  RUN("A  reuse order    KT10 G5  3cta PF12", (launch_va<10, 5, 3, 12>(P, d10)));
  RUN("A  reuse order    KT10 G10 2cta PF12", (launch_va<10, 10, 2, 12>(P, d10)));
  RUN("A  reuse order    KT10 G5  2cta PF12", (launch_va<10, 5, 2, 12>(P, d10)));
  RUN("B  span CTA KT10 G10 W5 TZS4 NST4 2cta", (launch_vb<10, 10, 5, 4, 4, 2>(P, d10)));
  RUN("B  span CTA KT10 G5  W5 TZS4 NST4 2cta", (launch_vb<10, 5, 5, 4, 4, 2>(P, d10)));
  RUN("B  span CTA KT10 G2  W5 TZS4 NST4 2cta", (launch_vb<10, 2, 5, 4, 4, 2>(P, d10)));
  RUN("B  span CTA KT10 G10 W5 TZS4 NST4 1cta", (launch_vb<10, 10, 5, 4, 4, 1>(P, d10)));
  RUN("B  span CTA KT10 G5  W5 TZS4 NST4 3cta", (launch_vb<10, 5, 5, 4, 4, 3>(P, d10)));
  RUN("B  span CTA KT10 G10 W5 TZS2 NST6 2cta", (launch_vb<10, 10, 5, 2, 6, 2>(P, d10)));
  RUN("B  span CTA KT5  G5  W10 TZS4 NST4 2cta", (launch_vb<5, 5, 10, 4, 4, 2>(P, d5)));
  RUN("B  span CTA KT5  G5  W10 TZS4 NST4 1cta", (launch_vb<5, 5, 10, 4, 4, 1>(P, d5)));
  RUN("B  span CTA KT5  G5  W10 TZS4 NST4 3cta", (launch_vb<5, 5, 10, 4, 4, 3>(P, d5)));
  RUN("C  UR weights  KT10 G2  3cta PF12", (launch_vc<10, 2, 3, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G2  4cta PF12", (launch_vc<10, 2, 4, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G5  4cta PF12", (launch_vc<10, 5, 4, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G10 3cta PF12", (launch_vc<10, 10, 3, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G5  5cta PF12", (launch_vc<10, 5, 5, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G5  3cta PF12", (launch_vc<10, 5, 3, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G10 4cta PF12", (launch_vc<10, 10, 4, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G1  4cta PF12", (launch_vc<10, 1, 4, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G1  5cta PF12", (launch_vc<10, 1, 5, 12>(P, d10, t10)));
  RUN("C  UR weights  KT10 G5  4cta PF6 ", (launch_vc<10, 5, 4, 6>(P, d10, t10)));
  RUN("C  UR weights  KT10 G5  4cta PF24", (launch_vc<10, 5, 4, 24>(P, d10, t10)));
  RUN("C  UR weights  KT13 G5  4cta PF12", (launch_vc<13, 5, 4, 12>(P, d13, t13)));
  RUN("C  UR weights  KT13 G5  3cta PF12", (launch_vc<13, 5, 3, 12>(P, d13, t13)));
  RUN("C  UR weights  KT17 G6  3cta PF12", (launch_vc<17, 6, 3, 12>(P, d17, t17)));
  RUN("C  UR weights  KT17 G6  4cta PF12", (launch_vc<17, 6, 4, 12>(P, d17, t17)));
  RUN("C  UR weights  KT25 G5  3cta PF12", (launch_vc<25, 5, 3, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G5  2cta PF12", (launch_vc<25, 5, 2, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G1  2cta PF12", (launch_vc<25, 1, 2, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G1  3cta PF12", (launch_vc<25, 1, 3, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G25 2cta PF12", (launch_vc<25, 25, 2, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G2  3cta PF12", (launch_vc<25, 2, 3, 12>(P, d25, t25)));
  RUN("C  UR weights  KT25 G5  2cta PF4 ", (launch_vc<25, 5, 2, 4>(P, d25, t25)));
  return 0;
}
