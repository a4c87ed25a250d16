// Density path (Component.evaluate_ybar, /root/reference/popkinmocks/components/base.py:836-859)
// as two FP64 kernels.  With  F = dv * rfft(roll(p))  and the not-a-knot spline written as
// P-hat[k] = sum_q w[k][q] c[i_k + q],  c = A^{-1} F,  the datacube spectrum of one pixel is
//
//     Y-hat[k] = sum_{t,z} S'[k,t,z] * sum_{q<4} w[k][q] * c[i_k+q, t, z]
//
// kernel 1 (coeff):    c = (A^{-1} D) p   - a small dense real->complex transform per
//                      (t,x,z) column, done as an FP64 tensor-core (DMMA) GEMM on TMA-staged
//                      density slabs, with the even/odd fold of the real DFT (halves the
//                      flops) and exp(log p) fused into the fold.  Variants: k_coeff_flow
//                      (default: warp = z x 8 pixels, no CTA barrier in the tile loop, stores
//                      straight from the accumulators), k_coeff_oct (the same items with a TMA
//                      tensor store per tile) and the staged k_coeff (any shape / alignment).
// kernel 2 (contract): streams S' (TMA-staged, L2 resident) and the coefficient rows, evaluates
//                      the 4-tap spline on the fly and accumulates over (t,z) in registers.
// P-hat (25 MB per pixel in the reference, base.py:846-853) is never materialised.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "exp_tab.cuh"
#include "pkm_internal.h"

#ifndef PKM_K1_APREFETCH
#define PKM_K1_APREFETCH 0
#endif

// ============================================================================
// async-copy helpers (1-D TMA bulk copies completing on an mbarrier)
// ============================================================================
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one 2-D box of a tiled tensor map -> shared memory (dense [box rows][box cols])
__device__ __forceinline__ void tensor2d_g2s(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// ============================================================================
// kernel 1, staged variant (the fallback; the default is k_coeff_oct further down)
// ============================================================================
// Persistent CTAs (one per SM), each looping over tiles = (age t, block of PXB pixels), i.e.
// NC = PXB*nz density columns of nv velocities.  The folded DFT*spline matrices stay resident
// in shared memory for the CTA's lifetime.  The density slab is double buffered: while tile i is
// being transformed, the [nv][NC] slab of tile i+1 streams into the other buffer as ONE 2-D TMA
// tensor copy (box = nv velocity rows x NC contiguous columns of the [nt*nv][npix*nz] matrix view
// of the density) completing on an mbarrier; layouts the tensor map cannot describe fall back
// to nv 1-D bulk copies, unaligned ones to plain loads.  Per tile:
//   phase 1: fold v <-> -v in place (e = sum, o = difference of the rolled profile), exp(log p)
//            fused here when the caller passes log densities
//   phase 2: warp-tiled FP64 GEMM; lanes = 8 row groups x 4 column groups, thread tile TM x 4;
//            one work item = (16-column chunk, real | imaginary part) per warp
//   phase 3: results staged as [z][m][x] in the (now consumed) slab buffer
//   phase 4: written as 16*PXB-byte runs so the contraction kernel reads pixel-contiguous rows.
// The kernel sits at the FP64 / HBM balance point: 10.3 kFLOP and 1.6 KB of traffic per column.
constexpr int K1_THREADS = 384;
constexpr int K1_WARPS = K1_THREADS / 32;

struct K1Geom {
  long long npix_total, pix0;
  int npix_tile, xpad, nv, nz, nfi, ne, no, pxb, nt, is_log, v0;
  int use_tma;  // 0 = plain loads, 1 = one 1-D bulk copy per velocity row, 2 = one 2-D tensor copy per tile
  int ngroup;   // independent warp groups per CTA (1 or 2), each running its own tile pipeline
  int out_f32;  // store the coefficients as float2 (float32 arithmetic mode of the contraction)
  int ncp;      // octet kernel: slab pitch in doubles (box width of the tensor copy)
};

// velocity row holding rolled position +u / -u  (np.roll(p, nv - v0), base.py:843)
__device__ __forceinline__ int row_plus(int u, int v0, int nv) {
  int r = u + v0;
  return r >= nv ? r - nv : r;
}
__device__ __forceinline__ int row_minus(int u, int v0, int nv) {
  int r = v0 - u;
  return r < 0 ? r + nv : r;
}

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor cores.  Lane l holds A[l/4][l%4], B[l%4][l/4]
// and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#ifdef PKM_K1_PROF
__device__ unsigned long long k1_prof[32];
#define K1_TICK(slot)                                                        \
  do {                                                                       \
    if (blockIdx.x == 0 && lane == 0 && warp == 0) {                         \
      const long long now_ = clock64();                                      \
      atomicAdd(&k1_prof[grp * 16 + (slot)], (unsigned long long)(now_ - tick_)); \
      tick_ = now_;                                                          \
    }                                                                        \
  } while (0)
#else
#define K1_TICK(slot) do {} while (0)
#endif

template <int TM>
__global__ void __launch_bounds__(K1_THREADS, 1)
k_coeff(const double* __restrict__ dens, const __grid_constant__ CUtensorMap tmap, K1Geom g,
        const double* __restrict__ CRt, const double* __restrict__ CIt, double2* __restrict__ cT) {
  extern __shared__ __align__(128) double smem[];
  constexpr int MB = TM;              // 8-row blocks of the output (MB * 8 >= nfi)
  const int nv = g.nv, nz = g.nz, nfi = g.nfi, ne = g.ne, no = g.no, pxb = g.pxb;
  const int nks = (ne + 3) >> 2;      // k-steps of 4 rolled velocities (real part; imaginary has no <= ne)
  const int frag_elems = nks * MB * 32;  // one A operand in DMMA fragment order
  const int NC = pxb * nz;
  const int NCp = (NC + 15) & ~15;
  // one slab buffer holds the [nv][NCp] density slab and later the [nz][nfi][pxb] complex staging
  const int buf_elems = (max(nv * NCp, 2 * nz * nfi * pxb) + 15) & ~15;
  double* As_r = smem;                                   // [nks][MB][32]  A fragments, real part
  double* As_i = As_r + frag_elems;                      // [nks][MB][32]  imaginary part
  double* Bbuf0 = As_i + frag_elems;                     // [ngroup][2][buf_elems]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Bbuf0 + 2 * (size_t)g.ngroup * buf_elems);
  const int v0 = g.v0;
  const int xmask = pxb - 1;                             // staging swizzle (pxb is a power of two)
  // The CTA is split into `ngroup` warp groups that run the tile loop independently (own slab
  // buffers, own mbarriers, own named barrier) and share only the A matrices: while one group
  // is in its GEMM phase the other folds / stages / stores, which keeps the FP64 pipe busy.
  const int GT = K1_THREADS / g.ngroup;                  // threads per group
  const int grp = threadIdx.x / GT;
  const int tid = threadIdx.x - grp * GT;                // thread index within the group
  const int warp = tid >> 5, lane = tid & 31;
  double* Bbuf = Bbuf0 + (size_t)grp * 2 * buf_elems;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(GT) : "memory"); };

  for (int i = threadIdx.x; i < frag_elems; i += K1_THREADS) {
    As_r[i] = CRt[i];
    As_i[i] = CIt[i];
  }
  for (int i = threadIdx.x; i < 2 * g.ngroup * buf_elems; i += K1_THREADS) Bbuf0[i] = 0.0;
  const unsigned bar0 = smem_u32(bars) + 16u * grp;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // generic-proxy writes above (zero fill) must be ordered before async-proxy writes (TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  const int nxb = (g.npix_tile + pxb - 1) / pxb;
  const long long ntile = (long long)nxb * g.nt;
  const int nchunk = NCp >> 4;
  const int nitem = nchunk * 2;
  const size_t vstride = (size_t)g.npix_total * nz;      // elements between velocity rows

  // load of one tile's slab into buffer `b`; returns true if it was issued as TMA (uniform)
  auto load_tile = [&](long long tile, int b) -> bool {
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    const int ncol_valid = min(pxb, g.npix_tile - xb * pxb) * nz;
    const double* src = dens + (size_t)t * nv * vstride + (size_t)(g.pix0 + (long long)xb * pxb) * nz;
    double* dst = Bbuf + (size_t)b * buf_elems;
    const bool tma = g.use_tma == 2 || (g.use_tma == 1 && (ncol_valid & 1) == 0);
    if (g.use_tma == 2) {
      // out-of-range columns / rows of the box are zero filled by the TMA unit; columns past
      // this call's pixel range but inside the array are real (finite) data of other pixels
      if (tid == 0) {
        const unsigned bar = bar0 + 8u * b;
        mbar_expect_tx(bar, (unsigned)(nv * NC) * 8u);
        tensor2d_g2s(smem_u32(dst), &tmap, (int)((g.pix0 + (long long)xb * pxb) * nz), t * nv, bar);
      }
    } else if (tma) {
      if (warp == 0) {
        const unsigned bar = bar0 + 8u * b;
        const unsigned row_bytes = (unsigned)ncol_valid * 8u;
        if (lane == 0) mbar_expect_tx(bar, row_bytes * (unsigned)nv);
        __syncwarp();
        for (int v = lane; v < nv; v += 32)
          bulk_g2s(smem_u32(dst + (size_t)v * NCp), src + (size_t)v * vstride, row_bytes, bar);
      }
    } else {
      for (int i = tid; i < nv * NCp; i += GT) {
        const int v = i / NCp, col = i - v * NCp;
        dst[i] = col < ncol_valid ? __ldg(src + (size_t)v * vstride + col) : 0.0;
      }
    }
    return tma;
  };

  unsigned phases = 0u;      // bit b = parity the next TMA completion on buffer b will have
#ifdef PKM_K1_PROF
  long long tick_ = clock64();
#endif
  bool cur_tma = false, next_tma = false;
  const long long tstep = (long long)gridDim.x * g.ngroup;
  long long tile = (long long)blockIdx.x * g.ngroup + grp;
  if (tile < ntile) cur_tma = load_tile(tile, 0);
  for (int it = 0; tile < ntile; tile += tstep, ++it, cur_tma = next_tma) {
    const int b = it & 1;
    // the other buffer was released by the barrier that ended the previous iteration
    if (tile + tstep < ntile) next_tma = load_tile(tile + tstep, b ^ 1);
    if (cur_tma) {
      mbar_wait(bar0 + 8u * b, (phases >> b) & 1u);
      phases ^= 1u << b;
    } else {
      group_sync();
    }
    K1_TICK(0);
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    double* Bs = Bbuf + (size_t)b * buf_elems;
    // phase 1 (self-paired rows carry weight 1/2 in CRt, so e = a + b unconditionally)
    for (int i = tid; i < ne * (NCp >> 1); i += GT) {
      const int u = i / (NCp >> 1), c2 = i - u * (NCp >> 1);
      const int rp = row_plus(u, v0, nv), rm = row_minus(u, v0, nv);
      double2* pa = reinterpret_cast<double2*>(Bs + (size_t)rp * NCp) + c2;
      double2* pb = reinterpret_cast<double2*>(Bs + (size_t)rm * NCp) + c2;
      double2 a = *pa, bb = *pb;
      if (g.is_log) {
        a.x = exp(a.x); a.y = exp(a.y);
        bb.x = exp(bb.x); bb.y = exp(bb.y);
      }
      *pa = make_double2(a.x + bb.x, a.y + bb.y);
      if (rp != rm) *pb = make_double2(a.x - bb.x, a.y - bb.y);
    }
    K1_TICK(1);
    group_sync();
    K1_TICK(2);
    // phase 2: FP64 tensor-core GEMM (DMMA m8n8k4).  One work item per warp = (16-column chunk,
    // real | imaginary part): MB x 2 accumulator tiles of 8x8; per k-step of 4 rolled velocities
    // a lane loads MB A-fragment doubles (pre-arranged, conflict-free) and 2 B-fragment doubles
    // from the folded slab and issues 2*MB DMMAs - 0.6 shared-memory bytes per FMA instead of
    // the 3.4 a register-tiled FMA loop needs (which made the phase shared-memory bound).
    // (the host picks PXB so that there is at most one item per warp: accumulators stay in
    // registers across the barrier that frees the slab for staging)
    double acc[MB][2][2];
    const int item = warp;
    const int part = item & 1, chunk = item >> 1;
    const int fr = lane >> 2, fk = lane & 3;      // fragment row / k index of this lane
    if (item < nitem) {
      const double* A = (part ? As_i : As_r) + lane;
      const int K = part ? no : ne;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        acc[mb][0][0] = acc[mb][0][1] = 0.0;
        acc[mb][1][0] = acc[mb][1][1] = 0.0;
      }
      const double* Bcol = Bs + chunk * 16 + fr;
      for (int ks = 0; ks < nks; ++ks) {
        // E[u] lives in row +u, O[u] (u >= 1) in row -u of the folded slab; u >= K has zero A
        const int u = min(ks * 4 + fk, K - 1);
        const int row = part ? row_minus(u + 1, v0, nv) : row_plus(u, v0, nv);
        const double b0 = Bcol[(size_t)row * NCp];
        const double b1 = Bcol[(size_t)row * NCp + 8];
        double a[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) a[mb] = A[(ks * MB + mb) * 32];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          dmma_m8n8k4(acc[mb][0][0], acc[mb][0][1], a[mb], b0);
          dmma_m8n8k4(acc[mb][1][0], acc[mb][1][1], a[mb], b1);
        }
      }
    }
    K1_TICK(3);
    group_sync();  // every warp of the group is done reading the slab: reuse it as the staging area
    K1_TICK(4);
    // phase 3: Cs[z][m][x] complex; lane holds C[row fr][cols 2 fk, 2 fk + 1] of every 8x8 tile
    double* Cs = Bs;
    if (item < nitem) {
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = chunk * 16 + nb * 8 + 2 * fk + e;
          if (col < NC) {
            const int x = col / nz, z = col - x * nz;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const int m = mb * 8 + fr;
              if (m < nfi) Cs[(((size_t)z * nfi + m) * pxb + (x ^ ((m + (z >> 2)) & xmask))) * 2 + part] = acc[mb][nb][e];
            }
          }
        }
    }
    K1_TICK(5);
    group_sync();
    K1_TICK(6);
    // phase 4: cT[x/32][t*nz + z][m][x%32]
    for (int i = tid; i < nz * nfi * pxb; i += GT) {
      const int zm = i / pxb, xs = i - zm * pxb;
      const int z = zm / nfi, m = zm - z * nfi;
      const int x = xs ^ ((m + (z >> 2)) & xmask);
      const double2 v = reinterpret_cast<const double2*>(Cs)[i];
      const int xg = xb * pxb + x;  // pixel-group-major layout cT[x/32][t*nz+z][m][x%32]
      const size_t ci = (((size_t)(xg >> 5) * g.nt * nz + (size_t)t * nz + z) * nfi + m) * 32 + (xg & 31);
      if (g.out_f32) reinterpret_cast<float2*>(cT)[ci] = make_float2((float)v.x, (float)v.y);
      else cT[ci] = v;
    }
    // order this tile's generic-proxy accesses to the buffer before the TMA refill of it
    K1_TICK(7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    group_sync();
    K1_TICK(8);
  }
}

// ----------------------------------------------------------------------------
// Variant used whenever the slab can be fetched as one 2-D tensor copy: one warp = (metallicity z,
// 8 consecutive pixels), real AND imaginary part.  The 8 columns of a warp's DMMA tile are then 8
// pixels of one (t,z) row of the output layout cT[x/32][tz][m][x%32], and a lane ends up holding
// the complex coefficients of two adjacent pixels: they go from the accumulator registers
// straight to global memory as 16-byte stores.  No staging pass, no second sweep over shared
// memory, two CTA barriers per tile instead of five, and all 12 warps issue DMMAs at the same time.
// Tile = (age t, PXB = 8 or 16 pixels); the slab buffer is free as soon as the GEMM has read it, so
// the tensor copy of tile i+2 is issued there while tile i is being stored.
// ----------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(K1_THREADS, 1)
k_coeff_oct(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap, K1Geom g,
            const double* __restrict__ CRt, const double* __restrict__ CIt,
            const double* __restrict__ exp2_tab) {
  extern __shared__ __align__(128) double smem[];
  constexpr int MB = TM;
  const int nv = g.nv, nz = g.nz, nfi = g.nfi, ne = g.ne, pxb = g.pxb, v0 = g.v0;
  const int nks = (ne + 3) >> 2;
  const int frag_elems = nks * MB * 32;
  const int NC = pxb * nz;                       // density columns of a tile (a multiple of 16)
  // slab pitch: the tensor copy fetches a box two columns wider than the tile (the extra columns
  // belong to the next pixel or are zero filled), because a 96-double pitch puts the four rolled
  // velocities of a B fragment on the same shared-memory banks (8-way conflicts; 4-way with 98)
  const int NCP = g.ncp;
  const int slab_elems = nv * NCP;
  const int buf_elems = (max(nv * NCP, 2 * nfi * NC) + 15) & ~15;   // slab, later the [nz][nfi][2 pxb] staging box
  double* As_r = smem;
  double* As_i = As_r + frag_elems;
  double* Bbuf = As_i + frag_elems;
  double* etab = Bbuf + 2 * (size_t)buf_elems;                           // [PKM_EXP_TAB] 2^(j/256)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(etab + PKM_EXP_TAB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < frag_elems; i += K1_THREADS) {
    As_r[i] = CRt[i];
    As_i[i] = CIt[i];
  }
  for (int i = tid; i < PKM_EXP_TAB; i += K1_THREADS) etab[i] = exp2_tab[i];
  const unsigned bar0 = smem_u32(bars);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nxb = (g.npix_tile + pxb - 1) / pxb;
  const long long ntile = (long long)nxb * g.nt;
  const long long tstep = gridDim.x;
  auto issue = [&](long long tile, int b) {      // thread 0 only
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    const unsigned bar = bar0 + 8u * b;
    mbar_expect_tx(bar, (unsigned)slab_elems * 8u);
    tensor2d_g2s(smem_u32(Bbuf + (size_t)b * buf_elems), &tmap,
                 (int)((g.pix0 + (long long)xb * pxb) * nz), t * nv, bar);
  };
  long long tile = blockIdx.x;
  if (tid == 0) {
    if (tile < ntile) issue(tile, 0);
    if (tile + tstep < ntile) issue(tile + tstep, 1);
  }
  // this warp's GEMM item and this thread's slice of the fold (both tile independent)
  const int nitem = nz * (pxb >> 3);
  const int iz = warp % nz, oct = warp / nz;
  const int fr = lane >> 2, fk = lane & 3;

  unsigned phases = 0u;
#ifdef PKM_K1_PROF
  const int grp = 0;
  long long tick_ = clock64();
#endif
  for (int it = 0; tile < ntile; tile += tstep, ++it) {
    const int b = it & 1;
    mbar_wait(bar0 + 8u * b, (phases >> b) & 1u);
    phases ^= 1u << b;
    K1_TICK(0);
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    double* Bs = Bbuf + (size_t)b * buf_elems;
    // the other buffer holds the previous tile's staged coefficients: once the store engine has
    // read them the slab of the next tile streams in behind this tile's GEMM
    if (tid == 0 && it > 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (tile + tstep < ntile) issue(tile + tstep, b ^ 1);
    }
    // GEMM on the FP64 tensor cores: C[m = mb*8 + fr][pixels 2 fk, 2 fk + 1] of this warp's 8 pixels
    // at metallicity iz, real and imaginary part.  The v <-> -v fold of the real DFT (e = sum,
    // o = difference of the rolled profile; halves the GEMM) and exp(log p) happen HERE, while the
    // B fragments are formed: every slab element belongs to exactly one (lane, k-step) of exactly
    // one warp, so nothing is evaluated twice, the slab stays read-only and the separate fold pass
    // (4.3k of 16.6k cycles per tile, one CTA barrier) is gone.  Even and odd part use the same
    // rolled index u (the odd operand has a zero column for u = 0).
    double accr[MB][2], acci[MB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) accr[mb][0] = accr[mb][1] = acci[mb][0] = acci[mb][1] = 0.0;
    if (warp < nitem) {
      const double* Bcol = Bs + (oct * 8 + fr) * nz + iz;
      const double* Ar = As_r + lane;
      const double* Ai = As_i + lane;
      auto fetch = [&](int ks, double& pa, double& pb) {
        const int u = min(ks * 4 + fk, ne - 1);               // u >= ne has zero A
        pa = Bcol[(size_t)row_plus(u, v0, nv) * NCP];
        pb = Bcol[(size_t)row_minus(u, v0, nv) * NCP];
      };
      double pa, pb;
      fetch(0, pa, pb);
#if PKM_K1_APREFETCH
      // the A fragments of the next k-step are loaded while this step's DMMAs issue
      double ar[MB], ai[MB];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        ar[mb] = Ar[mb * 32];
        ai[mb] = Ai[mb * 32];
      }
#endif
      for (int ks = 0; ks < nks; ++ks) {
        double qa = pa, qb = pb;
        if (ks + 1 < nks) fetch(ks + 1, pa, pb);               // next step's raw values, ahead of the DMMAs
        if (g.is_log) {
          qa = exp_tab(qa, etab);
          qb = exp_tab(qb, etab);
        }
        const double be = qa + qb, bo = qa - qb;
#if PKM_K1_APREFETCH
        double nr[MB], ni[MB];
        const int kn = min(ks + 1, nks - 1);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          nr[mb] = Ar[(kn * MB + mb) * 32];
          ni[mb] = Ai[(kn * MB + mb) * 32];
        }
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          dmma_m8n8k4(accr[mb][0], accr[mb][1], ar[mb], be);
          dmma_m8n8k4(acci[mb][0], acci[mb][1], ai[mb], bo);
        }
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          ar[mb] = nr[mb];
          ai[mb] = ni[mb];
        }
#else
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          dmma_m8n8k4(accr[mb][0], accr[mb][1], Ar[(ks * MB + mb) * 32], be);
          dmma_m8n8k4(acci[mb][0], acci[mb][1], Ai[(ks * MB + mb) * 32], bo);
        }
#endif
      }
    }
    K1_TICK(3);
    __syncthreads();   // every warp is done reading the slab: it becomes the staging box
    K1_TICK(4);
    // stage as the output box [z][m][2 pxb] (pixel fastest, re/im interleaved): a lane owns the
    // complex coefficients of two adjacent pixels = 32 contiguous bytes
    if (warp < nitem) {
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int m = mb * 8 + fr;
        if (m < nfi) {
          const size_t o = ((size_t)iz * nfi + m) * (2 * pxb) + oct * 16 + 4 * fk;
          if (g.out_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(Bs) + o) =
                make_float4((float)accr[mb][0], (float)acci[mb][0], (float)accr[mb][1], (float)acci[mb][1]);
          } else {
            *reinterpret_cast<double2*>(Bs + o) = make_double2(accr[mb][0], acci[mb][0]);
            *reinterpret_cast<double2*>(Bs + o + 2) = make_double2(accr[mb][1], acci[mb][1]);
          }
        }
      }
    }
    // generic-proxy writes -> async-proxy read: one 4-D tensor store moves the whole box to
    // cT[x/32][t*nz + z][m][x%32]; it drains behind the next tile's fold and GEMM
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const int xg0 = xb * pxb;
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                   ::"l"(reinterpret_cast<unsigned long long>(&omap)), "r"(2 * (xg0 & 31)), "r"(0),
                     "r"(t * nz), "r"(xg0 >> 5), "r"(smem_u32(Bs)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    K1_TICK(5);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------
// Barrier-free form of the octet kernel (the default).  Same tiles, same warp items, same fused
// fold / exp, but no CTA-wide barrier inside the tile loop:
//   * a PRODUCER warp issues the 2-D tensor copies: slab buffer b is refilled as soon as every
//     consumer warp has arrived on empty[b] (mbarrier, one arrival per warp item) - the consumer
//     side is the usual full[b] complete_tx wait;
//   * a consumer warp releases the slab right after its last B fragment and then writes its
//     accumulators straight to cT[x/32][t*nz + z][m][x%32] with 256-bit stores (a lane holds the
//     complex coefficients of two adjacent pixels = one 32-byte sector; a warp store covers eight
//     full 128-byte lines) - no staging box, so the slab is read-only for its whole life.
// The warps of a CTA drift apart by up to one tile: one warp's stores, slab wait and exp chains run
// under the other warps' DMMAs instead of every warp idling at two barriers per tile.
// ----------------------------------------------------------------------------
// Two table exponentials (exp_tab.cuh: same constants, same operation order, same results) cut
// into stages so that the caller can interleave them with independent work.  The range / NaN
// handling is one integer test on the high word (|x| >= 708 or NaN takes the rare slow path)
// instead of three FP64 compares per value.
// (The FP64 operations are volatile asm statements: that pins their order relative to the DMMAs,
// which the compiler otherwise clusters.)
__device__ __forceinline__ double vfma(double a, double b, double c) {
  double d;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}
__device__ __forceinline__ double vmul(double a, double b) {
  double d;
  asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
  return d;
}
__device__ __forceinline__ double vadd(double a, double b) {
  double d;
  asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
  return d;
}
struct ExpPair {
  double xa, xb, ya, yb;
  double ta, tb, fa, fb, ra, rb, qa, qb, sa, sb, Ta, Tb;
  __device__ __forceinline__ void stage(int st, const double* __restrict__ tab) {
    switch (st) {
      case 0:
        ta = vfma(xa, 369.32993046757463228, 6755399441055744.0);
        tb = vfma(xb, 369.32993046757463228, 6755399441055744.0);
        fa = vadd(ta, -6755399441055744.0);
        fb = vadd(tb, -6755399441055744.0);
        break;
      case 1:
        ra = vfma(fa, -2.70760617331688990816e-03, xa);
        rb = vfma(fb, -2.70760617331688990816e-03, xb);
        Ta = tab[__double2loint(ta) & (PKM_EXP_TAB - 1)];
        Tb = tab[__double2loint(tb) & (PKM_EXP_TAB - 1)];
        break;
      case 2:
        ra = vfma(fa, -7.45396456746323320321e-13, ra);
        rb = vfma(fb, -7.45396456746323320321e-13, rb);
        break;
      case 3:
        qa = vfma(ra, 1.0 / 24.0, 1.0 / 6.0);
        qb = vfma(rb, 1.0 / 24.0, 1.0 / 6.0);
        sa = vmul(ra, ra);
        sb = vmul(rb, rb);
        break;
      case 4:
        qa = vfma(qa, ra, 0.5);
        qb = vfma(qb, rb, 0.5);
        break;
      case 5:
        qa = vfma(sa, qa, ra);                   // e^r - 1
        qb = vfma(sb, qb, rb);
        break;
      default: {
        qa = vfma(Ta, qa, Ta);                   // in [1, 2.01)
        qb = vfma(Tb, qb, Tb);
        const int ka = __double2loint(ta), kb = __double2loint(tb);
        ya = __hiloint2double(__double2hiint(qa) + ((ka >> 8) << 20), __double2loint(qa));
        yb = __hiloint2double(__double2hiint(qb) + ((kb >> 8) << 20), __double2loint(qb));
        // |x| >= 708, inf or NaN: rare, and a warp-uniform branch keeps the FP64 compares out of
        // the common path
        const bool rare = max(__double2hiint(xa) & 0x7fffffff, __double2hiint(xb) & 0x7fffffff) >= 0x40862000;
        if (__any_sync(0xffffffffu, rare)) {
          ya = xa < -708.0 ? 0.0 : (xa > 709.0 ? INFINITY : (xa != xa ? xa : ya));
          yb = xb < -708.0 ? 0.0 : (xb > 709.0 ? INFINITY : (xb != xb ? xb : yb));
        }
        break;
      }
    }
  }
};

constexpr int K1F_THREADS = K1_THREADS + 32;   // 12 consumer warps + the producer warp
template <int TM, bool IS_LOG, bool APF>
__global__ void __launch_bounds__(APF ? K1_THREADS : K1F_THREADS, 1)
k_coeff_flow(const __grid_constant__ CUtensorMap tmap, K1Geom g, const double* __restrict__ CRt,
             const double* __restrict__ CIt, const double* __restrict__ exp2_tab,
             double2* __restrict__ cT) {
  extern __shared__ __align__(128) double smem[];
  constexpr int MB = TM;
  const int nv = g.nv, nz = g.nz, nfi = g.nfi, ne = g.ne, pxb = g.pxb, v0 = g.v0;
  const int nks = (ne + 3) >> 2;
  const int frag_elems = nks * MB * 32;
  const int NCP = g.ncp;                         // slab pitch (see k_coeff_oct)
  const int slab_elems = nv * NCP;
  const int buf_elems = (slab_elems + 15) & ~15;
  double* As_r = smem;
  double* As_i = As_r + frag_elems;
  double* Bbuf = As_i + frag_elems;
  double* etab = Bbuf + 2 * (size_t)buf_elems;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(etab + PKM_EXP_TAB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nitem = nz * (pxb >> 3);
  constexpr int NTHR = APF ? K1_THREADS : K1F_THREADS;
  for (int i = tid; i < frag_elems; i += NTHR) {
    As_r[i] = CRt[i];
    As_i[i] = CIt[i];
  }
  for (int i = tid; i < PKM_EXP_TAB; i += NTHR) etab[i] = exp2_tab[i];
  const unsigned full0 = smem_u32(bars), empty0 = full0 + 16u;
  const unsigned cnt0 = full0 + 32u;             // APF: arrival counters (the last arriver refills)
  if (tid == 0) {
    mbar_init(full0, 1);
    mbar_init(full0 + 8, 1);
    mbar_init(empty0, (unsigned)nitem);
    mbar_init(empty0 + 8, (unsigned)nitem);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt0), "r"(0u) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt0 + 4u), "r"(0u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nxb = (g.npix_tile + pxb - 1) / pxb;
  const long long ntile = (long long)nxb * g.nt;
  const long long tstep = gridDim.x;

  auto issue_tile = [&](long long tile, int b) {  // one thread
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    mbar_expect_tx(full0 + 8u * b, (unsigned)slab_elems * 8u);
    tensor2d_g2s(smem_u32(Bbuf + (size_t)b * buf_elems), &tmap,
                 (int)((g.pix0 + (long long)xb * pxb) * nz), t * nv, full0 + 8u * b);
  };
  if (APF) {
    // no producer warp (12 warps = 3 per scheduler leave 168 registers per thread for the second
    // A-fragment set): thread 0 issues the first two slabs, later the LAST consumer warp to
    // release a slab refills it
    if (tid == 0) {
      if ((long long)blockIdx.x < ntile) issue_tile(blockIdx.x, 0);
      if ((long long)blockIdx.x + tstep < ntile) issue_tile(blockIdx.x + tstep, 1);
    }
  } else if (warp == K1_WARPS) {                 // ---- producer ----
    if (lane == 0) {
      long long tile = blockIdx.x;
      for (int it = 0; tile < ntile; tile += tstep, ++it) {
        const int b = it & 1;
        if (it >= 2) mbar_wait(empty0 + 8u * b, (unsigned)((it >> 1) - 1) & 1u);
        const int t = (int)(tile / nxb);
        const int xb = (int)(tile - (long long)t * nxb);
        mbar_expect_tx(full0 + 8u * b, (unsigned)slab_elems * 8u);
        tensor2d_g2s(smem_u32(Bbuf + (size_t)b * buf_elems), &tmap,
                     (int)((g.pix0 + (long long)xb * pxb) * nz), t * nv, full0 + 8u * b);
      }
    }
    return;
  }
  if (warp >= nitem) return;                     // fewer items than consumer warps

  // ---- consumers: warp = (metallicity iz, pixel octet oct) ----
  const int iz = warp % nz, oct = warp / nz;
  const int fr = lane >> 2, fk = lane & 3;
  const double* Ar = As_r + lane;
  const double* Ai = As_i + lane;
  const size_t ntz = (size_t)g.nt * nz;
  long long tile = blockIdx.x;
  for (int it = 0; tile < ntile; tile += tstep, ++it) {
    const int b = it & 1;
    mbar_wait(full0 + 8u * b, (unsigned)(it >> 1) & 1u);
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    const double* Bcol = Bbuf + (size_t)b * buf_elems + (oct * 8 + fr) * nz + iz;
    double accr[MB][2], acci[MB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) accr[mb][0] = accr[mb][1] = acci[mb][0] = acci[mb][1] = 0.0;
    auto fetch = [&](int ks, double& pa, double& pb) {
      const int u = min(ks * 4 + fk, ne - 1);               // u >= ne has zero A
      pa = Bcol[(size_t)row_plus(u, v0, nv) * NCP];
      pb = Bcol[(size_t)row_minus(u, v0, nv) * NCP];
    };
    // Software pipeline: raw values two k-steps ahead; the exp / fold of k-step ks+1 is cut into
    // seven stages that are issued BETWEEN the DMMA pairs of k-step ks (DMMA and DFMA share the
    // FP64 pipe: left to the compiler the two dependent exp chains - ~10 FP64 latencies - sit
    // after the 14 DMMAs, where nothing of this warp covers them).
    double pa, pb, be, bo;
    fetch(0, pa, pb);
    if (IS_LOG) {
      pa = exp_tab(pa, etab);
      pb = exp_tab(pb, etab);
    }
    be = pa + pb;
    bo = pa - pb;
    fetch(nks > 1 ? 1 : 0, pa, pb);
    constexpr bool is_log = IS_LOG;
    // APF: the A fragments of k-step ks+1 are loaded into a second register set while the DMMAs
    // of k-step ks issue (no load -> DMMA -> reload chain on the same registers); the k-loop is
    // unrolled by two by hand so that both sets stay in registers
    double f0r[MB], f0i[MB], f1r[MB], f1i[MB];
    if (APF) {
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        f0r[mb] = Ar[mb * 32];
        f0i[mb] = Ai[mb * 32];
      }
    }
    auto kstep = [&](int ks, double (&cr)[MB], double (&ci)[MB], double (&nr)[MB], double (&ni)[MB]) {
      ExpPair ep;
      ep.xa = pa;
      ep.xb = pb;
      double nbe = 0.0, nbo = 0.0;
      const int kn = min(ks + 1, nks - 1);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (APF) {
          dmma_m8n8k4(accr[mb][0], accr[mb][1], cr[mb], be);
          dmma_m8n8k4(acci[mb][0], acci[mb][1], ci[mb], bo);
          nr[mb] = Ar[(kn * MB + mb) * 32];
          ni[mb] = Ai[(kn * MB + mb) * 32];
        } else {
          dmma_m8n8k4(accr[mb][0], accr[mb][1], Ar[(ks * MB + mb) * 32], be);
          dmma_m8n8k4(acci[mb][0], acci[mb][1], Ai[(ks * MB + mb) * 32], bo);
        }
        // stages [mb * 7 / MB, (mb + 1) * 7 / MB) of the next step's B fragment
#pragma unroll
        for (int st = (mb * 7) / MB; st < ((mb + 1) * 7) / MB; ++st) {
          if (st == 0 && ks + 2 < nks) fetch(ks + 2, pa, pb);
          if (is_log) ep.stage(st, etab);
          if (st == 6) {
            const double ya = is_log ? ep.ya : ep.xa, yb = is_log ? ep.yb : ep.xb;
            nbe = vadd(ya, yb);
            nbo = vadd(ya, -yb);
          }
        }
      }
      be = nbe;
      bo = nbo;
    };
    if (APF) {
      int ks = 0;
      for (; ks + 1 < nks; ks += 2) {
        kstep(ks, f0r, f0i, f1r, f1i);
        kstep(ks + 1, f1r, f1i, f0r, f0i);
      }
      if (ks < nks) kstep(ks, f0r, f0i, f1r, f1i);
    } else {
      for (int ks = 0; ks < nks; ++ks) kstep(ks, f0r, f0i, f1r, f1i);
    }
    // the slab is consumed: one arrival per warp lets the producer refill it
    __syncwarp();
    if (APF) {
      if (lane == 0) {
        unsigned old;
        asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt0 + 4u * b) : "memory");
        if (old == (unsigned)nitem - 1u) {
          asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(cnt0 + 4u * b), "r"(0u) : "memory");
          if (tile + 2 * tstep < ntile) issue_tile(tile + 2 * tstep, b);
        }
      }
    } else if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8u * b) : "memory");
    }
    const int xg0 = xb * pxb + oct * 8;
    const size_t base = (((size_t)(xg0 >> 5) * ntz + (size_t)t * nz + iz) * nfi) * 32 + (xg0 & 31) + 2 * fk;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int m = mb * 8 + fr;
      if (m < nfi) {
        const size_t ci = base + (size_t)m * 32;
        if (g.out_f32)
          *reinterpret_cast<float4*>(reinterpret_cast<float2*>(cT) + ci) =
              make_float4((float)accr[mb][0], (float)acci[mb][0], (float)accr[mb][1], (float)acci[mb][1]);
        else
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(cT + ci), "d"(accr[mb][0]),
                       "d"(acci[mb][0]), "d"(accr[mb][1]), "d"(acci[mb][1]) : "memory");
      }
    }
  }
}

// A operand in DMMA fragment order: element [ks][mb][lane] = C[m = mb*8 + lane/4][u = ks*4 + lane%4]
// (zero outside the matrix), so a warp's fragment load is one contiguous 256-byte read.
// `shift` = 1 stores column u of the fragment table from matrix column u - 1 (zero for u = 0): the
// octet kernel indexes the odd part O[u] with the same rolled position u as the even part.
static void upload_fragments(const std::vector<double>& C, int nfi, int K, int MB, int nks, double** d_out,
                             int shift = 0) {
  std::vector<double> h((size_t)nks * MB * 32, 0.0);
  for (int ks = 0; ks < nks; ++ks)
    for (int mb = 0; mb < MB; ++mb)
      for (int lane = 0; lane < 32; ++lane) {
        const int m = mb * 8 + lane / 4, u = ks * 4 + lane % 4 - shift;
        if (m < nfi && u >= 0 && u < K) h[((size_t)ks * MB + mb) * 32 + lane] = C[(size_t)m * K + u];
      }
  PKM_CUDA(cudaMalloc(d_out, h.size() * sizeof(double)));
  PKM_CUDA(cudaMemcpy(*d_out, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
}

void density_tables_upload(pkm_plan* pl) {
  pl->TM = (pl->nfi + 7) / 8;   // number of 8-row blocks
  // plain row-major copies for the generic kernel (nfi > 64, or PKM_COEFF_GENERIC=1 for testing)
  PKM_CUDA(cudaMalloc(&pl->d_CR, std::max<size_t>(pl->fold.CR.size(), 1) * sizeof(double)));
  PKM_CUDA(cudaMemcpy(pl->d_CR, pl->fold.CR.data(), pl->fold.CR.size() * sizeof(double), cudaMemcpyHostToDevice));
  PKM_CUDA(cudaMalloc(&pl->d_CI, std::max<size_t>(pl->fold.CI.size(), 1) * sizeof(double)));
  PKM_CUDA(cudaMemcpy(pl->d_CI, pl->fold.CI.data(), pl->fold.CI.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (pl->TM > 8) return;       // the tensor-core kernels keep the operands resident in shared memory: nfi <= 64
  const int nks = (pl->fold.ne + 3) / 4;
  pl->nfiP = nks * pl->TM * 32;  // doubles per A operand
  upload_fragments(pl->fold.CR, pl->nfi, pl->fold.ne, pl->TM, nks, &pl->d_CRt);
  upload_fragments(pl->fold.CI, pl->nfi, pl->fold.no, pl->TM, nks, &pl->d_CIt);
  upload_fragments(pl->fold.CI, pl->nfi, pl->fold.no, pl->TM, nks, &pl->d_CIt_oct, 1);
}

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    looked_up = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    else
      cudaGetLastError();
  }
  return encode;
}

// [rows][cols] float64 matrix view of the density, box = box_rows x box_cols (tiled, no swizzle)
static bool encode_density_map(CUtensorMap* map, const double* base, int64_t cols, int64_t rows,
                               int box_cols, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (!encode) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static size_t coeff_smem_bytes(const pkm_plan* pl, int pxb, int ngroup) {
  const int NC = pxb * pl->d.nz, NCp = (NC + 15) & ~15;
  const size_t buf_elems = (std::max<size_t>((size_t)pl->d.nv * NCp, (size_t)2 * pl->d.nz * pl->nfi * pxb) + 15) & ~size_t(15);
  size_t dbl = (size_t)2 * pl->nfiP + 2 * ngroup * buf_elems;
  return dbl * sizeof(double) + 2 * ngroup * sizeof(unsigned long long);
}

// Tile shape: prefer two warp groups per CTA (6 warps each, pipelines overlap), then one group
// of 12 warps; within each, the largest pixel block of {16, 8, 4, 2} whose slab buffers fit
// shared memory and whose GEMM items (2 per 16 columns) do not exceed one per warp of a group.
int coeff_choose_pxb(const pkm_plan* pl, int* ngroup_out) {
  const char* force = getenv("PKM_COEFF_GROUPS");
  for (int ngroup : {2, 1}) {
    if (force && atoi(force) != ngroup) continue;
    for (int pxb : {16, 8, 4, 2}) {
      const int NCp = (pxb * pl->d.nz + 15) & ~15;
      if ((NCp >> 4) * 2 > K1_WARPS / ngroup) continue;
      if (coeff_smem_bytes(pl, pxb, ngroup) <= 227 * 1024) {
        if (ngroup_out) *ngroup_out = ngroup;
        return pxb;
      }
    }
  }
  PKM_THROW(PKM_ERR_INVALID, "coefficient kernel: nv=%d nz=%d does not fit shared memory", pl->d.nv, pl->d.nz);
}

template <int TM>
static void launch_coeff_t(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total,
                           int64_t pix0, int npix_tile, int xpad, int pxb, int ngroup, double2* cT,
                           cudaStream_t st) {
  size_t smem = coeff_smem_bytes(pl, pxb, ngroup);
  PKM_CUDA(cudaFuncSetAttribute(k_coeff<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t ntile = (int64_t)((npix_tile + pxb - 1) / pxb) * pl->d.nt;
  int grid = (int)std::min<int64_t>((ntile + ngroup - 1) / ngroup, pl->num_sms);
  K1Geom g;
  g.ngroup = ngroup;
  g.npix_total = npix_total; g.pix0 = pix0; g.npix_tile = npix_tile; g.xpad = xpad;
  g.nv = pl->d.nv; g.nz = pl->d.nz; g.nfi = pl->nfi; g.ne = pl->fold.ne; g.no = pl->fold.no;
  g.pxb = pxb; g.nt = pl->d.nt; g.is_log = is_log; g.v0 = pl->d.v0_idx;
  g.out_f32 = pl->d.precision == 1;
  // 1-D bulk copies need 16-byte aligned row starts and sizes; the tensor map needs a 16-byte
  // aligned base and row pitch and writes a dense [nv][NC] box (so NC must equal the padded pitch)
  const int nz = pl->d.nz;
  const int NC = pxb * nz;
  const bool base_ok = reinterpret_cast<uintptr_t>(dens) % 16 == 0 && (npix_total * nz) % 2 == 0;
  g.use_tma = (base_ok && (pix0 * nz) % 2 == 0 && NC % 2 == 0) ? 1 : 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (base_ok && NC % 16 == 0 && NC <= 256 && pl->d.nv <= 256 && npix_total * nz < (1LL << 31) &&
      encode_density_map(&tmap, dens, npix_total * nz, (int64_t)pl->d.nt * pl->d.nv, NC, pl->d.nv))
    g.use_tma = 2;
  if (const char* e = getenv("PKM_COEFF_TMA_MODE")) g.use_tma = std::min(g.use_tma, atoi(e));
  KernelTimer tm(pl, PKM_K_COEFF, st);
  k_coeff<TM><<<grid, K1_THREADS, smem, st>>>(dens, tmap, g, pl->d_CRt, pl->d_CIt, cT);
  PKM_CUDA(cudaGetLastError());
#ifdef PKM_K1_PROF
  {
    unsigned long long h[32];
    PKM_CUDA(cudaStreamSynchronize(st));
    PKM_CUDA(cudaMemcpyFromSymbol(h, k1_prof, sizeof(h)));
    const long long tiles_per_group = (ntile / ngroup + grid - 1) / grid;
    for (int gi = 0; gi < ngroup; ++gi) {
      fprintf(stderr, "k1prof grp%d tiles~%lld cyc/tile:", gi, tiles_per_group);
      for (int k = 0; k < 9; ++k) fprintf(stderr, " %llu", h[gi * 16 + k] / (unsigned long long)std::max<long long>(tiles_per_group, 1));
      fprintf(stderr, "\n");
    }
    memset(h, 0, sizeof(h));
    PKM_CUDA(cudaMemcpyToSymbol(k1_prof, h, sizeof(h)));
  }
#endif
  pl->launches++;
}

#ifndef PKM_K1_APREFETCH
#define PKM_K1_APREFETCH 0
#endif
#ifndef PKM_K1_PITCH_PAD
#define PKM_K1_PITCH_PAD 2
#endif
static size_t coeff_oct_smem(const pkm_plan* pl, int NC) {
  const size_t buf_elems = (std::max<size_t>((size_t)pl->d.nv * (NC + PKM_K1_PITCH_PAD), (size_t)2 * pl->nfi * NC) + 15) & ~size_t(15);
  return ((size_t)2 * pl->nfiP + 2 * buf_elems + PKM_EXP_TAB) * sizeof(double) + 16;
}

static size_t coeff_flow_smem(const pkm_plan* pl, int NC) {
  const size_t buf_elems = ((size_t)pl->d.nv * (NC + PKM_K1_PITCH_PAD) + 15) & ~size_t(15);
  return ((size_t)2 * pl->nfiP + 2 * buf_elems + PKM_EXP_TAB) * sizeof(double) + 48;
}

// 4-D view of the coefficient array cT[x/32][t*nz+z][m][x%32] (complex) in scalar elements:
// dims {64 = 32 pixels x re/im, nfi, nt*nz, xpad/32}; box = {2 pxb, nfi, nz, 1} = one tile
static bool encode_coeff_map(CUtensorMap* map, void* cT, bool f32, int nfi, int ntz, int ngroup32,
                             int pxb, int nz) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (!encode) return false;
  const size_t es = f32 ? sizeof(float) : sizeof(double);
  cuuint64_t dims[4] = {64, (cuuint64_t)nfi, (cuuint64_t)ntz, (cuuint64_t)ngroup32};
  cuuint64_t strides[3] = {64 * es, (cuuint64_t)nfi * 64 * es, (cuuint64_t)ntz * nfi * 64 * es};
  cuuint32_t box[4] = {(cuuint32_t)(2 * pxb), (cuuint32_t)nfi, (cuuint32_t)nz, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, cT,
                      dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Pixel block of the octet variant: 16 or 8 pixels such that there are 6..12 (z, octet) items
// (one per warp), the box is a legal tensor-map box and two slab buffers fit shared memory.
static int coeff_oct_pxb(const pkm_plan* pl) {
  const char* force = getenv("PKM_COEFF_KERNEL");      // "0" = always the staged kernel
  if (force && atoi(force) == 0) return 0;
  const char* mode = getenv("PKM_COEFF_TMA_MODE");     // the fallbacks under test need the staged kernel
  if (mode && atoi(mode) < 2) return 0;
  for (int pxb : {16, 8}) {
    const int NC = pxb * pl->d.nz, nitem = pl->d.nz * (pxb / 8);
    if (NC % 16 != 0 || NC > 256 || pl->d.nv > 256 || nitem > K1_WARPS || nitem < 6) continue;
    const size_t bytes = coeff_oct_smem(pl, NC);
    if (bytes <= 227 * 1024) return pxb;
  }
  return 0;
}

template <int TM>
static bool launch_coeff_oct_t(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total,
                               int64_t pix0, int npix_tile, int xpad, double2* cT, cudaStream_t st) {
  const int pxb = coeff_oct_pxb(pl);
  const int nz = pl->d.nz, nv = pl->d.nv;
  if (!pxb) return false;
  const int NC = pxb * nz;
  const bool base_ok = reinterpret_cast<uintptr_t>(dens) % 16 == 0 && (npix_total * nz) % 2 == 0 &&
                       npix_total * nz < (1LL << 31);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!base_ok || !encode_density_map(&tmap, dens, npix_total * nz, (int64_t)pl->d.nt * nv, NC + PKM_K1_PITCH_PAD, nv))
    return false;
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  if (nz > 256 || pl->nfi > 256 ||
      !encode_coeff_map(&omap, cT, pl->d.precision == 1, pl->nfi, pl->d.nt * nz, xpad / 32, pxb, nz))
    return false;
  const size_t smem = coeff_oct_smem(pl, NC);
  PKM_CUDA(cudaFuncSetAttribute(k_coeff_oct<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntile = (int64_t)((npix_tile + pxb - 1) / pxb) * pl->d.nt;
  const int grid = (int)std::min<int64_t>(ntile, pl->num_sms);
  K1Geom g;
  g.ngroup = 1;
  g.npix_total = npix_total; g.pix0 = pix0; g.npix_tile = npix_tile; g.xpad = xpad;
  g.nv = nv; g.nz = nz; g.nfi = pl->nfi; g.ne = pl->fold.ne; g.no = pl->fold.no;
  g.pxb = pxb; g.nt = pl->d.nt; g.is_log = is_log; g.v0 = pl->d.v0_idx;
  g.out_f32 = pl->d.precision == 1;
  g.use_tma = 2;
  g.ncp = NC + PKM_K1_PITCH_PAD;
  KernelTimer tm(pl, PKM_K_COEFF, st);
  const char* flow = getenv("PKM_COEFF_FLOW");          // "0" = the staged-store (TMA tensor store) form
  if (!flow || atoi(flow) != 0) {
    const size_t fsmem = coeff_flow_smem(pl, NC);
    const char* apf_env = getenv("PKM_COEFF_APF");      // "0" = producer-warp form without the second A-fragment set
    const bool apf = !(apf_env && atoi(apf_env) == 0 && apf_env[0] != '\0');
    auto kern = is_log ? (apf ? k_coeff_flow<TM, true, true> : k_coeff_flow<TM, true, false>)
                       : (apf ? k_coeff_flow<TM, false, true> : k_coeff_flow<TM, false, false>);
    PKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    kern<<<grid, apf ? K1_THREADS : K1F_THREADS, fsmem, st>>>(tmap, g, pl->d_CRt, pl->d_CIt_oct, pl->d_exp_tab, cT);
  } else {
    k_coeff_oct<TM><<<grid, K1_THREADS, smem, st>>>(tmap, omap, g, pl->d_CRt, pl->d_CIt_oct, pl->d_exp_tab);
  }
  PKM_CUDA(cudaGetLastError());
#ifdef PKM_K1_PROF
  {
    unsigned long long h[32];
    PKM_CUDA(cudaStreamSynchronize(st));
    PKM_CUDA(cudaMemcpyFromSymbol(h, k1_prof, sizeof(h)));
    const long long tiles = std::max<long long>((ntile + grid - 1) / grid, 1);
    fprintf(stderr, "k1oct pxb=%d tiles~%lld cyc/tile:", pxb, tiles);
    for (int k = 0; k < 6; ++k) fprintf(stderr, " %llu", h[k] / (unsigned long long)tiles);
    fprintf(stderr, "\n");
    memset(h, 0, sizeof(h));
    PKM_CUDA(cudaMemcpyToSymbol(k1_prof, h, sizeof(h)));
  }
#endif
  pl->launches++;
  return true;
}

// ----------------------------------------------------------------------------
// Generic coefficient kernel for velocity grids whose folded matrices do not fit the resident-
// operand tensor-core kernels above (nfi = nv/2 + 1 > 64, e.g. IFUCube's default nv = 200/201):
// plain FP64 FMAs, the folded matrices read through L1 as warp-uniform loads.  Any shape, any
// alignment; several times slower per column than the DMMA kernels - a correctness path.
// CTA = (age t, 64 consecutive density columns); thread = (column, group of 4 output rows).
// ----------------------------------------------------------------------------
constexpr int KG_COLS = 64, KG_ROWS = 4;
__global__ void __launch_bounds__(256)
k_coeff_generic(const double* __restrict__ dens, K1Geom g, const double* __restrict__ CR,
                const double* __restrict__ CI, double2* __restrict__ cT) {
  extern __shared__ double kg_smem[];          // e [ne][64] | o [no][64]
  const int nv = g.nv, nz = g.nz, nfi = g.nfi, ne = g.ne, no = g.no, v0 = g.v0;
  double* se = kg_smem;
  double* so = kg_smem + (size_t)ne * KG_COLS;
  const int t = blockIdx.y;
  const long long col0 = (long long)blockIdx.x * KG_COLS;
  const long long ncol = (long long)g.npix_tile * nz;
  const size_t vstride = (size_t)g.npix_total * nz;
  const double* src = dens + (size_t)t * nv * vstride + (size_t)g.pix0 * nz + col0;
  for (int i = threadIdx.x; i < ne * KG_COLS; i += blockDim.x) {
    const int u = i / KG_COLS, c = i - u * KG_COLS;
    double a = 0.0, b = 0.0;
    if (col0 + c < ncol) {
      a = __ldg(src + (size_t)row_plus(u, v0, nv) * vstride + c);
      b = __ldg(src + (size_t)row_minus(u, v0, nv) * vstride + c);
      if (g.is_log) { a = exp(a); b = exp(b); }
    }
    se[i] = a + b;                              // self-paired rows carry weight 1/2 in CR
    if (u >= 1 && u - 1 < no) so[(size_t)(u - 1) * KG_COLS + c] = a - b;
  }
  __syncthreads();
  const int c = threadIdx.x % KG_COLS, rg = threadIdx.x / KG_COLS;
  const long long col = col0 + c;
  if (col >= ncol) return;
  const int x = (int)(col / nz), z = (int)(col - (long long)x * nz);
  const size_t obase = (((size_t)(x >> 5) * g.nt * nz + (size_t)t * nz + z) * nfi) * 32 + (x & 31);
  for (int m0 = rg * KG_ROWS; m0 < nfi; m0 += KG_ROWS * (256 / KG_COLS)) {
    double ar[KG_ROWS], ai[KG_ROWS];
#pragma unroll
    for (int r = 0; r < KG_ROWS; ++r) ar[r] = ai[r] = 0.0;
    for (int u = 0; u < ne; ++u) {
      const double e = se[(size_t)u * KG_COLS + c];
#pragma unroll
      for (int r = 0; r < KG_ROWS; ++r)
        if (m0 + r < nfi) ar[r] = fma(__ldg(CR + (size_t)(m0 + r) * ne + u), e, ar[r]);
    }
    for (int u = 0; u < no; ++u) {
      const double o = so[(size_t)u * KG_COLS + c];
#pragma unroll
      for (int r = 0; r < KG_ROWS; ++r)
        if (m0 + r < nfi) ai[r] = fma(__ldg(CI + (size_t)(m0 + r) * no + u), o, ai[r]);
    }
#pragma unroll
    for (int r = 0; r < KG_ROWS; ++r) {
      if (m0 + r >= nfi) continue;
      const size_t ci = obase + (size_t)(m0 + r) * 32;
      if (g.out_f32) reinterpret_cast<float2*>(cT)[ci] = make_float2((float)ar[r], (float)ai[r]);
      else cT[ci] = make_double2(ar[r], ai[r]);
    }
  }
}

static void launch_coeff_generic(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total,
                                 int64_t pix0, int npix_tile, int xpad, double2* cT, cudaStream_t st) {
  K1Geom g;
  g.ngroup = 1; g.use_tma = 0;
  g.npix_total = npix_total; g.pix0 = pix0; g.npix_tile = npix_tile; g.xpad = xpad;
  g.nv = pl->d.nv; g.nz = pl->d.nz; g.nfi = pl->nfi; g.ne = pl->fold.ne; g.no = pl->fold.no;
  g.pxb = 0; g.nt = pl->d.nt; g.is_log = is_log; g.v0 = pl->d.v0_idx;
  g.out_f32 = pl->d.precision == 1;
  const size_t smem = sizeof(double) * (size_t)(g.ne + g.no) * KG_COLS;
  PKM_REQUIRE(smem <= 200 * 1024, "nv=%d too large for the coefficient kernels", pl->d.nv);
  PKM_REQUIRE(pl->d.nt <= 65535, "nt=%d too large for the generic coefficient kernel", pl->d.nt);
  PKM_CUDA(cudaFuncSetAttribute(k_coeff_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ncol = (long long)npix_tile * g.nz;
  dim3 grid((unsigned)((ncol + KG_COLS - 1) / KG_COLS), (unsigned)pl->d.nt);
  KernelTimer tm(pl, PKM_K_COEFF, st);
  k_coeff_generic<<<grid, 256, smem, st>>>(dens, g, pl->d_CR, pl->d_CI, cT);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}

void launch_coeff(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total, int64_t pix0,
                  int npix_tile, int xpad, double2* cT, cudaStream_t st) {
  PKM_REQUIRE(xpad % 32 == 0 && xpad >= npix_tile, "internal: xpad=%d not a multiple of 32", xpad);
  if (pl->TM > 8 || getenv("PKM_COEFF_GENERIC")) {
    launch_coeff_generic(pl, dens, is_log, npix_total, pix0, npix_tile, xpad, cT, st);
    return;
  }
  int ngroup = 1;
  const int pxb = coeff_choose_pxb(pl, &ngroup);
  bool done = false;
  switch (pl->TM) {
#define CASE(T) case T: done = launch_coeff_oct_t<T>(pl, dens, is_log, npix_total, pix0, npix_tile, xpad, cT, st); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: break;
  }
  if (done) return;
  switch (pl->TM) {
#define CASE(T) case T: launch_coeff_t<T>(pl, dens, is_log, npix_total, pix0, npix_tile, xpad, pxb, ngroup, cT, st); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: PKM_THROW(PKM_ERR_INVALID, "unsupported TM=%d", pl->TM);
  }
}

// ============================================================================
// kernel 2
// ============================================================================
// Y-hat[k,x] = sum_{tz} S'[k,tz] * sum_{q<4} w[k][q] c[i_k+q, tz, x]     (10.2 FP64 ops per (k,tz,x))
//
// The kernel is bound by the FP64 pipe, and on B200 a DFMA whose three operands all come from
// the vector register file issues every 3 cycles instead of every 2.17 (tools/dfma_rf.cu: 24.5
// vs 34.2 TFLOP/s).  Everything is arranged around that:
//   * work item = one warp = (warp segment, block of 32 pixels): the lanes are 32 consecutive
//     pixels, every lane owns the SAME KT <= 25 consecutive frequencies of one spline span,
//     so the 4 coefficient rows it needs are shared by all its frequencies (4 coalesced
//     512-byte loads per (t,z), prefetched one (t,z) ahead in registers);
//   * the 4 warps of a CTA work on the SAME warp segment (4 pixel blocks), so the segment index
//     depends on blockIdx only and the 3 KT spline weights are read from the KERNEL-PARAMETER
//     bank with a uniform index: ptxas keeps them in uniform registers (LDCU.64 + DFMA with a UR
//     operand).  The 6 spline DFMAs of a term then read two vector registers and run at the full
//     rate; only the 4 complex-MAC DFMAs stay register-file limited.  The parameter bank holds
//     the weights of K2_WP_SEGS segments, so one pixel tile takes ceil(nwseg / K2_WP_SEGS) launches;
//   * with the weights out of the vector registers KT = 25 fits (2 segments per spline span at
//     the headline shape): the coefficient loads, the 6 DADDs forming d_q = c_q - c_3 and the
//     loop overhead are amortised over 2.5x more DFMAs than with KT = 10;
//   * S' is warp-uniform: it is staged in shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), one chunk = 12 (t,z) rows x KT frequencies per
//     copy, 3 stages per warp, and read back as broadcast LDS.128;
//   * warps are fully independent pipelines (own stages, own mbarriers, no CTA barrier).
// tools/contract_probe.cu is the harness the variants were measured with (one 7168-pixel tile:
// 10.19 ms for the round-1 kernel, 8.18 ms = 28.2 TFLOP/s for this one).
constexpr int K2_WARPS = 4;
constexpr int K2_STAGES = 3;
#ifndef PKM_K2_CTAS
#define PKM_K2_CTAS 3
#endif
#ifndef PKM_K2_G
#define PKM_K2_G 5
#endif
constexpr int K2_CTAS_PER_SM = PKM_K2_CTAS;
constexpr int K2_PF = 12;      // L2 prefetch distance of the coefficient rows in (t,z) steps
constexpr int K2_G = PKM_K2_G;  // frequencies whose FMA chains are interleaved
constexpr int K2_CHUNK_ELEMS = PKM_TZC * PKM_KT;                      // double2 per staged chunk
constexpr unsigned K2_CHUNK_BYTES = K2_CHUNK_ELEMS * sizeof(double2);  // 4800
constexpr int K2_WP_SEGS = 30000 / (PKM_KT * 3 * 8);                  // segments per launch (parameter bank: 32 KB)
struct K2Weights {
  double w[K2_WP_SEGS][PKM_KT][3];
};

__device__ __forceinline__ double2 ldg_nc_f64x2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(K2_WARPS * 32, K2_CTAS_PER_SM)
k_contract(const __grid_constant__ K2Weights wp, int ws0, int nws, const double2* __restrict__ Sw,
           const int4* __restrict__ wsegs, const double2* __restrict__ cT, double2* __restrict__ yhat,
           int ntz, int nchunk, int nfi, int np, int nfo, int npg) {
  extern __shared__ __align__(128) unsigned char k2_smem[];
  double2 (*sS)[K2_STAGES][K2_CHUNK_ELEMS] = reinterpret_cast<double2 (*)[K2_STAGES][K2_CHUNK_ELEMS]>(k2_smem);
  unsigned long long (*sBar)[K2_STAGES] = reinterpret_cast<unsigned long long (*)[K2_STAGES]>(
      k2_smem + sizeof(double2) * K2_WARPS * K2_STAGES * K2_CHUNK_ELEMS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wsl = blockIdx.x % nws;                   // uniform: warp segment within this launch
  const int pg = (blockIdx.x / nws) * K2_WARPS + warp;
  if (pg >= npg) return;  // warps never synchronise with each other
  const int ws = ws0 + wsl;
  const int4 sg = __ldg(wsegs + ws);  // {first coefficient row, k0, count, -}
  const int x = pg * 32 + lane;
  const int xl = min(x, np - 1);      // lanes past the tile re-read the last pixel, never store

  // The 4 B-spline taps of a frequency sum to one, so P-hat = c3 + sum_{q<3} w_q (c_q - c3):
  // three weights per frequency and 3 instead of 4 FP64 ops per component.
  const double (&w)[PKM_KT][3] = wp.w[wsl];
  double2 acc[PKM_KT];
#pragma unroll
  for (int r = 0; r < PKM_KT; ++r) acc[r] = make_double2(0.0, 0.0);

  const double2* Sg = Sw + (size_t)ws * nchunk * K2_CHUNK_ELEMS;
  const unsigned bar0 = smem_u32(&sBar[warp][0]);   // + 8 * stage
  const unsigned buf0 = smem_u32(&sS[warp][0][0]);  // + K2_CHUNK_BYTES * stage
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < K2_STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < K2_STAGES; ++s)
      if (s < nchunk) {
        mbar_expect_tx(bar0 + 8u * s, K2_CHUNK_BYTES);
        bulk_g2s(buf0 + K2_CHUNK_BYTES * s, Sg + (size_t)s * K2_CHUNK_ELEMS, K2_CHUNK_BYTES, bar0 + 8u * s);
      }
  }
  __syncwarp();

  // coefficients are stored pixel-group major, cT[x/32][tz][m][x%32]: the ntz*nfi rows one warp
  // walks are a single contiguous stream (TLB and DRAM-page friendly), 512 bytes per row
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
    const int nrow = min(PKM_TZC, ntz - ci * PKM_TZC);
    for (int j = 0; j < nrow; ++j, ++tz) {
      // d_q = c_q - c_3 frees the c registers: the loads of the next (t,z) step are issued
      // right here, a whole step ahead of their use (pinned with volatile asm)
      double2 d[3];
      const double2 c3 = c[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) d[q] = make_double2(c[q].x - c3.x, c[q].y - c3.y);
      cp += (tz + 1 < ntz) ? c_step : 0;  // branch-free: the last step re-reads its own rows
      // the coefficient tile streams from HBM: pull the rows a few steps ahead into L2
      if (tz + K2_PF < ntz) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)(K2_PF - 1) * c_step + q * 32));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) c[q] = ldg_nc_f64x2(cp + q * 32);
      const double2* srow = sbuf + j * PKM_KT;
#pragma unroll
      for (int h = 0; h < PKM_KT; h += K2_G) {
        double pr[K2_G], pi[K2_G];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
          for (int g = 0; g < K2_G; ++g)
            if (h + g < PKM_KT) pr[g] = fma(w[h + g][q], d[q].x, q == 0 ? c3.x : pr[g]);
#pragma unroll
          for (int g = 0; g < K2_G; ++g)
            if (h + g < PKM_KT) pi[g] = fma(w[h + g][q], d[q].y, q == 0 ? c3.y : pi[g]);
        }
#pragma unroll
        for (int g = 0; g < K2_G; ++g) {
          if (h + g >= PKM_KT) continue;
          const double2 s = srow[h + g];
          acc[h + g].x = fma(s.x, pr[g], acc[h + g].x);
          acc[h + g].y = fma(s.x, pi[g], acc[h + g].y);
          acc[h + g].x = fma(-s.y, pi[g], acc[h + g].x);
          acc[h + g].y = fma(s.y, pr[g], acc[h + g].y);
        }
      }
    }
    __syncwarp();  // every lane is done reading this stage
    if (lane == 0 && ci + K2_STAGES < nchunk) {
      mbar_expect_tx(bar0 + 8u * stage, K2_CHUNK_BYTES);
      bulk_g2s(buf0 + K2_CHUNK_BYTES * stage, Sg + (size_t)(ci + K2_STAGES) * K2_CHUNK_ELEMS, K2_CHUNK_BYTES,
               bar0 + 8u * stage);
    }
    if (++stage == K2_STAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (x < np) {
    double2* yo = yhat + (size_t)x * nfo + sg.y;
#pragma unroll
    for (int r = 0; r < PKM_KT; ++r)
      if (r < sg.z) yo[r] = acc[r];
  }
}

// ----------------------------------------------------------------------------
// float32-arithmetic variant of the contraction (plan precision = 1): coefficients and S' are
// float2, products and the spline run on the FP32 pipe; every staged chunk of 12 (t,z) steps is
// accumulated in float32 and then folded into float64 accumulators, so the 636-term sum does
// not lose accuracy to float32 accumulation.  Same work decomposition and TMA staging as above.
// ----------------------------------------------------------------------------
constexpr int K2F_CHUNK_ELEMS = PKM_TZC * PKM_KT32;                       // float2 per staged chunk
constexpr unsigned K2F_CHUNK_BYTES = K2F_CHUNK_ELEMS * sizeof(float2);   // 960

__global__ void __launch_bounds__(K2_WARPS * 32, 3)
k_contract_f32(const float2* __restrict__ Sw, const double* __restrict__ Ww,
               const int4* __restrict__ wsegs, const float2* __restrict__ cT,
               double2* __restrict__ yhat, int nwseg, int ntz, int nchunk, int nfi, int np, int nfo,
               long long nitems) {
  __shared__ __align__(128) float2 sS[K2_WARPS][K2_STAGES][K2F_CHUNK_ELEMS];
  __shared__ __align__(8) unsigned long long sBar[K2_WARPS][K2_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * K2_WARPS + warp;
  if (item >= nitems) return;
  const int pg = (int)(item / nwseg), ws = (int)(item - (long long)pg * nwseg);
  const int4 sg = __ldg(wsegs + ws);
  const int x = pg * 32 + lane;
  const int xl = min(x, np - 1);

  float w[PKM_KT32][3];
#pragma unroll
  for (int r = 0; r < PKM_KT32; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) w[r][q] = (float)__ldg(Ww + ((size_t)ws * PKM_KT32 + r) * 4 + q);
  double2 acc[PKM_KT32];
#pragma unroll
  for (int r = 0; r < PKM_KT32; ++r) acc[r] = make_double2(0.0, 0.0);

  const float2* Sg = Sw + (size_t)ws * nchunk * K2F_CHUNK_ELEMS;
  const unsigned bar0 = smem_u32(&sBar[warp][0]);
  const unsigned buf0 = smem_u32(&sS[warp][0][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < K2_STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < K2_STAGES; ++s)
      if (s < nchunk) {
        mbar_expect_tx(bar0 + 8u * s, K2F_CHUNK_BYTES);
        bulk_g2s(buf0 + K2F_CHUNK_BYTES * s, Sg + (size_t)s * K2F_CHUNK_ELEMS, K2F_CHUNK_BYTES, bar0 + 8u * s);
      }
  }
  __syncwarp();

  const size_t c_step = (size_t)nfi * 32;
  const float2* cp = cT + (((size_t)pg * ntz) * nfi + sg.x) * 32 + (xl & 31);
  float2 c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q] = __ldg(cp + q * 32);

  int stage = 0;
  unsigned parity = 0;
  int tz = 0;
  for (int ci = 0; ci < nchunk; ++ci) {
    mbar_wait(bar0 + 8u * stage, parity);
    const float2* sbuf = &sS[warp][stage][0];
    const int nrow = min(PKM_TZC, ntz - ci * PKM_TZC);
    float2 a32[PKM_KT32];
#pragma unroll
    for (int r = 0; r < PKM_KT32; ++r) a32[r] = make_float2(0.f, 0.f);
    for (int j = 0; j < nrow; ++j, ++tz) {
      const float2 c3 = c[3];
      float2 d[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) d[q] = make_float2(c[q].x - c3.x, c[q].y - c3.y);
      cp += (tz + 1 < ntz) ? c_step : 0;
      if (tz + K2_PF < ntz)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)(K2_PF - 1) * c_step + (ws & 3) * 32));
#pragma unroll
      for (int q = 0; q < 4; ++q) c[q] = __ldg(cp + q * 32);
#pragma unroll
      for (int r = 0; r < PKM_KT32; ++r) {
        const float2 s = sbuf[j * PKM_KT32 + r];
        float pr = c3.x, pi = c3.y;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          pr = fmaf(w[r][q], d[q].x, pr);
          pi = fmaf(w[r][q], d[q].y, pi);
        }
        a32[r].x = fmaf(s.x, pr, a32[r].x);
        a32[r].x = fmaf(-s.y, pi, a32[r].x);
        a32[r].y = fmaf(s.x, pi, a32[r].y);
        a32[r].y = fmaf(s.y, pr, a32[r].y);
      }
    }
#pragma unroll
    for (int r = 0; r < PKM_KT32; ++r) {
      acc[r].x += (double)a32[r].x;
      acc[r].y += (double)a32[r].y;
    }
    __syncwarp();
    if (lane == 0 && ci + K2_STAGES < nchunk) {
      mbar_expect_tx(bar0 + 8u * stage, K2F_CHUNK_BYTES);
      bulk_g2s(buf0 + K2F_CHUNK_BYTES * stage, Sg + (size_t)(ci + K2_STAGES) * K2F_CHUNK_ELEMS, K2F_CHUNK_BYTES,
               bar0 + 8u * stage);
    }
    if (++stage == K2_STAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (x < np) {
    double2* yo = yhat + (size_t)x * nfo + sg.y;
#pragma unroll
    for (int r = 0; r < PKM_KT32; ++r)
      if (r < sg.z) yo[r] = acc[r];
  }
}

double contract_wave_pixel_groups(const pkm_plan* pl) {
  // resident warps per GPU / warps per pixel group (one warp = one warp segment x 32 pixels)
  const double warps = (double)pl->num_sms * K2_WARPS * (pl->d.precision == 1 ? 3 : K2_CTAS_PER_SM);
  return std::max(1e-3, warps / std::max(1, pl->wsegs.nwseg));
}

// side stream (+ fork / join events) for the alternate launches of one tile, created on first use
static bool contract_side_stream(pkm_plan* pl, int lane) {
  if (pl->s_k2[lane]) return true;
  if (getenv("PKM_CONTRACT_NO_FORK")) return false;
  PKM_CUDA(cudaStreamCreateWithFlags(&pl->s_k2[lane], cudaStreamNonBlocking));
  PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_k2_fork[lane], cudaEventDisableTiming));
  PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_k2_join[lane], cudaEventDisableTiming));
  return true;
}

void launch_contract(pkm_plan* pl, const double2* cT, int npix_tile, int xpad, double2* yhat,
                     cudaStream_t st, int lane) {
  if (pl->d.precision == 1) {
    const long long npg = (npix_tile + 31) / 32;
    const long long nitems = npg * pl->wsegs.nwseg;
    const unsigned grid = (unsigned)((nitems + K2_WARPS - 1) / K2_WARPS);
    KernelTimer tm(pl, PKM_K_CONTRACT, st);
    k_contract_f32<<<grid, K2_WARPS * 32, 0, st>>>(pl->d_Sw32, pl->d_wseg_w, reinterpret_cast<const int4*>(pl->d_wseg),
                                                  reinterpret_cast<const float2*>(cT), yhat, pl->wsegs.nwseg,
                                                  pl->ntz, pl->nchunk, pl->nfi, npix_tile, pl->nfo, nitems);
    PKM_CUDA(cudaGetLastError());
    pl->launches++;
    return;
  }
  (void)xpad;
  const int npg = (npix_tile + 31) / 32;
  const size_t smem = sizeof(double2) * K2_WARPS * K2_STAGES * K2_CHUNK_ELEMS +
                      sizeof(unsigned long long) * K2_WARPS * K2_STAGES;
  PKM_CUDA(cudaFuncSetAttribute(k_contract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelTimer tm(pl, PKM_K_CONTRACT, st);
  static_assert(sizeof(K2Weights) <= 30000, "the spline weights of one launch must fit the parameter bank");
  // The launches of one tile are independent: every other one goes to a side stream so that the
  // CTAs of the next launch fill the SMs while the last wave of the previous one drains.
  const int nlaunch = (pl->wsegs.nwseg + K2_WP_SEGS - 1) / K2_WP_SEGS;
  const bool fork = nlaunch > 1 && contract_side_stream(pl, lane);
  if (fork) {
    PKM_CUDA(cudaEventRecord(pl->ev_k2_fork[lane], st));
    PKM_CUDA(cudaStreamWaitEvent(pl->s_k2[lane], pl->ev_k2_fork[lane], 0));
  }
  K2Weights wp;
  int li = 0;
  for (int ws0 = 0; ws0 < pl->wsegs.nwseg; ws0 += K2_WP_SEGS, ++li) {
    const int nws = std::min(K2_WP_SEGS, pl->wsegs.nwseg - ws0);
    for (int i = 0; i < nws; ++i)
      for (int r = 0; r < PKM_KT; ++r)
        for (int q = 0; q < 3; ++q) wp.w[i][r][q] = pl->wsegs.w[((size_t)(ws0 + i) * PKM_KT + r) * 4 + q];
    const unsigned grid = (unsigned)(((npg + K2_WARPS - 1) / K2_WARPS) * nws);
    cudaStream_t ls = (fork && (li & 1)) ? pl->s_k2[lane] : st;
    k_contract<<<grid, K2_WARPS * 32, smem, ls>>>(wp, ws0, nws, pl->d_Sw, reinterpret_cast<const int4*>(pl->d_wseg), cT,
                                              yhat, pl->ntz, pl->nchunk, pl->nfi, npix_tile, pl->nfo, npg);
    PKM_CUDA(cudaGetLastError());
    pl->launches++;
  }
  if (fork) {
    PKM_CUDA(cudaEventRecord(pl->ev_k2_join[lane], pl->s_k2[lane]));
    PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_k2_join[lane], 0));
  }
}
