// Density path (Component.evaluate_ybar, /root/reference/popkinmocks/components/base.py:836-859)
// as two FP64 kernels.  With  F = dv * rfft(roll(p))  and the not-a-knot spline written as
// P-hat[k] = sum_q w[k][q] c[i_k + q],  c = A^{-1} F,  the datacube spectrum of one pixel is
//
//     Y-hat[k] = sum_{t,z} S'[k,t,z] * sum_{q<4} w[k][q] * c[i_k+q, t, z]
//
// kernel 1 (coeff):    c = (A^{-1} D) p   - a small dense real->complex transform per
//                      (t,x,z) column, done as a register-tiled FP64 GEMM with the even/odd
//                      fold of the real DFT (halves the flops), exp(log p) fused into the load.
// kernel 2 (contract): streams S' (L2 resident) and the coefficient rows, evaluates the 4-tap
//                      spline on the fly and accumulates over (t,z) in registers.
// P-hat (25 MB per pixel in the reference, base.py:846-853) is never materialised.
#include <cstdio>

#include "pkm_internal.h"

// ============================================================================
// kernel 1
// ============================================================================
// Persistent CTAs (one per SM), each looping over tiles = (age t, block of PXB pixels), i.e.
// NC = PXB*nz density columns of nv velocities.  The folded DFT*spline matrices stay resident
// in shared memory for the CTA's lifetime.  Per tile:
//   phase 1: load the [nv][NC] slab (coalesced: for fixed v the columns are contiguous), exp fused
//   phase 2: fold v <-> -v in place (e = sum, o = difference of the rolled profile)
//   phase 3: warp-tiled FP64 GEMM; lanes = 8 row groups x 4 column groups, thread tile TM x 4;
//            one work item = (16-column chunk, real | imaginary part) per warp
//   phase 4: results staged as [z][m][x] in shared memory, written as 16*PXB-byte runs so the
//            contraction kernel reads pixel-contiguous coefficient rows.
constexpr int K1_THREADS = 384;
constexpr int K1_WARPS = K1_THREADS / 32;

template <int TM>
__global__ void __launch_bounds__(K1_THREADS, 1)
k_coeff(const double* __restrict__ dens, int is_log, long long npix_total, long long pix0,
        int npix_tile, int xpad, int nv, int nz, int nfi, int ne, int no, int pxb, int nt,
        const double* __restrict__ CRt, const double* __restrict__ CIt,
        const int* __restrict__ idx_plus, const int* __restrict__ idx_minus,
        double2* __restrict__ cT) {
  extern __shared__ double smem[];
  constexpr int TMP = TM + (TM & 1);  // even row padding -> 16-byte aligned double2 loads
  constexpr int NFIP = 8 * TMP;
  const int NC = pxb * nz;
  const int NCp = (NC + 15) & ~15;
  double* As_r = smem;                                   // [ne][NFIP]
  double* As_i = As_r + (size_t)ne * NFIP;               // [max(no,1)][NFIP]
  double* Bs = As_i + (size_t)max(no, 1) * NFIP;         // [nv][NCp]
  double* Cs = Bs + (size_t)nv * NCp;                    // [nz][nfi][pxb][2]
  int* rows_e = reinterpret_cast<int*>(Cs + (size_t)nz * nfi * pxb * 2);  // [ne]
  int* rows_o = rows_e + ne;                             // [ne] (entry u-1 used for u>=1)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int rg = lane & 7, cg = lane >> 3;

  for (int i = tid; i < ne * NFIP; i += K1_THREADS) As_r[i] = CRt[i];
  for (int i = tid; i < no * NFIP; i += K1_THREADS) As_i[i] = CIt[i];
  for (int i = tid; i < ne; i += K1_THREADS) {
    rows_e[i] = idx_plus[i];
    rows_o[i] = idx_minus[i];
  }
  __syncthreads();

  const int nxb = (npix_tile + pxb - 1) / pxb;
  const long long ntile = (long long)nxb * nt;
  const int nchunk = NCp >> 4;
  const int nitem = nchunk * 2;

  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int t = (int)(tile / nxb);
    const int xb = (int)(tile - (long long)t * nxb);
    // phase 1
    const long long col_base = (pix0 + (long long)xb * pxb) * nz;
    const int ncol_valid = min(pxb, npix_tile - xb * pxb) * nz;
    for (int i = tid; i < nv * NCp; i += K1_THREADS) {
      int v = i / NCp, col = i - v * NCp;
      double val = 0.0;
      if (col < ncol_valid) {
        val = __ldg(dens + ((long long)t * nv + v) * npix_total * nz + col_base + col);
        if (is_log) val = exp(val);
      }
      Bs[(size_t)v * NCp + col] = val;
    }
    __syncthreads();
    // phase 2 (self-paired rows carry weight 1/2 in CRt, so e = a + b unconditionally)
    for (int i = tid; i < ne * NCp; i += K1_THREADS) {
      int u = i / NCp, col = i - u * NCp;
      int rp = rows_e[u], rm = rows_o[u];
      double a = Bs[(size_t)rp * NCp + col], b = Bs[(size_t)rm * NCp + col];
      Bs[(size_t)rp * NCp + col] = a + b;
      if (rp != rm) Bs[(size_t)rm * NCp + col] = a - b;
    }
    __syncthreads();
    // phase 3 + staging
    for (int item = warp; item < nitem; item += K1_WARPS) {
      const int part = item & 1, chunk = item >> 1;
      const double* A = part ? As_i : As_r;
      const int* rows = part ? (rows_o + 1) : rows_e;
      const int K = part ? no : ne;
      const int col0 = chunk * 16 + cg * 4;
      double acc[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 2
      for (int u = 0; u < K; ++u) {
        const double2* bp = reinterpret_cast<const double2*>(Bs + (size_t)rows[u] * NCp + col0);
        const double2 b01 = bp[0], b23 = bp[1];
        const double2* ap = reinterpret_cast<const double2*>(A + (size_t)u * NFIP + rg * TMP);
        double a[TMP];
#pragma unroll
        for (int i = 0; i < TMP / 2; ++i) {
          double2 v2 = ap[i];
          a[2 * i] = v2.x;
          a[2 * i + 1] = v2.y;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][0] = fma(a[i], b01.x, acc[i][0]);
          acc[i][1] = fma(a[i], b01.y, acc[i][1]);
          acc[i][2] = fma(a[i], b23.x, acc[i][2]);
          acc[i][3] = fma(a[i], b23.y, acc[i][3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int col = col0 + j;
        if (col < NC) {
          int x = col / nz, z = col - x * nz;
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            int m = rg * TM + i;
            if (m < nfi) Cs[(((size_t)z * nfi + m) * pxb + x) * 2 + part] = acc[i][j];
          }
        }
      }
    }
    __syncthreads();
    // phase 4: cT[((t*nz + z)*nfi + m)*xpad + xb*pxb + x]
    for (int i = tid; i < nz * nfi * pxb; i += K1_THREADS) {
      int zm = i / pxb, x = i - zm * pxb;
      double2 v = reinterpret_cast<const double2*>(Cs)[i];
      cT[((size_t)t * nz * nfi + zm) * xpad + (size_t)xb * pxb + x] = v;
    }
    // the next tile's phase 1/2 only touch Bs; Cs is rewritten after two more barriers
  }
}

// The GEMM phase reads A rows as As[u][rg*TMP + i]; the host stores row m = rg*TM + i there.
static void upload_padded_transposed(const std::vector<double>& C, int nfi, int K, int TM,
                                     double** d_out) {
  const int TMP = TM + (TM & 1), NFIP = 8 * TMP;
  std::vector<double> h((size_t)std::max(K, 1) * NFIP, 0.0);
  for (int u = 0; u < K; ++u)
    for (int m = 0; m < nfi; ++m) {
      int rg = m / TM, i = m % TM;
      h[(size_t)u * NFIP + rg * TMP + i] = C[(size_t)m * K + u];
    }
  PKM_CUDA(cudaMalloc(d_out, h.size() * sizeof(double)));
  PKM_CUDA(cudaMemcpy(*d_out, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
}

void density_tables_upload(pkm_plan* pl) {
  // rows per thread: 8 row groups must cover nfi rows
  pl->TM = (pl->nfi + 7) / 8;
  PKM_REQUIRE(pl->TM <= 8, "nv=%d too large for the coefficient kernel (nfi=%d > 64)", pl->d.nv, pl->nfi);
  const int TMP = pl->TM + (pl->TM & 1);
  pl->nfiP = 8 * TMP;
  upload_padded_transposed(pl->fold.CR, pl->nfi, pl->fold.ne, pl->TM, &pl->d_CRt);
  upload_padded_transposed(pl->fold.CI, pl->nfi, pl->fold.no, pl->TM, &pl->d_CIt);
  PKM_CUDA(cudaMalloc(&pl->d_idx_plus, pl->fold.ne * sizeof(int32_t)));
  PKM_CUDA(cudaMalloc(&pl->d_idx_minus, pl->fold.ne * sizeof(int32_t)));
  PKM_CUDA(cudaMemcpy(pl->d_idx_plus, pl->fold.idx_plus.data(), pl->fold.ne * sizeof(int32_t), cudaMemcpyHostToDevice));
  PKM_CUDA(cudaMemcpy(pl->d_idx_minus, pl->fold.idx_minus.data(), pl->fold.ne * sizeof(int32_t), cudaMemcpyHostToDevice));
}

static size_t coeff_smem_bytes(const pkm_plan* pl, int pxb) {
  const int NC = pxb * pl->d.nz, NCp = (NC + 15) & ~15;
  size_t dbl = (size_t)(pl->fold.ne + std::max(pl->fold.no, 1)) * pl->nfiP + (size_t)pl->d.nv * NCp +
               (size_t)pl->d.nz * pl->nfi * pxb * 2;
  return dbl * sizeof(double) + 2 * (size_t)pl->fold.ne * sizeof(int);
}

// pixels per tile: largest of {16, 8, 4, 2, 1} whose slab + staging fits shared memory
int coeff_choose_pxb(const pkm_plan* pl) {
  for (int pxb : {16, 8, 4, 2, 1})
    if (coeff_smem_bytes(pl, pxb) <= 227 * 1024) return pxb;
  PKM_THROW(PKM_ERR_INVALID, "coefficient kernel: nv=%d nz=%d does not fit shared memory", pl->d.nv, pl->d.nz);
}

template <int TM>
static void launch_coeff_t(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total,
                           int64_t pix0, int npix_tile, int xpad, int pxb, double2* cT,
                           cudaStream_t st) {
  size_t smem = coeff_smem_bytes(pl, pxb);
  PKM_CUDA(cudaFuncSetAttribute(k_coeff<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t ntile = (int64_t)((npix_tile + pxb - 1) / pxb) * pl->d.nt;
  int grid = (int)std::min<int64_t>(ntile, pl->num_sms);
  KernelTimer tm(pl, PKM_K_COEFF, st);
  k_coeff<TM><<<grid, K1_THREADS, smem, st>>>(dens, is_log, npix_total, pix0, npix_tile, xpad,
                                              pl->d.nv, pl->d.nz, pl->nfi, pl->fold.ne, pl->fold.no,
                                              pxb, pl->d.nt, pl->d_CRt, pl->d_CIt, pl->d_idx_plus,
                                              pl->d_idx_minus, cT);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}

void launch_coeff(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total, int64_t pix0,
                  int npix_tile, int xpad, double2* cT, cudaStream_t st) {
  const int pxb = coeff_choose_pxb(pl);
  PKM_REQUIRE(xpad % pxb == 0 && xpad >= npix_tile, "internal: xpad=%d not a multiple of %d", xpad, pxb);
  switch (pl->TM) {
#define CASE(T) case T: launch_coeff_t<T>(pl, dens, is_log, npix_total, pix0, npix_tile, xpad, pxb, cT, st); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: PKM_THROW(PKM_ERR_INVALID, "unsupported TM=%d", pl->TM);
  }
}

// ============================================================================
// kernel 2
// ============================================================================
// warp = 4 groups x 8 pixels; thread = KT consecutive frequencies (one spline span) of one pixel.
// Per (t,z): 4 complex S' values (64 B contiguous per instruction across the 4 groups of the
// warp), the 4 coefficient taps of the pixel (128 B contiguous across the 8 pixel lanes), then
// 12 flops per frequency: 8 for the real-weight spline evaluation, 4 for the complex MAC.
__global__ void __launch_bounds__(256, 2)
k_contract(const double2* __restrict__ Sp, const double2* __restrict__ cT,
           const int* __restrict__ group_span, const double* __restrict__ slot_w,
           const int* __restrict__ slot_k, double2* __restrict__ yhat, int ntz, int nfi, int xpad,
           int nslot, int nunit, int nfo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 8 + warp;
  if (unit >= nunit) return;
  const int px = lane & 7, gsub = lane >> 3;
  const int x = blockIdx.y * 8 + px;
  const int g = unit * PKM_GPW + gsub;
  const int span = group_span[g];
  double w[PKM_KT][4];
#pragma unroll
  for (int r = 0; r < PKM_KT; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) w[r][q] = slot_w[(size_t)((unit * PKM_KT + r) * PKM_GPW + gsub) * 4 + q];
  double2 acc[PKM_KT];
#pragma unroll
  for (int r = 0; r < PKM_KT; ++r) acc[r] = make_double2(0.0, 0.0);

  const double2* sp = Sp + (size_t)unit * PKM_KT * PKM_GPW + gsub;
  const double2* cp = cT + (size_t)span * xpad + x;
  const size_t c_step = (size_t)nfi * xpad;
#pragma unroll 2
  for (int tz = 0; tz < ntz; ++tz) {
    double2 s[PKM_KT], c[4];
#pragma unroll
    for (int r = 0; r < PKM_KT; ++r) s[r] = __ldg(sp + r * PKM_GPW);
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = __ldg(cp + (size_t)q * xpad);
    sp += nslot;
    cp += c_step;
#pragma unroll
    for (int r = 0; r < PKM_KT; ++r) {
      double pr = w[r][0] * c[0].x;
      double pi = w[r][0] * c[0].y;
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        pr = fma(w[r][q], c[q].x, pr);
        pi = fma(w[r][q], c[q].y, pi);
      }
      acc[r].x = fma(s[r].x, pr, acc[r].x);
      acc[r].x = fma(-s[r].y, pi, acc[r].x);
      acc[r].y = fma(s[r].x, pi, acc[r].y);
      acc[r].y = fma(s[r].y, pr, acc[r].y);
    }
  }
#pragma unroll
  for (int r = 0; r < PKM_KT; ++r) {
    int k = slot_k[(unit * PKM_KT + r) * PKM_GPW + gsub];
    if (k >= 0) yhat[(size_t)x * nfo + k] = acc[r];
  }
}

void launch_contract(pkm_plan* pl, const double2* cT, int npix_tile, int xpad, double2* yhat,
                     cudaStream_t st) {
  dim3 grid((pl->slots.nunit + 7) / 8, (npix_tile + 7) / 8);
  KernelTimer tm(pl, PKM_K_CONTRACT, st);
  k_contract<<<grid, 256, 0, st>>>(pl->d_Sp, cT, pl->d_group_span, pl->d_slot_w, pl->d_slot_k, yhat,
                                   pl->ntz, pl->nfi, xpad, pl->slots.nslot, pl->slots.nunit, pl->nfo);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}
