// Device-side generation of the O(pixels) inputs of the Gaussian-LOSVD path for the parametric
// components: the maps p(x|t), mu_v(t,x), sig_v(t,x), t_dep(x) of a
// GrowingDisk (/root/reference/popkinmocks/components/growing_disk.py:30-239) and of a Stream
// (stream.py:28-138), the per-row normalisation, and the assembly of P(t,x,z) from the stored
// factors (parametric.py:541-566 with the enrichment model of parametric.py:253-289 tabulated per
// depletion-time row on the host).  The host keeps only O(nt), O(rows) work (beta cdf of the star
// formation history, the enrichment table); nothing of size nx1*nx2 crosses PCIe.
#include <algorithm>
#include <cmath>

#include "pkm_internal.h"

namespace {

struct Frame {
  double cx, cy, c, s;   // centre and cos / sin of the rotation (parametric.py:43-49)
  int nx1, nx2;
};

// coordinates of pixel (i, j) in the component's own frame
__device__ __forceinline__ void frame_xy(const Frame& f, const double* __restrict__ x1,
                                         const double* __restrict__ x2, long long pix, double& xp, double& yp) {
  const int i = (int)(pix / f.nx2), j = (int)(pix - (long long)i * f.nx2);
  const double d1 = x1[i] - f.cx, d2 = x2[j] - f.cy;
  xp = __dsub_rn(__dmul_rn(f.c, d1), __dmul_rn(f.s, d2));     // no fma contraction: NumPy's arithmetic
  yp = __dadd_rn(__dmul_rn(f.s, d1), __dmul_rn(f.c, d2));
}

__device__ __forceinline__ double ell_radius(double xp, double yp, double q) {
  const double yq = yp / q;
  return sqrt(__dadd_rn(__dmul_rn(xp, xp), __dmul_rn(yq, yq)));
}

// age_pars rows (each [nt]): 0 q, 1 rc, 2 alpha of p(x|t); 3 q, 4 rc, 5 K of mu_v;
// 6 q, 7 alpha, 8 sig_in, 9 sig_out of sig_v
__global__ void k_disk_maps(Frame f, const double* __restrict__ x1, const double* __restrict__ x2,
                            const double* __restrict__ ap, int nt, double scale, double* __restrict__ log_rho,
                            double* __restrict__ mu_v, double* __restrict__ sig_v) {
  const long long npix = (long long)f.nx1 * f.nx2;
  const int t = blockIdx.y;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    double xp, yp;
    frame_xy(f, x1, x2, pix, xp, yp);
    const size_t o = (size_t)t * npix + pix;
    // surface density: rho ~ (r + rc)^-alpha  (growing_disk.py:30-71)
    log_rho[o] = -ap[2 * nt + t] * log(ell_radius(xp, yp, ap[0 * nt + t]) + ap[1 * nt + t]);
    // rotation: K r / (r + rc)^3 cos(theta)  (growing_disk.py:122-170)
    {
      const double q = ap[3 * nt + t], rc = ap[4 * nt + t], K = ap[5 * nt + t];
      const double r = ell_radius(xp, yp, q), theta = atan2(yp / q, xp);
      const double d = r + rc;
      mu_v[o] = K * r / (d * d * d) * cos(theta);
    }
    // dispersion: power law from sig_in towards sig_out, saturating  (growing_disk.py:172-239)
    {
      const double q = ap[6 * nt + t], al = ap[7 * nt + t], s_in = ap[8 * nt + t], s_out = ap[9 * nt + t];
      const double sg = s_in + (s_out - s_in) * pow(ell_radius(xp, yp, q) / scale, al);
      const bool beyond = s_out > s_in ? sg > s_out : sg < s_out;
      sig_v[o] = beyond ? s_out : sg;
    }
  }
}

// depletion time map and the nearest row of the tabulated depletion-time grid (growing_disk.py:73-120,
// parametric.py:262: argmin |t_dep - grid|, first minimum on ties)
__global__ void k_disk_t_dep(Frame f, const double* __restrict__ x1, const double* __restrict__ x2, double q,
                             double alpha, double t_in, double t_out, double scale,
                             const double* __restrict__ grid, int ngrid, double* __restrict__ t_dep,
                             int* __restrict__ row) {
  const long long npix = (long long)f.nx1 * f.nx2;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    double xp, yp;
    frame_xy(f, x1, x2, pix, xp, yp);
    double td = t_in + (t_out - t_in) * pow(ell_radius(xp, yp, q) / scale, alpha);
    const bool beyond = t_in <= t_out ? td > t_out : td < t_out;
    if (beyond) td = t_out;
    t_dep[pix] = td;
    int lo = 0, hi = ngrid - 1;                       // grid ascending: last index with grid <= td
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (grid[mid] <= td) lo = mid; else hi = mid;
    }
    int best = lo;
    if (ngrid > 1) {
      if (td < grid[0]) best = 0;
      else best = (fabs(td - grid[hi]) < fabs(td - grid[lo])) ? hi : lo;
    }
    row[pix] = best;
  }
}

// log of the standard normal cdf, accurate in both tails
__device__ __forceinline__ double log_ndtr_m(double x) {
  const double rs2 = 0.70710678118654752440;
  if (x < -1.0) return log(0.5 * erfcx(-x * rs2)) - 0.5 * x * x;
  if (x < 0.0) return log(0.5 * erfc(-x * rs2));
  return log1p(-0.5 * erfc(x * rs2));
}
__device__ __forceinline__ double log_diff_exp(double hi, double lo) {   // log(e^hi - e^lo), hi >= lo
  if (hi == -INFINITY) return -INFINITY;
  return hi + log(1.0 - exp(lo - hi));
}

// Stream: unnormalised log pixel mass of nsmp equally weighted isotropic Gaussians along the arc,
// integrated exactly over each pixel (stream.py:28-90), and the mean-velocity map (stream.py:105-127)
__global__ void k_stream_maps(Frame f, const double* __restrict__ x1, const double* __restrict__ x2,
                              const double* __restrict__ cxs, const double* __restrict__ cys, int nsmp,
                              double sig, double hwx, double hwy, double th_lo, double th_hi, double v_lo,
                              double v_hi, double* __restrict__ log_p, double* __restrict__ mu_v) {
  extern __shared__ double s_c[];            // centres [2][nsmp]
  for (int i = threadIdx.x; i < nsmp; i += blockDim.x) {
    s_c[i] = cxs[i];
    s_c[nsmp + i] = cys[i];
  }
  __syncthreads();
  const long long npix = (long long)f.nx1 * f.nx2;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    double xp, yp;
    frame_xy(f, x1, x2, pix, xp, yp);
    double m = -INFINITY, sum = 0.0;          // online log-sum-exp over the samples
    for (int k = 0; k < nsmp; ++k) {
      const double lx = log_diff_exp(log_ndtr_m((xp + hwx - s_c[k]) / sig), log_ndtr_m((xp - hwx - s_c[k]) / sig));
      const double ly = log_diff_exp(log_ndtr_m((yp + hwy - s_c[nsmp + k]) / sig),
                                     log_ndtr_m((yp - hwy - s_c[nsmp + k]) / sig));
      const double a = lx + ly;
      if (a > m) {
        sum = sum * exp(m - a) + 1.0;
        m = a;
      } else if (a > -INFINITY) {
        sum += exp(a - m);
      }
    }
    log_p[pix] = m > -INFINITY ? m + log(sum) : -INFINITY;
    const double theta = atan2(yp, xp);
    double mu;
    if (theta <= th_lo) mu = v_lo;
    else if (theta >= th_hi) mu = v_hi;
    else mu = (theta - th_lo) / (th_hi - th_lo) * (v_hi - v_lo) + v_lo;
    mu_v[pix] = mu;
  }
}

// a[r][:] -= logsumexp_c(a[r][c] + log_cell): one CTA per row (rows are ages; 4e4 columns)
__global__ void k_normalise_rows(double* __restrict__ a, long long ncol, double log_cell) {
  __shared__ double red[256];
  double* row = a + (size_t)blockIdx.x * ncol;
  double m = -INFINITY;
  for (long long c = threadIdx.x; c < ncol; c += blockDim.x) m = fmax(m, row[c]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + st]);
    __syncthreads();
  }
  m = red[0];
  __syncthreads();
  double s = 0.0;
  if (m > -INFINITY)
    for (long long c = threadIdx.x; c < ncol; c += blockDim.x) s += exp(row[c] - m);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  const double norm = m + log(red[0]) + log_cell;
  for (long long c = threadIdx.x; c < ncol; c += blockDim.x) row[c] -= norm;
}

// P(t,x,z) = exp(log P(t) + log p(x|t) + log dxdy + log P(z | t, row(x)))  and, optionally, the log
// DENSITY log p(t,x,z) = log p(t) + log p(x|t) + log p(z|t,x)   (parametric.py:541-566)
__global__ void k_assemble_txz(const double* __restrict__ log_P_t, const double* __restrict__ log_dt,
                               const double* __restrict__ log_px, long long px_age_stride, double log_dxdy,
                               const double* __restrict__ log_Pz, const double* __restrict__ log_dz,
                               const int* __restrict__ row, int nt, long long npix, int nz,
                               double* __restrict__ P_txz, double* __restrict__ log_p_txz) {
  const long long n = (long long)nt * npix * nz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % nz);
    const long long tp = i / nz;
    const long long pix = tp % npix;
    const int t = (int)(tp / npix);
    const int r = row ? row[pix] : 0;
    const double lpx = log_px[(size_t)t * px_age_stride + pix];
    const double lPz = log_Pz[((size_t)r * nt + t) * nz + z];
    const double lP = log_P_t[t] + (lpx + log_dxdy) + lPz;
    if (P_txz) P_txz[i] = exp(lP);
    if (log_p_txz) log_p_txz[i] = (log_P_t[t] - log_dt[t]) + lpx + (lPz - log_dz[z]);
  }
}

int grid_for(long long n, int num_sms) {
  return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)num_sms * 16));
}

}  // namespace

void launch_disk_maps(const ModelFrame& mf, const double* d_x1, const double* d_x2, const double* d_age_pars,
                      int nt, double scale, double* log_rho, double* mu_v, double* sig_v, int num_sms,
                      cudaStream_t st) {
  Frame f{mf.cx, mf.cy, mf.c, mf.s, mf.nx1, mf.nx2};
  const long long npix = (long long)mf.nx1 * mf.nx2;
  dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((npix + 255) / 256, num_sms * 4)), (unsigned)nt);
  k_disk_maps<<<grid, 256, 0, st>>>(f, d_x1, d_x2, d_age_pars, nt, scale, log_rho, mu_v, sig_v);
  PKM_CUDA(cudaGetLastError());
}

void launch_disk_t_dep(const ModelFrame& mf, const double* d_x1, const double* d_x2, double q, double alpha,
                       double t_in, double t_out, double scale, const double* d_grid, int ngrid, double* t_dep,
                       int* row, int num_sms, cudaStream_t st) {
  Frame f{mf.cx, mf.cy, mf.c, mf.s, mf.nx1, mf.nx2};
  const long long npix = (long long)mf.nx1 * mf.nx2;
  k_disk_t_dep<<<grid_for(npix, num_sms), 256, 0, st>>>(f, d_x1, d_x2, q, alpha, t_in, t_out, scale, d_grid,
                                                         ngrid, t_dep, row);
  PKM_CUDA(cudaGetLastError());
}

void launch_stream_maps(const ModelFrame& mf, const double* d_x1, const double* d_x2, const double* d_cx,
                        const double* d_cy, int nsmp, double sig, double hwx, double hwy, double th_lo,
                        double th_hi, double v_lo, double v_hi, double* log_p, double* mu_v, int num_sms,
                        cudaStream_t st) {
  Frame f{mf.cx, mf.cy, mf.c, mf.s, mf.nx1, mf.nx2};
  const long long npix = (long long)mf.nx1 * mf.nx2;
  const size_t smem = sizeof(double) * 2 * (size_t)nsmp;
  PKM_REQUIRE(smem <= 40 * 1024, "too many stream samples (nsmp=%d)", nsmp);
  k_stream_maps<<<grid_for(npix, num_sms), 256, smem, st>>>(f, d_x1, d_x2, d_cx, d_cy, nsmp, sig, hwx, hwy, th_lo,
                                                            th_hi, v_lo, v_hi, log_p, mu_v);
  PKM_CUDA(cudaGetLastError());
}

void launch_normalise_rows(double* a, int nrow, int64_t ncol, double log_cell, cudaStream_t st) {
  if (nrow <= 0 || ncol <= 0) return;
  k_normalise_rows<<<nrow, 256, 0, st>>>(a, ncol, log_cell);
  PKM_CUDA(cudaGetLastError());
}

void launch_assemble_txz(const double* d_log_P_t, const double* d_log_dt, const double* log_px,
                         int64_t px_age_stride, double log_dxdy, const double* d_log_Pz, const double* d_log_dz,
                         const int* row, int nt, int64_t npix, int nz, double* P_txz, double* log_p_txz,
                         int num_sms, cudaStream_t st) {
  const long long n = (long long)nt * npix * nz;
  k_assemble_txz<<<grid_for(n, num_sms), 256, 0, st>>>(d_log_P_t, d_log_dt, log_px, px_age_stride, log_dxdy,
                                                       d_log_Pz, d_log_dz, row, nt, npix, nz, P_txz, log_p_txz);
  PKM_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------------
// SSP operand pipeline: the constant operand S-hat of the hot path,
// milesSSPs.logarithmically_resample + calculate_fourier_transform
// (/root/reference/popkinmocks/model_grids.py:240-281), for all SSPs at once:
//   Xw[j][s]  = sum_{q<4} tap_w[j][q] * coef[tap_idx[j] + q][s]          cubic B-spline evaluation
//   FXw[k][s] = sum_{j<n_pix} Xw[j][s] * exp(-2 pi i (j k mod n_fft) / n_fft)      real DFT, zero padded
// The host keeps the O(n_lambda) parts (the banded not-a-knot solve for the B-spline coefficients and
// the 4 basis values per output pixel, both from SciPy); the O(n_lambda x n_ssp) evaluation and the
// O(n^2 x n_ssp) transform run here.  Any n_fft (the DFT is direct, with exact integer phase
// reduction against a long-double table), so pad="ppxf" and pad=False take the same path.
// ----------------------------------------------------------------------------
namespace {
__global__ void k_ssp_resample(const double* __restrict__ coef, const int* __restrict__ tap_idx,
                               const double* __restrict__ tap_w, int n_pix, int nssp, double* __restrict__ Xw) {
  const long long n = (long long)n_pix * nssp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / nssp), s = (int)(i - (long long)j * nssp);
    const int i0 = tap_idx[j];
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = fma(tap_w[4 * j + q], coef[(size_t)(i0 + q) * nssp + s], acc);
    Xw[i] = acc;
  }
}

// thread = (4 consecutive frequencies, one SSP); the phase index advances by k modulo n (no products)
constexpr int DFT_KB = 4;
__global__ void __launch_bounds__(128)
k_ssp_rdft(const double* __restrict__ Xw, const double2* __restrict__ wn, int n_pix, int nssp, int n_fft, int nfo,
           double2* __restrict__ FXw) {
  const int s = blockIdx.y * blockDim.x + threadIdx.x;
  const int k0 = blockIdx.x * DFT_KB;
  if (s >= nssp) return;
  int r[DFT_KB];
  double2 acc[DFT_KB];
#pragma unroll
  for (int b = 0; b < DFT_KB; ++b) {
    r[b] = 0;
    acc[b] = make_double2(0.0, 0.0);
  }
  for (int j = 0; j < n_pix; ++j) {
    const double x = __ldg(Xw + (size_t)j * nssp + s);
#pragma unroll
    for (int b = 0; b < DFT_KB; ++b) {
      const double2 w = __ldg(wn + r[b]);          // exp(-2 pi i r / n): the same address for the whole warp
      acc[b].x = fma(x, w.x, acc[b].x);
      acc[b].y = fma(x, w.y, acc[b].y);
      r[b] += k0 + b;
      if (r[b] >= n_fft) r[b] -= n_fft;
    }
  }
#pragma unroll
  for (int b = 0; b < DFT_KB; ++b)
    if (k0 + b < nfo) FXw[(size_t)(k0 + b) * nssp + s] = acc[b];
}
}  // namespace

void launch_ssp_operand(const double* d_coef, const int* d_tap_idx, const double* d_tap_w, const double2* d_wn,
                        int n_pix, int nssp, int n_fft, double* d_Xw, double2* d_FXw, int num_sms, cudaStream_t st) {
  const int nfo = n_fft / 2 + 1;
  k_ssp_resample<<<grid_for((long long)n_pix * nssp, num_sms), 256, 0, st>>>(d_coef, d_tap_idx, d_tap_w, n_pix, nssp, d_Xw);
  PKM_CUDA(cudaGetLastError());
  dim3 grid((unsigned)((nfo + DFT_KB - 1) / DFT_KB), (unsigned)((nssp + 127) / 128));
  k_ssp_rdft<<<grid, 128, 0, st>>>(d_Xw, d_wn, n_pix, nssp, n_fft, nfo, d_FXw);
  PKM_CUDA(cudaGetLastError());
}
