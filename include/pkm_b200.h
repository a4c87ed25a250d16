/*
 * pkm_b200.h - C ABI of the B200-native popkinmocks datacube hot path.
 *
 * Shared library: popkinmocks_b200/csrc/libpkm_b200.so  (built by `make -C popkinmocks_b200/csrc`
 * or `__graft_entry__.build()`; nvcc -gencode arch=compute_100a,code=sm_100a).
 *
 * Every entry point replaces one piece of the reference's pure-Python hot path
 * (reference = prashjet/popkinmocks v1.0.1; file:line relative to /root/reference):
 *
 *   pkm_ybar_density   <- Component.evaluate_ybar            popkinmocks/components/base.py:789-882
 *   pkm_ybar_gaussian  <- ParametricComponent.evaluate_ybar  popkinmocks/components/parametric.py:335-407
 *   pkm_axpy_cubes     <- Mixture.evaluate_ybar              popkinmocks/components/mixture.py:35-48
 *   pkm_reduce_logp    <- Component._get_log_p_* marginals   popkinmocks/components/base.py:884-1174
 *   pkm_moments        <- get_mean/get_noncentral_moment ... popkinmocks/components/base.py:179-452
 *   pkm_noise          <- ShotNoise/ConstantSNR + Diagonal.sample  popkinmocks/noise.py:16-71
 *   pkm_plan_create    <- the constant operands read from IFUCube / milesSSPs
 *                         (ssps.FXw, par_dims, n_fft, dv, delta_t, delta_z; cube.dx, dy, v_edg)
 *                         popkinmocks/ifu_cube.py:20-50, popkinmocks/model_grids.py:240-293
 *
 * Conventions
 *   - plain C types only; all arrays are C-order float64 unless stated; complex = interleaved (re, im).
 *   - bulk array arguments documented "host or device" are classified with
 *     cudaPointerGetAttributes; host buffers are staged by the library (pinned
 *     or pageable both work, pinned is faster).
 *   - every function returns PKM_OK (0) or a negative pkm_status; pkm_last_error()
 *     returns a thread-local message for the last failure.  No C++ exception crosses the ABI.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls with device buffers are asynchronous on `stream`; calls with any host
 *     buffer return after the result is in the host buffer.
 *   - there is NO CPU fallback: without a CUDA device the compute entry points
 *     return PKM_ERR_CUDA.
 */
#ifndef PKM_B200_H
#define PKM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PKM_ABI_VERSION 1

#if defined(__GNUC__)
#define PKM_API __attribute__((visibility("default")))
#else
#define PKM_API
#endif

typedef enum pkm_status {
  PKM_OK = 0,
  PKM_ERR_INVALID = -1,   /* bad argument / shape / unsupported size */
  PKM_ERR_CUDA = -2,      /* CUDA runtime error (message holds cudaGetErrorString) */
  PKM_ERR_OOM = -3,       /* device or host allocation failed */
  PKM_ERR_NO_ZERO_BIN = -4 /* density path needs a velocity bin centred on 0 (reference raises IndexError, base.py:822-823) */
} pkm_status;

/* Output layouts of the ybar entry points. */
#define PKM_LAYOUT_CUBE 0      /* [n_fft][rows][nx2]  - the reference layout (omega slowest)      */
#define PKM_LAYOUT_PIXEL_MAJOR 1 /* [rows*nx2][n_fft] - spaxel-major, used for the multi-GPU gather */

/* Constant description of one (cube, ssps) pair.  Pixels are a batch dimension and
 * are given per call, so one plan serves any nx1 x nx2 with the same grids. */
typedef struct pkm_cube_desc {
  int32_t nt;      /* number of SSP ages            (ssps.par_dims[1]) */
  int32_t nz;      /* number of SSP metallicities   (ssps.par_dims[0]) */
  int32_t nv;      /* number of velocity bins       (cube.nv)          */
  int32_t n_fft;   /* ssps.n_fft (= len(ssps.w) with pad=False); any length, primes included */
  int32_t v0_idx;  /* index of the velocity bin centred on 0 (base.py:842); -1 if there is none */
  int32_t device;  /* CUDA device ordinal */
  int32_t precision; /* 0 = float64 arithmetic (default); 1 = float32 arithmetic (inputs/outputs stay float64) */
  int32_t reserved;
  double dv;       /* ssps.dv  */
  double dx, dy;   /* cube.dx, cube.dy */
} pkm_cube_desc;

typedef struct pkm_plan pkm_plan;

PKM_API const char* pkm_last_error(void);
PKM_API int32_t pkm_abi_version(void);
/* Number of visible CUDA devices (0 without a GPU/driver; never fails). */
PKM_API int32_t pkm_device_count(void);

/* Build the device-resident constant operands.
 *   FXw      host, complex128 [n_fft/2+1][nz][nt]  (ssps.FXw reshaped, model_grids.py:277-281)
 *   delta_t  host [nt], delta_z host [nz]
 * Uploads S-hat' = FXw*dz*dt (density path) and FXw (Gaussian path) in kernel-friendly
 * layouts, builds the not-a-knot spline tables, the folded DFT matrices, chirp-z tables. */
PKM_API int32_t pkm_plan_create(const pkm_cube_desc* desc, const double* FXw, const double* delta_t,
                        const double* delta_z, pkm_plan** plan_out);
PKM_API int32_t pkm_plan_destroy(pkm_plan* plan);
/* Cap the device workspace (bytes) used for intermediate spline coefficients; the
 * pixel batch is tiled to fit.  0 restores the default (8 GiB). */
PKM_API int32_t pkm_plan_set_workspace_limit(pkm_plan* plan, int64_t bytes);
/* Number of kernels launched by this plan since creation (for bench accounting). */
PKM_API int64_t pkm_plan_launch_count(const pkm_plan* plan);

/* Per-kernel device timing.  While enabled, every hot-path launch of this plan is bracketed by
 * a CUDA-event pair recorded on the launch stream.  pkm_plan_kernel_times waits for the pending
 * launches and returns, per kernel class, the summed milliseconds and the launch count since
 * profiling was (re-)enabled.  Classes: 0 coefficient GEMM, 1 spline contraction, 2 Gaussian
 * contraction, 3 inverse FFT, 4 transpose. */
#define PKM_NUM_KERNEL_CLASSES 5
PKM_API int32_t pkm_plan_set_profiling(pkm_plan* plan, int32_t enable);
PKM_API int32_t pkm_plan_kernel_times(pkm_plan* plan, double* ms_out, int64_t* count_out);

/* Density path.  dens: [nt][nv][nx1][nx2][nz] float64, host or device; is_log != 0 means
 * it holds log p (exp is fused into the first kernel; -inf -> 0).  Rows [row_begin,row_end)
 * of the x1 axis are evaluated (spatial sharding); ybar_out (host or device) receives
 * (row_end-row_begin)*nx2*n_fft float64 in `out_layout`. */
PKM_API int32_t pkm_ybar_density(pkm_plan* plan, const double* dens, int32_t is_log, int32_t nx1,
                         int32_t nx2, int32_t row_begin, int32_t row_end, double* ybar_out,
                         int32_t out_layout, void* stream);

/* Multi-GPU variant of pkm_ybar_density: rows [row_begin,row_end) are written straight into a
 * LARGER cube [n_fft][cube_nx1][nx2] at row offset cube_row0.  `cube` is device memory of this GPU
 * or of a PEER GPU (a pointer obtained with pkm_ipc_open): every pixel tile is transposed locally
 * and moved into the cube by one strided 2-D copy-engine transfer over NVLink on a side stream,
 * overlapped with the next tile's kernels, so the cube of a sharded evaluation is assembled on
 * rank 0 without a separate gather pass (each rank needs one barrier before rank 0 reads). */
PKM_API int32_t pkm_ybar_density_into_cube(pkm_plan* plan, const double* dens, int32_t is_log,
                                           int32_t nx1, int32_t nx2, int32_t row_begin,
                                           int32_t row_end, double* cube, int32_t cube_nx1,
                                           int32_t cube_row0, void* stream);
/* CUDA IPC plumbing for the peer-memory cube (one process per GPU): export a 64-byte handle of a
 * device allocation's BASE address, open it in another process, close it. */
PKM_API int32_t pkm_device_alloc(int32_t device, int64_t bytes, void** ptr_out); /* plain cudaMalloc: IPC-exportable */
PKM_API int32_t pkm_device_free(void* ptr);
PKM_API int32_t pkm_ipc_export(const void* dev_ptr, uint8_t* handle64);
PKM_API int32_t pkm_ipc_open(const uint8_t* handle64, int32_t device, void** ptr_out);
PKM_API int32_t pkm_ipc_close(void* ptr);

/* Page-lock a caller-owned host buffer in place (cudaHostRegister) so that the pipelined staging of
 * pkm_ybar_density / pkm_reduce_logp reads it at full PCIe rate; a buffer that is already pinned
 * is left alone.  The Python host side registers a component's log_p_tvxz on first use and
 * unregisters it when the array is garbage collected. */
PKM_API int32_t pkm_host_register(void* ptr, int64_t bytes);
PKM_API int32_t pkm_host_unregister(void* ptr);

/* Gaussian-LOSVD path.  P_txz [nt][nx1][nx2][nz] (volume-weighted probabilities),
 * mu_v / sig_v addressed as base[t*s[0] + x1*s[1] + x2*s[2]] with ELEMENT strides
 * (0 strides allow numpy broadcast views, stream.py:126); omega [n_fft/2+1] is the
 * reference's linspace(0, pi, nfo)/dv (parametric.py:360-362), host.  All of P_txz,
 * mu_v, sig_v must be on the same side (all host or all device). */
PKM_API int32_t pkm_ybar_gaussian(pkm_plan* plan, const double* P_txz, const double* mu_v,
                          const int64_t mu_strides[3], const double* sig_v,
                          const int64_t sig_strides[3], const double* omega, int32_t nx1,
                          int32_t nx2, int32_t row_begin, int32_t row_end, double* ybar_out,
                          int32_t out_layout, void* stream);

/* out[i] = sum_c weights[c] * ybars[c][i], i < n.  ybars: host array of n_cmps pointers
 * (all host or all device, same side as out); weights host. */
PKM_API int32_t pkm_axpy_cubes(int32_t n_cmps, const double* const* ybars, const double* weights,
                       int64_t n, double* out, int32_t device, void* stream);

/* Component-collapsed log density of a Mixture (Mixture.get_log_p(..., collapse_cmps=True),
 * popkinmocks/components/mixture.py:50-80): out[i] = log sum_c exp(log_weights[c] + log_ps[c][i]).
 * log_ps: n_cmps pointers to n float64 each; log_ps and out host or device (same side),
 * log_weights [n_cmps] host. */
PKM_API int32_t pkm_logsumexp_cubes(int32_t n_cmps, const double* const* log_ps, const double* log_weights,
                                    int64_t n, double* out, int32_t device, void* stream);

/* Spaxel-major [npix][n] -> cube layout [n][npix] (device pointers only). */
PKM_API int32_t pkm_transpose_to_cube(const double* pixel_major, int64_t npix, int64_t n, double* cube,
                              int32_t device, void* stream);

/* Batched inverse real FFT of arbitrary length (numpy.fft.irfft semantics, base.py:858):
 * Yhat complex128 [npix][n_fft/2+1] device -> out float64 [npix][n_fft] device, times `scale`.
 * Exposed for tests; pkm_ybar_* call the same kernels. */
PKM_API int32_t pkm_irfft_batch(pkm_plan* plan, const double* Yhat, int64_t npix, double scale,
                        double* out, void* stream);
/* O(n^2) inverse real DFT with exact integer phase reduction - independent cross-check. */
PKM_API int32_t pkm_irfft_naive(pkm_plan* plan, const double* Yhat, int64_t npix, double scale,
                        double* out, void* stream);

/* log-marginals of the pixelated density (base.py:884-1174, 1017-1041).
 *   log_p_tvxz [nt][nv][npix][nz] host or device; log_dv host [nv] = log(velocity bin widths);
 *   light_weights host [nt][nz] or NULL (mass-weighted);
 *   keep_mask bit0=t bit1=v bit2=x bit3=z: axes kept in the output (others are
 *   log-sum-exp'ed); out (same side as input) has the kept axes in t,v,x,z order.
 *   Returns log P (volume weighted) if density == 0, else log density. */
PKM_API int32_t pkm_reduce_logp(pkm_plan* plan, const double* log_p_tvxz, int64_t npix,
                        const double* log_dv, const double* light_weights, uint32_t keep_mask,
                        int32_t density, double* out, void* stream);

/* Several log-marginals of the same array in one call.  keep_masks / density [n_out] and the output
 * pointers outs [n_out] are host arrays; every output is laid out like pkm_reduce_logp's.  Nested
 * requests share work: a marginal whose axes are a subset of another requested one is reduced from
 * that (small) result instead of the 5-D array, so e.g. {vx, x, v} or {txz, tz, t, z, x} read the
 * 5-D array ONCE.  (The reductions are bound by one exponential per source element, not by HBM,
 * so requests that are not nested each make their own pass.)  All outputs share the weighting
 * (light_weights NULL = mass weighted).  n_out <= 16. */
PKM_API int32_t pkm_reduce_logp_multi(pkm_plan* plan, const double* log_p_tvxz, int64_t npix,
                                      const double* log_dv, const double* light_weights, int32_t n_out,
                                      const uint32_t* keep_masks, const int32_t* density,
                                      double* const* outs, void* stream);

/* A log-marginal reduced from an already computed one.  log_P_src (device) holds volume-weighted
 * log PROBABILITIES (density = 0 output of pkm_reduce_logp*) over the axes of src_mask; keep_mask
 * must be a subset.  light_weights != NULL applies log L(t,z) and renormalises (the source must be
 * mass weighted and keep t and z): this is how light-weighted marginals without v follow from the
 * mass-weighted p(t,x,z) without touching the 5-D array again.  Device buffers only. */
PKM_API int32_t pkm_reduce_logp_from(pkm_plan* plan, const double* log_P_src, uint32_t src_mask, int64_t npix,
                                     const double* log_dv, const double* light_weights, uint32_t keep_mask,
                                     int32_t density, double* out, void* stream);

/* A marginal or conditional distribution assembled on the device from log-probability marginals
 * (base.py:130-177): out[dependent axes..., conditioning axes...] = log P(joint) - log P(given)
 * - (density ? log dV(dependent axes) : 0), exponentiated if `exponentiate` (get_p instead of
 * get_log_p).  log_P_joint holds the axes of joint_mask, log_P_given (or NULL for a plain marginal)
 * those of given_mask (a subset), both as produced by pkm_reduce_logp* with density = 0; axes inside
 * each group are in t, v, x, z order, which is the reference's array layout.  Device buffers only. */
PKM_API int32_t pkm_conditional(pkm_plan* plan, const double* log_P_joint, uint32_t joint_mask,
                                const double* log_P_given, uint32_t given_mask, int64_t npix,
                                const double* log_dv, int32_t density, int32_t exponentiate, double* out,
                                void* stream);

/* Moments of a conditional 1-D distribution straight from device-resident log marginals
 * (base.py:130-177 + 208-260 in one kernel): P_i = exp(log_joint[a][i][b] - log_given[a][b]) with the
 * variable on the middle axis (nouter x nvar x ninner, C order; log_given [nouter][ninner] or NULL),
 * out[4][nouter*ninner] = mean and central moments 2..4.  Device buffers only. */
PKM_API int32_t pkm_moments_logp(const double* log_joint, const double* log_given, const double* values,
                                 int64_t nouter, int64_t nvar, int64_t ninner, double* out, int32_t device,
                                 void* stream);

/* Moments of a conditional 1-D distribution held as probabilities P[nvar][ncond]
 * (axis 0 = variable, base.py:208-260): out[0][c]=mean, out[1..3][c]= central moments 2,3,4.
 * P, values[nvar], out[4][ncond]: host or device (same side). */
PKM_API int32_t pkm_moments(const double* P, const double* values, int64_t nvar, int64_t ncond,
                    double* out, int32_t device, void* stream);

/* Observation noise (noise.py:16-71).  mode 0 = ShotNoise (sigma = sqrt(max ybar)/snr * sqrt(ybar)),
 * mode 1 = ConstantSNR (sigma = ybar/snr).  yobs = ybar + N(0, sigma) (Philox counter RNG,
 * seeded); sigma_out may be NULL.  ybar_max < 0 => the library reduces max(ybar) itself;
 * otherwise it is used as given (multi-GPU: pass the all-reduced maximum).
 * ybar, yobs, sigma_out: host or device (same side). */
PKM_API int32_t pkm_noise(const double* ybar, int64_t n, double snr, int32_t mode, uint64_t seed,
                  uint64_t offset, double ybar_max, double* yobs, double* sigma_out,
                  int32_t device, void* stream);
/* max over a device/host float64 buffer (NaN-propagating like numpy.max). */
PKM_API int32_t pkm_max(const double* x, int64_t n, double* result_host, int32_t device, void* stream);

/* 5-D particle histogram with numpy.histogramdd bin semantics (FromParticle,
 * popkinmocks/components/particle.py:52-54): samples[a] (a = t, v, x1, x2, z; n values each, host
 * or device, all on the same side), edges[a] host arrays of nbin[a]+1 ascending bin edges.
 * out [nbin0][nbin1][nbin2][nbin3][nbin4] float64 (host or device) receives the counts, or the
 * sums of `weights` (may be NULL).  Bin = searchsorted(edges, x, "right") - 1, x == last edge
 * falls in the last bin, outside / NaN is dropped: indices are bit-exact w.r.t. NumPy. */
PKM_API int32_t pkm_histogram5d(const double* const* samples, const double* weights, int64_t n,
                                const double* const* edges, const int32_t* nbin, double* out,
                                int32_t device, void* stream);

/* Analytic pixelation of a Gaussian-LOSVD component on the device
 * (ParametricComponent._get_log_p_tvxz, popkinmocks/components/parametric.py:567-596, with the
 * velocity factor of _get_log_p_v_tx, parametric.py:180-211):
 *   out[t][v][x][z] = log_w_txz[t][x][z]
 *                   + log( Phi((v_edg[v+1]-mu[t][x])/sig[t][x]) - Phi((v_edg[v]-mu[t][x])/sig[t][x]) )
 *                   - (density ? log(v_edg[v+1]-v_edg[v]) : 0)
 * log_w_txz [nt][npix][nz] is log p(t,x,z) (density or volume-weighted, as the caller wants the
 * result), mu_v / sig_v [nt][npix] dense, v_edg [nv+1] a host array.  log_w_txz, mu_v, sig_v and
 * out [nt][nv][npix][nz] may each be host or device memory; a device `out` makes the 5-D array
 * (20.8 GB at 201x201 px) exist only in HBM, ready for pkm_ybar_density / pkm_reduce_logp. */
PKM_API int32_t pkm_pixelate_gaussian(const double* log_w_txz, const double* mu_v, const double* sig_v,
                                      const double* v_edg, int32_t nt, int32_t nv, int64_t npix,
                                      int32_t nz, int32_t density, double* out, int32_t device,
                                      void* stream);

/* Synthetic white-noise log density of the benchmark workload (SURVEY.md section 8d, config C4:
 * uniform(0,1) values; the reference has no generator of its own - its tests draw from NumPy).
 * out (device) receives [nt][nv][npix][nz] float64 for the pixel range [pix0, pix0+npix) of a
 * cube with total_pix pixels: out = log(u * scale), u = (splitmix64(seed + g) >> 11 + 0.5) * 2^-53,
 * g = the element's flat index in the WHOLE [nt][nv][total_pix][nz] array, so the values of a
 * pixel do not depend on how the cube is sharded. */
PKM_API int32_t pkm_synth_log_density(double* out, int32_t nt, int32_t nv, int64_t total_pix, int64_t pix0,
                                      int64_t npix, int32_t nz, uint64_t seed, double scale,
                                      int32_t device, void* stream);

/* ---- parametric-model maps on the device -----------------------
 * The O(pixels) inputs of the Gaussian-LOSVD path, generated in HBM from O(nt) parameter vectors.
 * x1 [nx1], x2 [nx2]: pixel centres (host); (cx, cy), rotation: the component frame
 * (popkinmocks/components/parametric.py:43-49).  All outputs are DEVICE buffers.
 *
 * pkm_disk_maps <- GrowingDisk.set_p_x_t / set_mu_v / set_sig_v / set_t_dep
 *                  (popkinmocks/components/growing_disk.py:30-239)
 *   age_pars host [10][nt]: rows q, rc, alpha of p(x|t); q, rc, K of mu_v; q, alpha, sig_in, sig_out of
 *   sig_v (already interpolated in age on the host); scale = max |x'| over the cube.
 *   log_p_x_t [nt][npix] receives the log DENSITY p(x|t) normalised per age (log_dxdy = log(dx dy)),
 *   mu_v, sig_v [nt][npix].  If t_dep and row_idx are given: t_dep_pars host {q, alpha, t_dep_in,
 *   t_dep_out}, t_dep_grid host [ngrid] ascending (the tabulated depletion times of
 *   parametric.py:213-247); t_dep [npix] and the index of the nearest grid entry row_idx [npix]
 *   (parametric.py:262, first minimum on ties). */
PKM_API int32_t pkm_disk_maps(const double* x1, const double* x2, int32_t nx1, int32_t nx2, double cx, double cy,
                              double rotation, const double* age_pars, int32_t nt, double scale,
                              const double* t_dep_pars, const double* t_dep_grid, int32_t ngrid,
                              double* log_p_x_t, double* mu_v, double* sig_v, double* t_dep, int32_t* row_idx,
                              double log_dxdy, int32_t device, void* stream);
/* pkm_stream_maps <- Stream.set_p_x_t / set_mu_v (popkinmocks/components/stream.py:28-127):
 *   centres_x/y host [nsmp] = r(theta) cos/sin(theta) of the samples along the arc, sig their width,
 *   half_dx/dy = half the pixel size.  log_p_x [npix] receives log P(pixel) normalised to one over the
 *   cube (the caller subtracts log(dx dy) for the density); mu_v [npix] the piecewise-linear velocity. */
PKM_API int32_t pkm_stream_maps(const double* x1, const double* x2, int32_t nx1, int32_t nx2, double cx, double cy,
                                double rotation, const double* centres_x, const double* centres_y, int32_t nsmp,
                                double sig, double half_dx, double half_dy, double theta_lo, double theta_hi,
                                double v_lo, double v_hi, double* log_p_x, double* mu_v, double log_dxdy,
                                int32_t device, void* stream);
/* pkm_assemble_txz <- ParametricComponent.get_p("txz", density=False) and get_log_p("txz")
 * (popkinmocks/components/parametric.py:541-566):
 *   P[t][x][z] = exp(log_P_t[t] + log_p_x[t or 0][x] + log_dxdy + log_Pz_rows[row_idx[x]][t][z])
 *   log_p[t][x][z] = the corresponding log density (log_dt, log_dz host = log bin widths).
 * log_P_t host [nt] (volume weighted), log_p_x device [nt][npix] (px_per_age != 0) or [npix],
 * log_Pz_rows host [nrow][nt][nz] = log P(z | t, depletion-time row) (volume weighted), row_idx device
 * [npix] or NULL (nrow == 1).  P_txz / log_p_txz: device [nt][npix][nz], either may be NULL. */
PKM_API int32_t pkm_assemble_txz(const double* log_P_t, const double* log_dt, const double* log_p_x,
                                 int32_t px_per_age, double log_dxdy, const double* log_Pz_rows, int32_t nrow,
                                 const double* log_dz, const int32_t* row_idx, int32_t nt, int64_t npix,
                                 int32_t nz, double* P_txz, double* log_p_txz, int32_t device, void* stream);

/* SSP operand pipeline on the device <- milesSSPs.logarithmically_resample
 * + calculate_fourier_transform (popkinmocks/model_grids.py:240-281):
 *   Xw[j][s]  = sum_{q<4} tap_w[j][q] * coef[tap_idx[j]+q][s]      (cubic B-spline evaluated at the uniform
 *               log-wavelength grid; coef [ncoef][nssp] = not-a-knot B-spline coefficients of the SSP
 *               spectra over log(lambda), tap_idx/tap_w [n_pix]([4]) = first coefficient and the four
 *               basis values of every output pixel - the O(n_lambda) tables come from the host)
 *   FXw[k][s] = sum_j Xw[j][s] exp(-2 pi i j k / n_fft), k <= n_fft/2   (= numpy.fft.rfft(Xw.T, n_fft).T; any
 *               n_fft >= n_pix, so pad=False and pad="ppxf" are the same call)
 * coef, tap_idx, tap_w: host.  Xw_out [n_pix][nssp] float64, FXw_out [n_fft/2+1][nssp] complex128: host or
 * device. */
PKM_API int32_t pkm_ssp_operand(const double* coef, int32_t ncoef, const int32_t* tap_idx, const double* tap_w,
                                int32_t n_pix, int32_t nssp, int32_t n_fft, double* Xw_out, double* FXw_out,
                                int32_t device, void* stream);

/* Measured FP64 FMA throughput of `device` in TFLOP/s (2 flops per FMA; best of `repeats`
 * timed launches of a register-resident DFMA chain on every SM).  The roofline denominator of
 * the FP64-pipe-bound kernels; MEASURED_PEAKS.json carries no FP64 figure. */
PKM_API int32_t pkm_fp64_fma_peak(int32_t device, int32_t repeats, double* tflops_out);

/* Host-only helpers (no GPU needed), exposed so the tables can be checked on CPU.
 * Not-a-knot cubic spline nfi -> nfo on linspace(0,1,.) (base.py:846-853):
 *   A_inv [nfi][nfi]; tap_idx [nfo] first coefficient of the 4-tap window; tap_w [nfo][4]. */
PKM_API int32_t pkm_spline_tables(int32_t nfi, int32_t nfo, double* A_inv, int32_t* tap_idx,
                          double* tap_w);
/* The folded matrices of the first density kernel: CR [nfi][nv/2+1], CI [nfi][(nv-1)/2] with
 * c = CR * e + i * CI * o, e/o = even/odd parts of the rolled velocity profile. */
PKM_API int32_t pkm_fold_tables(int32_t nv, int32_t v0_idx, double dv, double* CR, double* CI,
                        int32_t* idx_plus, int32_t* idx_minus);

#ifdef __cplusplus
}
#endif
#endif /* PKM_B200_H */
