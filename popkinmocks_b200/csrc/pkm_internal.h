// Internal declarations shared by the translation units of libpkm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "pkm_b200.h"

// ----------------------------------------------------------------------------
// error plumbing: every C-ABI function catches everything and returns a status
// ----------------------------------------------------------------------------
struct pkm_error {
  int code;
  std::string msg;
};
void pkm_set_error(const std::string& msg);

#define PKM_THROW(code_, ...)                                  \
  do {                                                         \
    char _b[512];                                              \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                     \
    throw pkm_error{(code_), std::string(_b)};                 \
  } while (0)

#define PKM_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      int _c = (_e == cudaErrorMemoryAllocation) ? PKM_ERR_OOM : PKM_ERR_CUDA;          \
      PKM_THROW(_c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                __LINE__);                                                              \
    }                                                                                   \
  } while (0)

#define PKM_REQUIRE(cond, ...)                        \
  do {                                                \
    if (!(cond)) PKM_THROW(PKM_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ----------------------------------------------------------------------------
// device buffer that grows on demand (plan workspace)
// ----------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }        // a throw between reserve() and the explicit release() must not leak
  void reserve(size_t bytes);
  void release();
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ----------------------------------------------------------------------------
// host-built tables (tables.cpp)
// ----------------------------------------------------------------------------
struct SplineTables {
  int nfi = 0, nfo = 0;
  std::vector<double> A_inv;    // [nfi][nfi]
  std::vector<int32_t> tap_idx; // [nfo]   first coefficient of the 4-tap window
  std::vector<double> tap_w;    // [nfo][4]
};
void build_spline_tables(int nfi, int nfo, SplineTables& out);

struct FoldTables {
  int nv = 0, nfi = 0, ne = 0, no = 0;
  std::vector<double> CR;        // [nfi][ne]
  std::vector<double> CI;        // [nfi][no]
  std::vector<int32_t> idx_plus; // [ne]  velocity index of rolled position  u
  std::vector<int32_t> idx_minus;// [ne]  velocity index of rolled position -u
};
void build_fold_tables(int nv, int v0_idx, double dv, const SplineTables& sp, FoldTables& out);

// Work items of the contraction kernel: a "warp segment" is a run of at most PKM_KT consecutive
// output frequencies that share one 4-tap coefficient window (one spline span).
constexpr int PKM_KT = 25;      // frequencies per thread of the float64 contraction kernel
constexpr int PKM_KT32 = 10;    // ... of the float32 contraction kernel (weights in vector registers)
constexpr int PKM_TZC = 12;     // (t,z) rows per staged S-hat' chunk (one bulk copy)
struct WsegTables {
  int nwseg = 0, kt = 0;
  std::vector<int32_t> seg;     // [nwseg][4] = {first coefficient row, first frequency k0, count, 0}
  std::vector<double> w;        // [nwseg][kt][4] spline tap weights (zero for padding slots)
};
void build_wseg_tables(const SplineTables& sp, int kt, WsegTables& out);

// chirp-z tables for the arbitrary-length inverse real FFT
struct CztTables {
  int n = 0, nfo = 0, M = 0, logM = 0;
  int R = 1, C = 0, logC = 0;  // four-step split M = R * C used when M does not fit shared memory
  std::vector<double> chi;     // [n] complex: exp(+i*pi*j^2/n)
  std::vector<double> alpha;   // [nfo] weights 1,2,...,2,(1 if n even)
  std::vector<double> twiddle; // [M] complex: exp(-2*pi*i*j/M)
  std::vector<double> bfilt;   // [M] complex: conj chirp laid out circularly (time domain)
  // halves kernel (M = 2 * 16^p): input weights pre[h][k] = alpha_k chi_k (W_M^k)^h, h < 2, k < nfo;
  // output weights post[0][m] = chi_m, post[1][m] = +-chi_m conj(W_M^(m mod M/2)), m < n
  std::vector<double> pre, post;
  std::vector<double> htw;     // radix-16 block twiddles exp(-2 pi i j/Nb), j < Nb/16, for Nb = M/2, M/32, ..., 256
};
bool czt_halves_ok(const CztTables& c);
void build_czt_tables(int n, CztTables& out);

// ----------------------------------------------------------------------------
// the plan
// ----------------------------------------------------------------------------
enum { PKM_K_COEFF = 0, PKM_K_CONTRACT = 1, PKM_K_GAUSSIAN = 2, PKM_K_IRFFT = 3, PKM_K_TRANSPOSE = 4,
       PKM_K_NCLASS = 5 };
struct TimedLaunch {
  int cls;
  cudaEvent_t e0, e1;
};

struct pkm_plan {
  pkm_cube_desc d;
  int nfo = 0, nfi = 0, ntz = 0;
  bool density_ok = false;  // v0_idx >= 0 and nfi >= 4
  int64_t launches = 0;
  bool profiling = false;              // CUDA-event pair around every hot-path launch
  std::vector<TimedLaunch> timed;      // pending (not yet collected) launches
  double k_ms[PKM_K_NCLASS] = {0, 0, 0, 0, 0};
  int64_t k_count[PKM_K_NCLASS] = {0, 0, 0, 0, 0};
  int num_sms = 148;
  size_t ws_limit = size_t(8) << 30;

  // density path
  FoldTables fold;
  WsegTables wsegs;
  int nchunk = 0;               // ceil(ntz / PKM_TZC)
  int TM = 0, nfiP = 0;         // kernel 1: TM = 8-row output blocks, nfiP = doubles per A operand
  double* d_CRt = nullptr;      // real-part operand (A^-1 x cosine fold) in DMMA fragment order [nks][TM][32]
  double* d_CIt = nullptr;      // imaginary-part operand, same layout
  double* d_CR = nullptr;       // folded matrices, plain row major [nfi][ne] / [nfi][no] (generic kernel)
  double* d_CI = nullptr;
  double* d_CIt_oct = nullptr;  // the same with the rolled index shifted by one (octet kernel: O[u] at column u)
  double2* d_Sw = nullptr;      // [nwseg][nchunk*TZC][KT]  S-hat' = FXw*dz*dt, tz = t*nz+z, zero padded
  float2* d_Sw32 = nullptr;     // float32 copy of d_Sw (precision = 1 only)
  int32_t* d_wseg = nullptr;    // [nwseg][4]
  double* d_wseg_w = nullptr;   // [nwseg][KT][4]

  // gaussian path
  int ga_nks = 0, ga_nkt = 0;   // k-steps of 4 metallicities, frequency tiles of 64
  double* d_SA = nullptr;       // FXw (unscaled) in DMMA fragment order [nkt][nt][16][nks][32]
  float2* d_SA32 = nullptr;     // FXw as float2 [nkt][nt][64][nz] (precision = 1 only)
  double* d_omega = nullptr;    // [nfo] the reference's omega array
  std::vector<double> h_omega;  // host copy of what d_omega holds (re-uploaded only when it changes)

  // inverse FFT
  CztTables czt;
  double2* d_chi = nullptr;
  double* d_alpha = nullptr;
  double2* d_twiddle = nullptr;
  double2* d_Bhat = nullptr;    // [M] forward transform of bfilt, in the kernel's permuted order
  double2* d_Bhat_h = nullptr;  // [M] the same in the order of the halves kernel (M = 2 * 16^p only)
  double2* d_pre = nullptr;     // [2][nfo] input weights of the halves kernel
  double2* d_post = nullptr;    // [2][n] output weights of the halves kernel
  double2* d_htw = nullptr;     // block twiddles of the halves kernel
  double2* d_wn = nullptr;      // [n] exp(+2 pi i r / n) for the naive cross-check transform

  double* d_exp_tab = nullptr;  // [256] 2^(j/256) for exp_tab()

  // statistics
  double* d_log_dt = nullptr;   // [nt]
  double* d_log_dz = nullptr;   // [nz]
  std::vector<double> h_delta_t, h_delta_z;

  // workspace
  DevBuf coef, yhat, ypx, stage_in[2], stage_out[2], cube_stage[2], misc, small;
  // host-buffer pipeline: s_copy = H2D stream, s_comp = D2H stream; per staging slot:
  // ev_in (H2D landed), ev_done (kernel 1 consumed it), ev_ready (output staged), ev_out (D2H done)
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  // per tile lane (tiles alternate between two lanes whose kernels fill each other's wave tails):
  // side stream of the contraction's alternate launches + its fork / join events
  cudaStream_t s_k2[2] = {nullptr, nullptr};
  cudaEvent_t ev_k2_fork[2] = {nullptr, nullptr}, ev_k2_join[2] = {nullptr, nullptr};
  cudaStream_t s_lane = nullptr;         // main stream of lane 1 (lane 0 runs on the caller's stream)
  cudaEvent_t ev_lane_fork = nullptr, ev_lane_join = nullptr;
  DevBuf coef1, yhat1, ypx1;             // lane 1's workspace (lane 0: coef, yhat, ypx)
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
              ev_out[2] = {nullptr, nullptr}, ev_ready[2] = {nullptr, nullptr};
};

// ----------------------------------------------------------------------------
// kernel launchers (all asynchronous on `st`)
// ----------------------------------------------------------------------------
// kernel 1: density tile -> spline coefficients cT[x/32][t*nz+z][m][x%32] (complex128, or float2 for precision 1)
void launch_coeff(pkm_plan* pl, const double* dens, int is_log, int64_t npix_total,
                  int64_t pix0, int npix_tile, int xpad, double2* cT, cudaStream_t st);
// kernel 2: coefficients -> Yhat[x][nfo]
// pixel groups (of 32) that fill the GPU once with contraction CTAs (real valued)
double contract_wave_pixel_groups(const pkm_plan* pl);
void launch_contract(pkm_plan* pl, const double2* cT, int npix_tile, int xpad, double2* yhat,
                     cudaStream_t st, int lane = 0);
// gaussian path -> Yhat[x][nfo]
void launch_gaussian(pkm_plan* pl, const double* P_txz, const double* mu, const int64_t* mu_st,
                     const double* sig, const int64_t* sig_st, int64_t nx2, int64_t npix_total,
                     int64_t pix0, int npix_tile, int xpad, double omega_step4, double2* yhat,
                     cudaStream_t st);
void gaussian_tables_upload(pkm_plan* pl, const double* FXw);
// inverse real FFT: Yhat[npix][nfo] -> out[npix][n] * scale
void launch_irfft_czt(pkm_plan* pl, const double2* yhat, int64_t npix, double scale, double* out,
                      cudaStream_t st);
void launch_irfft_naive(pkm_plan* pl, const double2* yhat, int64_t npix, double scale, double* out,
                        cudaStream_t st);
void czt_prepare_filter(pkm_plan* pl);  // computes d_Bhat on the device (plan creation)
// [npix][n] -> cube[m*row_stride + col0 + x]
void launch_transpose(const double* src, int64_t npix, int64_t n, double* dst, int64_t row_stride,
                      int64_t col0, cudaStream_t st);
void launch_axpy(int n_cmps, const double* const* d_ptrs_dev, const double* d_w, int64_t n,
                 double* out, cudaStream_t st);
void launch_logsumexp_cubes(int n_cmps, const double* const* d_ptrs_dev, const double* d_log_w, int64_t n,
                            double* out, cudaStream_t st);

// statistics kernels (stats.cu)
// Source of a log-sum-exp reduction: the 5-D log density (weighted = 1: the log volume elements
// are added on the fly) or an already volume-weighted log-probability marginal of it (weighted = 0),
// whose summed-out axes have extent 1.
struct ReduceSrc {
  int nt, nv, nz;
  int64_t npix;
  int weighted;
  int uniform_dv = 0;   // all velocity bins have the same width (log dv folded into the per-thread constant)
};
void launch_reduce_logp(pkm_plan* pl, const double* logp, const ReduceSrc& src, const double* d_log_dv,
                        const double* d_log_lw, uint32_t keep_mask, int density, double* out,
                        cudaStream_t st);
void launch_apply_log_dv(pkm_plan* pl, double* out, const ReduceSrc& src, const double* d_log_dv,
                         uint32_t keep_mask, double sign, cudaStream_t st);
void launch_conditional(pkm_plan* pl, const double* log_joint, uint32_t joint_mask, const double* log_given,
                        uint32_t given_mask, int64_t npix, const double* d_log_dv, int density, int do_exp,
                        double* out, cudaStream_t st);
void launch_moments_logp(const double* log_joint, const double* log_given, const double* values,
                         int64_t nouter, int64_t nvar, int64_t ninner, double* out, cudaStream_t st);
void density_tables_upload(pkm_plan* pl);
int coeff_choose_pxb(const pkm_plan* pl, int* ngroup_out = nullptr);
void launch_moments(const double* P, const double* values, int64_t nvar, int64_t ncond,
                    double* out, cudaStream_t st);
void launch_max(const double* x, int64_t n, double* d_result, cudaStream_t st);
void launch_noise(const double* ybar, int64_t n, double snr, int mode, uint64_t seed,
                  uint64_t offset, const double* d_max, double* yobs, double* sigma,
                  cudaStream_t st);

void launch_hist5d(const double* const samples[5], const double* weights, int64_t n, const double* d_edges,
                   const int nbin[5], double* out, int num_sms, cudaStream_t st);
void launch_pixelate(const double* log_w, const double* mu, const double* sig, const double* d_v_edg,
                     const double* d_log_dv, int nt, int nv, int64_t npix, int nz, double* out,
                     cudaStream_t st);
void launch_synth_logp(double* out, int nt, int nv, int64_t total_pix, int64_t pix0, int64_t npix, int nz,
                       uint64_t seed, double scale, int num_sms, cudaStream_t st);
// parametric-model maps (model_kernels.cu)
struct ModelFrame {
  double cx, cy, c, s;   // centre, cos / sin of the rotation
  int nx1, nx2;
};
void launch_disk_maps(const ModelFrame& mf, const double* d_x1, const double* d_x2, const double* d_age_pars,
                      int nt, double scale, double* log_rho, double* mu_v, double* sig_v, int num_sms,
                      cudaStream_t st);
void launch_disk_t_dep(const ModelFrame& mf, const double* d_x1, const double* d_x2, double q, double alpha,
                       double t_in, double t_out, double scale, const double* d_grid, int ngrid, double* t_dep,
                       int* row, int num_sms, cudaStream_t st);
void launch_stream_maps(const ModelFrame& mf, const double* d_x1, const double* d_x2, const double* d_cx,
                        const double* d_cy, int nsmp, double sig, double hwx, double hwy, double th_lo,
                        double th_hi, double v_lo, double v_hi, double* log_p, double* mu_v, int num_sms,
                        cudaStream_t st);
void launch_normalise_rows(double* a, int nrow, int64_t ncol, double log_cell, cudaStream_t st);
void launch_assemble_txz(const double* d_log_P_t, const double* d_log_dt, const double* log_px,
                         int64_t px_age_stride, double log_dxdy, const double* d_log_Pz, const double* d_log_dz,
                         const int* row, int nt, int64_t npix, int nz, double* P_txz, double* log_p_txz,
                         int num_sms, cudaStream_t st);
void launch_ssp_operand(const double* d_coef, const int* d_tap_idx, const double* d_tap_w, const double2* d_wn,
                        int n_pix, int nssp, int n_fft, double* d_Xw, double2* d_FXw, int num_sms, cudaStream_t st);
bool pkm_is_device_pointer(const void* p);

// RAII CUDA-event bracket of one launch on its own stream (no-op unless plan->profiling)
struct KernelTimer {
  pkm_plan* pl;
  cudaStream_t st;
  TimedLaunch t;
  bool on;
  KernelTimer(pkm_plan* pl_, int cls, cudaStream_t st_) : pl(pl_), st(st_), on(pl_->profiling) {
    if (!on) return;
    t.cls = cls;
    cudaEventCreate(&t.e0);
    cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, st);
  }
  ~KernelTimer() {
    if (!on) return;
    cudaEventRecord(t.e1, st);
    pl->timed.push_back(t);
  }
};
// NVTX range (header-only NVTX 3: a no-op unless a profiler is attached) around the hot-path calls,
// so nsys / ncu timelines show the entry points and their per-tile kernel groups by name
struct PkmRange {
  explicit PkmRange(const char* name) { nvtxRangePushA(name); }
  ~PkmRange() { nvtxRangePop(); }
  PkmRange(const PkmRange&) = delete;
  PkmRange& operator=(const PkmRange&) = delete;
};
double run_fma_probe(int num_sms, int repeats);  // TFLOP/s
