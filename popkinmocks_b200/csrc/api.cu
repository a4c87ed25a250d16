// C-ABI entry points of libpkm_b200.so (declared in include/pkm_b200.h).
//
// This file owns the plan (constant operands on the device), the pixel tiling of the two
// ybar paths, the host<->device staging for callers that pass host buffers, and the error
// plumbing.  All arithmetic lives in the kernel translation units.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "exp_tab.cuh"
#include "pkm_internal.h"

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
static thread_local std::string g_last_error;
void pkm_set_error(const std::string& msg) { g_last_error = msg; }

template <class F>
static int32_t guarded(F&& body) {
  try {
    body();
    return PKM_OK;
  } catch (const pkm_error& e) {
    pkm_set_error(e.msg);
    return e.code;
  } catch (const std::bad_alloc&) {
    pkm_set_error("host allocation failed");
    return PKM_ERR_OOM;
  } catch (const std::exception& e) {
    pkm_set_error(std::string("unexpected: ") + e.what());
    return PKM_ERR_INVALID;
  } catch (...) {
    pkm_set_error("unknown error");
    return PKM_ERR_INVALID;
  }
}

void DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return;
  release();
  PKM_CUDA(cudaMalloc(&p, bytes));
  cap = bytes;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

bool pkm_is_device_pointer(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory on old drivers
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static void use_device(int dev) { PKM_CUDA(cudaSetDevice(dev)); }

template <class T>
static T* upload(const std::vector<T>& h) {
  T* d = nullptr;
  PKM_CUDA(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) PKM_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

// ----------------------------------------------------------------------------
// misc exports
// ----------------------------------------------------------------------------
extern "C" const char* pkm_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t pkm_abi_version(void) { return PKM_ABI_VERSION; }
extern "C" int32_t pkm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int32_t pkm_spline_tables(int32_t nfi, int32_t nfo, double* A_inv, int32_t* tap_idx,
                                     double* tap_w) {
  return guarded([&] {
    SplineTables sp;
    build_spline_tables(nfi, nfo, sp);
    if (A_inv) std::memcpy(A_inv, sp.A_inv.data(), sp.A_inv.size() * sizeof(double));
    if (tap_idx) std::memcpy(tap_idx, sp.tap_idx.data(), sp.tap_idx.size() * sizeof(int32_t));
    if (tap_w) std::memcpy(tap_w, sp.tap_w.data(), sp.tap_w.size() * sizeof(double));
  });
}

extern "C" int32_t pkm_fold_tables(int32_t nv, int32_t v0_idx, double dv, double* CR, double* CI,
                                   int32_t* idx_plus, int32_t* idx_minus) {
  return guarded([&] {
    PKM_REQUIRE(nv >= 6, "nv=%d too small for the cubic spline (needs nv >= 6)", nv);
    SplineTables sp;
    build_spline_tables(nv / 2 + 1, 2, sp);
    FoldTables f;
    build_fold_tables(nv, v0_idx, dv, sp, f);
    if (CR) std::memcpy(CR, f.CR.data(), size_t(f.nfi) * f.ne * sizeof(double));
    if (CI) std::memcpy(CI, f.CI.data(), size_t(f.nfi) * f.no * sizeof(double));
    if (idx_plus) std::memcpy(idx_plus, f.idx_plus.data(), f.ne * sizeof(int32_t));
    if (idx_minus) std::memcpy(idx_minus, f.idx_minus.data(), f.ne * sizeof(int32_t));
  });
}

// ----------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------
static void collect_timers(pkm_plan* pl);

static void plan_free(pkm_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->d.device);
  auto fr = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  fr(pl->d_CRt); fr(pl->d_CIt); fr(pl->d_CIt_oct); fr(pl->d_CR); fr(pl->d_CI); fr(pl->d_Sw); fr(pl->d_Sw32);
  fr(pl->d_wseg); fr(pl->d_wseg_w); fr(pl->d_SA); fr(pl->d_SA32); fr(pl->d_omega);
  fr(pl->d_chi); fr(pl->d_alpha); fr(pl->d_twiddle); fr(pl->d_Bhat); fr(pl->d_Bhat_h); fr(pl->d_pre); fr(pl->d_post); fr(pl->d_htw); fr(pl->d_wn);
  fr(pl->d_log_dt); fr(pl->d_log_dz); fr(pl->d_exp_tab);
  collect_timers(pl);
  for (int i = 0; i < 2; ++i)
    for (cudaEvent_t e : {pl->ev_in[i], pl->ev_done[i], pl->ev_out[i], pl->ev_ready[i]})
      if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (pl->s_k2[i]) cudaStreamDestroy(pl->s_k2[i]);
    if (pl->ev_k2_fork[i]) cudaEventDestroy(pl->ev_k2_fork[i]);
    if (pl->ev_k2_join[i]) cudaEventDestroy(pl->ev_k2_join[i]);
  }
  if (pl->s_lane) cudaStreamDestroy(pl->s_lane);
  if (pl->ev_lane_fork) cudaEventDestroy(pl->ev_lane_fork);
  if (pl->ev_lane_join) cudaEventDestroy(pl->ev_lane_join);
  if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
  if (pl->s_comp) cudaStreamDestroy(pl->s_comp);
  pl->coef.release(); pl->yhat.release(); pl->ypx.release(); pl->misc.release(); pl->small.release();
  pl->coef1.release(); pl->yhat1.release(); pl->ypx1.release();
  for (int i = 0; i < 2; ++i) {
    pl->stage_in[i].release();
    pl->stage_out[i].release();
    pl->cube_stage[i].release();
  }
  delete pl;
}

extern "C" int32_t pkm_plan_create(const pkm_cube_desc* desc, const double* FXw,
                                   const double* delta_t, const double* delta_z,
                                   pkm_plan** plan_out) {
  pkm_plan* pl = nullptr;
  int32_t rc = guarded([&] {
    PKM_REQUIRE(desc && FXw && delta_t && delta_z && plan_out, "null argument");
    PKM_REQUIRE(desc->nt >= 1 && desc->nz >= 1, "bad SSP grid %d x %d", desc->nt, desc->nz);
    PKM_REQUIRE(desc->n_fft >= 2, "n_fft=%d too small", desc->n_fft);
    PKM_REQUIRE(desc->nv >= 1, "nv=%d", desc->nv);
    PKM_REQUIRE(desc->precision == 0 || desc->precision == 1, "precision=%d: must be 0 (float64) or 1 (float32)", desc->precision);
    int ndev = pkm_device_count();
    if (ndev <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(desc->device >= 0 && desc->device < ndev, "device %d out of range (%d visible)", desc->device, ndev);
    use_device(desc->device);
    pl = new pkm_plan();
    pl->d = *desc;
    pl->nfo = desc->n_fft / 2 + 1;
    pl->nfi = desc->nv / 2 + 1;
    pl->ntz = desc->nt * desc->nz;
    const int nt = desc->nt, nz = desc->nz, nfo = pl->nfo;
    cudaDeviceProp prop;
    PKM_CUDA(cudaGetDeviceProperties(&prop, desc->device));
    pl->num_sms = prop.multiProcessorCount;
    pl->h_delta_t.assign(delta_t, delta_t + nt);
    pl->h_delta_z.assign(delta_z, delta_z + nz);

    {
      std::vector<double> tab(PKM_EXP_TAB);
      pkm_fill_exp_tab(tab.data());
      pl->d_exp_tab = upload(tab);
    }

    // ---- density path tables
    pl->density_ok = desc->v0_idx >= 0 && desc->v0_idx < desc->nv && pl->nfi >= 4;
    if (pl->density_ok) {
      SplineTables sp;
      build_spline_tables(pl->nfi, nfo, sp);
      build_fold_tables(desc->nv, desc->v0_idx, desc->dv, sp, pl->fold);
      const int kt = desc->precision == 1 ? PKM_KT32 : PKM_KT;
      build_wseg_tables(sp, kt, pl->wsegs);
      density_tables_upload(pl);
      // S-hat' per warp segment: [nwseg][nchunk*TZC][kt], rows beyond ntz and slots beyond the
      // segment's count are zero, so every staged chunk is one full-size contiguous bulk copy
      pl->nchunk = (pl->ntz + PKM_TZC - 1) / PKM_TZC;
      const size_t rows = size_t(pl->nchunk) * PKM_TZC;
      std::vector<double> Sw(size_t(2) * pl->wsegs.nwseg * rows * kt, 0.0);
      for (int ws = 0; ws < pl->wsegs.nwseg; ++ws) {
        const int k0 = pl->wsegs.seg[4 * ws + 1], cnt = pl->wsegs.seg[4 * ws + 2];
        for (int t = 0; t < nt; ++t)
          for (int z = 0; z < nz; ++z) {
            const double w = delta_z[z] * delta_t[t];
            double* row = Sw.data() + size_t(2) * ((size_t(ws) * rows + size_t(t) * nz + z) * kt);
            for (int r = 0; r < cnt; ++r) {
              const double* sv = FXw + size_t(2) * ((size_t(k0 + r) * nz + z) * nt + t);
              row[2 * r] = sv[0] * w;
              row[2 * r + 1] = sv[1] * w;
            }
          }
      }
      pl->d_Sw = reinterpret_cast<double2*>(upload(Sw));
      if (desc->precision == 1) {
        std::vector<float> Sw32(Sw.begin(), Sw.end());
        pl->d_Sw32 = reinterpret_cast<float2*>(upload(Sw32));
      }
      pl->d_wseg = upload(pl->wsegs.seg);
      pl->d_wseg_w = upload(pl->wsegs.w);
    }

    // ---- gaussian path: FXw in tensor-core fragment order, omega buffer
    gaussian_tables_upload(pl, FXw);
    {
      std::vector<double> om(nfo, 0.0);
      pl->d_omega = upload(om);
    }

    // ---- inverse FFT tables
    build_czt_tables(desc->n_fft, pl->czt);
    pl->d_chi = reinterpret_cast<double2*>(upload(pl->czt.chi));
    pl->d_alpha = upload(pl->czt.alpha);
    pl->d_twiddle = reinterpret_cast<double2*>(upload(pl->czt.twiddle));
    PKM_CUDA(cudaMalloc(&pl->d_Bhat, sizeof(double2) * pl->czt.M));
    if (!pl->czt.pre.empty()) {
      pl->d_pre = reinterpret_cast<double2*>(upload(pl->czt.pre));
      pl->d_post = reinterpret_cast<double2*>(upload(pl->czt.post));
      pl->d_htw = reinterpret_cast<double2*>(upload(pl->czt.htw));
      PKM_CUDA(cudaMalloc(&pl->d_Bhat_h, sizeof(double2) * pl->czt.M));
    }
    czt_prepare_filter(pl);
    {
      const long double PI_L = 3.14159265358979323846264338327950288419716939937510L;
      std::vector<double> wn(size_t(2) * desc->n_fft);
      for (int r = 0; r < desc->n_fft; ++r) {
        long double ang = 2.0L * PI_L * (long double)r / (long double)desc->n_fft;
        wn[2 * r] = (double)cosl(ang);
        wn[2 * r + 1] = (double)sinl(ang);
      }
      pl->d_wn = reinterpret_cast<double2*>(upload(wn));
    }

    // ---- statistics
    {
      std::vector<double> ldt(nt), ldz(nz);
      for (int t = 0; t < nt; ++t) ldt[t] = std::log(delta_t[t]);
      for (int z = 0; z < nz; ++z) ldz[z] = std::log(delta_z[z]);
      pl->d_log_dt = upload(ldt);
      pl->d_log_dz = upload(ldz);
    }
    PKM_CUDA(cudaDeviceSynchronize());
    *plan_out = pl;
  });
  if (rc != PKM_OK) {
    plan_free(pl);
    if (plan_out) *plan_out = nullptr;
  }
  return rc;
}

extern "C" int32_t pkm_plan_destroy(pkm_plan* plan) {
  return guarded([&] { plan_free(plan); });
}

extern "C" int32_t pkm_plan_set_workspace_limit(pkm_plan* plan, int64_t bytes) {
  return guarded([&] {
    PKM_REQUIRE(plan, "null plan");
    PKM_REQUIRE(bytes >= 0, "negative workspace limit");
    plan->ws_limit = bytes == 0 ? (size_t(8) << 30) : size_t(bytes);
  });
}

extern "C" int64_t pkm_plan_launch_count(const pkm_plan* plan) { return plan ? plan->launches : 0; }

static void collect_timers(pkm_plan* pl) {
  for (TimedLaunch& t : pl->timed) {
    float ms = 0.f;
    if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
      pl->k_ms[t.cls] += ms;
      pl->k_count[t.cls] += 1;
    }
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  pl->timed.clear();
}

extern "C" int32_t pkm_plan_set_profiling(pkm_plan* plan, int32_t enable) {
  return guarded([&] {
    PKM_REQUIRE(plan, "null plan");
    use_device(plan->d.device);
    collect_timers(plan);
    plan->profiling = enable != 0;
    for (int i = 0; i < PKM_K_NCLASS; ++i) {
      plan->k_ms[i] = 0.0;
      plan->k_count[i] = 0;
    }
  });
}

extern "C" int32_t pkm_plan_kernel_times(pkm_plan* plan, double* ms_out, int64_t* count_out) {
  return guarded([&] {
    PKM_REQUIRE(plan && ms_out && count_out, "null argument");
    use_device(plan->d.device);
    collect_timers(plan);
    for (int i = 0; i < PKM_K_NCLASS; ++i) {
      ms_out[i] = plan->k_ms[i];
      count_out[i] = plan->k_count[i];
    }
  });
}

// ----------------------------------------------------------------------------
// shared tail of the two ybar paths: Yhat tile -> irfft -> output layout
// ----------------------------------------------------------------------------
namespace {

struct OutSpec {
  double* out;       // caller buffer (host or device)
  bool host;
  int layout;
  int64_t npix_out;  // pixels of the whole call
  int n;
};

int pixel_tile(const pkm_plan* pl, int64_t npix, size_t bytes_per_px) {
  int64_t t = (int64_t)(pl->ws_limit / std::max<size_t>(bytes_per_px, 1));
  t = std::min<int64_t>(t, npix + 31);
  t = std::min<int64_t>(t, 65535LL * 8);
  t &= ~int64_t(31);
  return (int)std::max<int64_t>(t, 32);
}

// ypx tile [np][n] (device) -> caller's buffer
void emit_tile(pkm_plan* pl, const OutSpec& o, const double* ypx, int64_t pix_off, int np,
               int slot, cudaStream_t st) {
  const size_t n = o.n;
  if (o.layout == PKM_LAYOUT_PIXEL_MAJOR) {
    double* dst = o.out + (size_t)pix_off * n;
    PKM_CUDA(cudaMemcpyAsync(dst, ypx, sizeof(double) * n * np,
                             o.host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    return;
  }
  if (!o.host) {
    {
      KernelTimer tm(pl, PKM_K_TRANSPOSE, st);
      launch_transpose(ypx, np, n, o.out, o.npix_out, pix_off, st);
    }
    pl->launches++;
    return;
  }
  pl->stage_out[slot].reserve(sizeof(double) * n * np);
  double* tmp = pl->stage_out[slot].as<double>();
  {
    KernelTimer tm(pl, PKM_K_TRANSPOSE, st);
    launch_transpose(ypx, np, n, tmp, np, 0, st);
  }
  pl->launches++;
  PKM_CUDA(cudaMemcpy2DAsync(o.out + pix_off, sizeof(double) * o.npix_out, tmp, sizeof(double) * np,
                             sizeof(double) * np, n, cudaMemcpyDeviceToHost, st));
}

}  // namespace

// tile lanes of the device-resident density path: PKM_TILE_LANES=1 serialises the tiles, 2 forces
// two lanes, unset (-1) = two lanes when the tiles are small
static int tile_lanes() {
  const char* e = getenv("PKM_TILE_LANES");
  return e ? atoi(e) : -1;
}

// pixel tiles of a range evaluated into an embedded (peer-memory) cube
static int peer_tiles() {
  const char* e = getenv("PKM_PEER_TILES");
  return e ? std::max(1, atoi(e)) : 8;
}

// streams / events of the host-buffer pipeline (created on first use)
static void pipe_init(pkm_plan* pl) {
  if (pl->s_copy) return;
  PKM_CUDA(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
  PKM_CUDA(cudaStreamCreateWithFlags(&pl->s_comp, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_in[i], cudaEventDisableTiming));
    PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
    PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_out[i], cudaEventDisableTiming));
    PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_ready[i], cudaEventDisableTiming));
  }
}

// ----------------------------------------------------------------------------
// density path
// ----------------------------------------------------------------------------
// cube_npix > 0: cube layout written into a larger cube [n_fft][cube_npix] at pixel offset
// cube_pix0 (local or peer device memory); otherwise the output covers exactly the rows asked for.
static int32_t density_body(pkm_plan* pl, const double* dens, int32_t is_log, int32_t nx1,
                            int32_t nx2, int32_t row_begin, int32_t row_end, double* ybar_out,
                            int32_t out_layout, int64_t cube_npix, int64_t cube_pix0, void* stream);

static int32_t density_impl(pkm_plan* pl, const double* dens, int32_t is_log, int32_t nx1,
                            int32_t nx2, int32_t row_begin, int32_t row_end, double* ybar_out,
                            int32_t out_layout, int64_t cube_npix, int64_t cube_pix0, void* stream) {
  const int32_t rc = density_body(pl, dens, is_log, nx1, nx2, row_begin, row_end, ybar_out, out_layout, cube_npix,
                                  cube_pix0, stream);
  if (rc != PKM_OK && pl) {
    // a failure in the middle of the tile loop may leave asynchronous copies on the caller's host
    // buffers in flight: drain the plan's streams before handing the buffers back
    if (pl->s_copy) cudaStreamSynchronize(pl->s_copy);
    if (pl->s_comp) cudaStreamSynchronize(pl->s_comp);
    for (int i = 0; i < 2; ++i)
      if (pl->s_k2[i]) cudaStreamSynchronize(pl->s_k2[i]);
    if (pl->s_lane) cudaStreamSynchronize(pl->s_lane);
    cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
    cudaGetLastError();
  }
  return rc;
}

static int32_t density_body(pkm_plan* pl, const double* dens, int32_t is_log, int32_t nx1,
                            int32_t nx2, int32_t row_begin, int32_t row_end, double* ybar_out,
                            int32_t out_layout, int64_t cube_npix, int64_t cube_pix0, void* stream) {
  return guarded([&] {
    PkmRange range_call("pkm_ybar_density");
    PKM_REQUIRE(pl && dens && ybar_out, "null argument");
    PKM_REQUIRE(nx1 >= 1 && nx2 >= 1 && row_begin >= 0 && row_end <= nx1 && row_begin <= row_end,
                "bad pixel range rows [%d,%d) of %d x %d", row_begin, row_end, nx1, nx2);
    PKM_REQUIRE(out_layout == PKM_LAYOUT_CUBE || out_layout == PKM_LAYOUT_PIXEL_MAJOR, "bad out_layout %d", out_layout);
    if (pl->d.v0_idx < 0)
      PKM_THROW(PKM_ERR_NO_ZERO_BIN, "no velocity bin centred on zero (nv=%d): the density path is undefined", pl->d.nv);
    PKM_REQUIRE(pl->density_ok, "density path unavailable for nv=%d (the cubic spline needs nv >= 6)", pl->d.nv);
    use_device(pl->d.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool in_host = !pkm_is_device_pointer(dens);
    const bool out_host = !pkm_is_device_pointer(ybar_out);
    const int nt = pl->d.nt, nv = pl->d.nv, nz = pl->d.nz, n = pl->d.n_fft;
    const int64_t npix_total = (int64_t)nx1 * nx2;
    const int64_t pix_begin = (int64_t)row_begin * nx2;
    const int64_t npix = (int64_t)(row_end - row_begin) * nx2;
    if (npix == 0) return;

    // Host buffers are pipelined: H2D of tile i+1 (copy stream) and D2H of tile i-1 (output
    // stream) overlap the kernels of tile i (caller's stream); two staging slots each way.
    size_t per_px = sizeof(double2) * (size_t)pl->ntz * pl->nfi + sizeof(double2) * pl->nfo + sizeof(double) * n;
    if (in_host) per_px += 2 * sizeof(double) * (size_t)nt * nv * nz;
    if (out_host) per_px += 2 * sizeof(double) * n;
    // Pixel tiles.  Upper bound: the workspace; host buffers: at least 8 tiles so that copies and
    // kernels overlap.  Every tile is sized to a WHOLE number of waves of the contraction kernel
    // (its CTAs are long: a last, nearly empty wave costs as much as a full one).  Embedded
    // (possibly peer-memory) cube: the transposes and the NVLink copies run on a side stream behind
    // the next tile's kernels, so the range is cut into four tiles.
    const bool side_transpose = cube_npix > 0;
    const int max_pg = std::max(1, pixel_tile(pl, npix, per_px) / 32);
    const int64_t total_pg = (npix + 31) / 32;
    const double wave = contract_wave_pixel_groups(pl);
    // largest whole number of waves within pg pixel groups (a contraction CTA covers 4 groups, so
    // tiles are whole CTA rows), at least one group
    auto whole_waves = [&](double pg) -> int {
      const double m = std::floor(pg / wave);
      int r = m >= 1.0 ? ((int)std::floor(m * wave) / 4) * 4 : (int)pg;
      if (r < 1) r = (int)pg;
      return std::max(1, std::min(r, max_pg));
    };
    std::vector<int> sched;                           // pixels per tile
    {
      int64_t rem = total_pg;
      if (side_transpose) {
        // eight near-equal tiles (whole CTA rows) alternating between the two lanes: measured best on
        // 8 GPUs (4 tiles 9.78, 6 tiles 9.57, 8 tiles 9.55 ms per cube); a geometric schedule with a
        // small last tile (less exposed copy) lost more to half-empty contraction launches
        const int nt_peer = peer_tiles();
        const int t = (int)std::min<int64_t>(max_pg, (((total_pg + nt_peer - 1) / nt_peer + 3) / 4) * 4);
        while (rem > 0) {
          sched.push_back((int)std::min<int64_t>(rem, t));
          rem -= sched.back();
        }
      } else {
        const double want = (in_host || out_host) ? std::min<double>(max_pg, std::ceil(total_pg / 8.0)) : max_pg;
        const int t = whole_waves(want);
        while (rem > 0) {
          sched.push_back((int)std::min<int64_t>(rem, t));
          rem -= sched.back();
        }
      }
      int64_t left = npix;
      for (int& t : sched) {                          // pixel groups -> pixels (the last tile may be ragged)
        t = (int)std::min<int64_t>((int64_t)t * 32, left);
        left -= t;
      }
    }
    int tile = 32;                                    // the largest tile, padded to whole groups of 32
    for (int t : sched) tile = std::max(tile, (t + 31) & ~31);

    pl->coef.reserve(sizeof(double2) * (size_t)pl->ntz * pl->nfi * tile);
    pl->yhat.reserve(sizeof(double2) * (size_t)pl->nfo * tile);
    pl->ypx.reserve(sizeof(double) * (size_t)n * tile);
    if (in_host || out_host || side_transpose) pipe_init(pl);
    // Device-resident input and output: the tiles alternate between two lanes (own streams and
    // workspaces), so that the kernels of one tile fill the wave tails - and the latency-bound
    // inverse FFT - of the other.
    // Worth it when the tiles are small (a few waves per kernel: 8 GPUs sharing one cube, small
    // workspaces); large tiles run in order, which also keeps per-kernel event times meaningful.
    const int lanes_env = tile_lanes();
    const bool lanes_auto = side_transpose || tile <= 4096;
    const int nlanes = (!in_host && !out_host && sched.size() > 1 && (lanes_env == 2 || (lanes_env < 0 && lanes_auto))) ? 2 : 1;
    if (nlanes == 2) {
      pl->coef1.reserve(sizeof(double2) * (size_t)pl->ntz * pl->nfi * tile);
      pl->yhat1.reserve(sizeof(double2) * (size_t)pl->nfo * tile);
      if (!side_transpose) pl->ypx1.reserve(sizeof(double) * (size_t)n * tile);
      if (!pl->s_lane) {
        PKM_CUDA(cudaStreamCreateWithFlags(&pl->s_lane, cudaStreamNonBlocking));
        PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_lane_fork, cudaEventDisableTiming));
        PKM_CUDA(cudaEventCreateWithFlags(&pl->ev_lane_join, cudaEventDisableTiming));
      }
      PKM_CUDA(cudaEventRecord(pl->ev_lane_fork, reinterpret_cast<cudaStream_t>(stream)));
      PKM_CUDA(cudaStreamWaitEvent(pl->s_lane, pl->ev_lane_fork, 0));
    }
    cudaStream_t const st0 = reinterpret_cast<cudaStream_t>(stream);
    cudaStream_t s_in = pl->s_copy, s_out = pl->s_comp;
    const bool pm = out_layout == PKM_LAYOUT_PIXEL_MAJOR;
    OutSpec o{ybar_out, out_host, out_layout, npix, n};
    if (cube_npix > 0) {
      PKM_REQUIRE(!out_host && !pm && cube_pix0 >= 0 && cube_pix0 + npix <= cube_npix,
                  "embedded cube output needs a device buffer, cube layout and a valid offset");
      o.npix_out = cube_npix;
      o.out = ybar_out + cube_pix0;
    }

    int64_t it = 0;
    for (int64_t p0 = 0; p0 < npix; p0 += sched[it], ++it) {
      const int np = sched[it];
      const int slot = (int)(it & 1);
      const int lane = nlanes == 2 ? slot : 0;
      cudaStream_t st = lane ? pl->s_lane : st0;
      DevBuf& coef = lane ? pl->coef1 : pl->coef;
      DevBuf& yhat = lane ? pl->yhat1 : pl->yhat;
      DevBuf& ypxb = lane ? pl->ypx1 : pl->ypx;
      PkmRange range_tile("pkm density tile: coeff + contract + irfft");
      const double* src = dens;
      int64_t src_npix = npix_total, src_pix0 = pix_begin + p0;
      if (in_host) {
        pl->stage_in[slot].reserve(sizeof(double) * (size_t)nt * nv * nz * tile);
        PKM_CUDA(cudaStreamWaitEvent(s_in, pl->ev_done[slot], 0));   // kernel 1 of tile it-2 is done with it
        PKM_CUDA(cudaMemcpy2DAsync(pl->stage_in[slot].p, sizeof(double) * (size_t)np * nz,
                                   dens + (size_t)(pix_begin + p0) * nz,
                                   sizeof(double) * (size_t)npix_total * nz,
                                   sizeof(double) * (size_t)np * nz, (size_t)nt * nv,
                                   cudaMemcpyHostToDevice, s_in));
        PKM_CUDA(cudaEventRecord(pl->ev_in[slot], s_in));
        PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_in[slot], 0));
        src = pl->stage_in[slot].as<double>();
        src_npix = np;
        src_pix0 = 0;
      }
      launch_coeff(pl, src, is_log, src_npix, src_pix0, np, tile, coef.as<double2>(), st);
      if (in_host) PKM_CUDA(cudaEventRecord(pl->ev_done[slot], st));
      launch_contract(pl, coef.as<double2>(), np, tile, yhat.as<double2>(), st, lane);
      if (side_transpose) {
        // two pixel-major buffers (stage_out doubles as the second one)
        pl->stage_out[slot].reserve(sizeof(double) * (size_t)n * tile);
        double* ypx = pl->stage_out[slot].as<double>();
        PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_out[slot], 0));      // transpose of tile it-2 has drained it
        launch_irfft_czt(pl, yhat.as<double2>(), np, pl->d.dx * pl->d.dy, ypx, st);
        PKM_CUDA(cudaEventRecord(pl->ev_ready[slot], st));
        PKM_CUDA(cudaStreamWaitEvent(s_out, pl->ev_ready[slot], 0));
        // transpose locally, then let a copy engine move the [n][np] block into the (possibly
        // remote) cube as one strided 2-D copy: DMA bursts use NVLink far better than the
        // 256-byte stores of a kernel writing to a peer, and cost no SM time
        pl->cube_stage[slot].reserve(sizeof(double) * (size_t)n * tile);
        double* cs = pl->cube_stage[slot].as<double>();
        {
          KernelTimer tm(pl, PKM_K_TRANSPOSE, s_out);
          launch_transpose(ypx, np, n, cs, np, 0, s_out);
        }
        pl->launches++;
        PKM_CUDA(cudaMemcpy2DAsync(o.out + p0, sizeof(double) * (size_t)o.npix_out, cs, sizeof(double) * (size_t)np,
                                   sizeof(double) * (size_t)np, (size_t)n, cudaMemcpyDefault, s_out));
        PKM_CUDA(cudaEventRecord(pl->ev_out[slot], s_out));
        continue;
      }
      if (!out_host) {
        launch_irfft_czt(pl, yhat.as<double2>(), np, pl->d.dx * pl->d.dy, ypxb.as<double>(), st);
        emit_tile(pl, o, ypxb.as<double>(), p0, np, 0, st);
        continue;
      }
      // host output: the staging slot must have been drained by the D2H of tile it-2
      pl->stage_out[slot].reserve(sizeof(double) * (size_t)n * tile);
      double* stage = pl->stage_out[slot].as<double>();
      PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_out[slot], 0));
      if (pm) {
        launch_irfft_czt(pl, yhat.as<double2>(), np, pl->d.dx * pl->d.dy, stage, st);
      } else {
        launch_irfft_czt(pl, yhat.as<double2>(), np, pl->d.dx * pl->d.dy, ypxb.as<double>(), st);
        KernelTimer tm(pl, PKM_K_TRANSPOSE, st);
        launch_transpose(ypxb.as<double>(), np, n, stage, np, 0, st);
        pl->launches++;
      }
      PKM_CUDA(cudaEventRecord(pl->ev_ready[slot], st));
      PKM_CUDA(cudaStreamWaitEvent(s_out, pl->ev_ready[slot], 0));
      if (pm)
        PKM_CUDA(cudaMemcpyAsync(ybar_out + (size_t)p0 * n, stage, sizeof(double) * (size_t)n * np,
                                 cudaMemcpyDeviceToHost, s_out));
      else
        PKM_CUDA(cudaMemcpy2DAsync(ybar_out + p0, sizeof(double) * npix, stage, sizeof(double) * np,
                                   sizeof(double) * np, n, cudaMemcpyDeviceToHost, s_out));
      PKM_CUDA(cudaEventRecord(pl->ev_out[slot], s_out));
    }
    if (nlanes == 2) {
      PKM_CUDA(cudaEventRecord(pl->ev_lane_join, pl->s_lane));
      PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_lane_join, 0));
    }
    if (side_transpose) {
      // stream semantics for the caller: everything is complete once `st` reaches this point
      PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_out[0], 0));
      PKM_CUDA(cudaStreamWaitEvent(st, pl->ev_out[1], 0));
    }
    if (in_host || out_host) {
      PKM_CUDA(cudaStreamSynchronize(st));
      PKM_CUDA(cudaStreamSynchronize(s_out));
      PKM_CUDA(cudaStreamSynchronize(s_in));
    }
  });
}

extern "C" int32_t pkm_ybar_density(pkm_plan* pl, const double* dens, int32_t is_log, int32_t nx1,
                                    int32_t nx2, int32_t row_begin, int32_t row_end,
                                    double* ybar_out, int32_t out_layout, void* stream) {
  return density_impl(pl, dens, is_log, nx1, nx2, row_begin, row_end, ybar_out, out_layout, 0, 0, stream);
}

extern "C" int32_t pkm_ybar_density_into_cube(pkm_plan* pl, const double* dens, int32_t is_log,
                                              int32_t nx1, int32_t nx2, int32_t row_begin,
                                              int32_t row_end, double* cube, int32_t cube_nx1,
                                              int32_t cube_row0, void* stream) {
  if (cube_nx1 < 1 || cube_row0 < 0) {
    pkm_set_error("bad cube geometry");
    return PKM_ERR_INVALID;
  }
  return density_impl(pl, dens, is_log, nx1, nx2, row_begin, row_end, cube, PKM_LAYOUT_CUBE,
                      (int64_t)cube_nx1 * nx2, (int64_t)cube_row0 * nx2, stream);
}

extern "C" int32_t pkm_host_register(void* ptr, int64_t bytes) {
  return guarded([&] {
    PKM_REQUIRE(ptr && bytes > 0, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible");
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) return;  // already pinned
    cudaGetLastError();
    PKM_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  });
}

extern "C" int32_t pkm_host_unregister(void* ptr) {
  return guarded([&] {
    PKM_REQUIRE(ptr, "null argument");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();   // not registered (any more): nothing to undo
  });
}

extern "C" int32_t pkm_device_alloc(int32_t device, int64_t bytes, void** ptr_out) {
  return guarded([&] {
    PKM_REQUIRE(ptr_out && bytes > 0, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible");
    use_device(device);
    PKM_CUDA(cudaMalloc(ptr_out, (size_t)bytes));
  });
}

extern "C" int32_t pkm_device_free(void* ptr) {
  return guarded([&] { PKM_CUDA(cudaFree(ptr)); });
}

extern "C" int32_t pkm_ipc_export(const void* dev_ptr, uint8_t* handle64) {
  return guarded([&] {
    PKM_REQUIRE(dev_ptr && handle64, "null argument");
    cudaIpcMemHandle_t h;
    PKM_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
  });
}

extern "C" int32_t pkm_ipc_open(const uint8_t* handle64, int32_t device, void** ptr_out) {
  return guarded([&] {
    PKM_REQUIRE(handle64 && ptr_out, "null argument");
    use_device(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PKM_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  });
}

extern "C" int32_t pkm_ipc_close(void* ptr) {
  return guarded([&] { PKM_CUDA(cudaIpcCloseMemHandle(ptr)); });
}

// ----------------------------------------------------------------------------
// Gaussian-LOSVD path
// ----------------------------------------------------------------------------
extern "C" int32_t pkm_ybar_gaussian(pkm_plan* pl, const double* P_txz, const double* mu_v,
                                     const int64_t mu_strides[3], const double* sig_v,
                                     const int64_t sig_strides[3], const double* omega,
                                     int32_t nx1, int32_t nx2, int32_t row_begin, int32_t row_end,
                                     double* ybar_out, int32_t out_layout, void* stream) {
  return guarded([&] {
    PkmRange range_call("pkm_ybar_gaussian");
    PKM_REQUIRE(pl && P_txz && mu_v && sig_v && omega && ybar_out && mu_strides && sig_strides, "null argument");
    PKM_REQUIRE(nx1 >= 1 && nx2 >= 1 && row_begin >= 0 && row_end <= nx1 && row_begin <= row_end,
                "bad pixel range rows [%d,%d) of %d x %d", row_begin, row_end, nx1, nx2);
    PKM_REQUIRE(out_layout == PKM_LAYOUT_CUBE || out_layout == PKM_LAYOUT_PIXEL_MAJOR, "bad out_layout %d", out_layout);
    use_device(pl->d.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool in_host = !pkm_is_device_pointer(P_txz);
    PKM_REQUIRE(in_host == !pkm_is_device_pointer(mu_v) && in_host == !pkm_is_device_pointer(sig_v),
                "P_txz, mu_v and sig_v must all be host or all be device buffers");
    const bool out_host = !pkm_is_device_pointer(ybar_out);
    const int nt = pl->d.nt, nz = pl->d.nz, n = pl->d.n_fft;
    const int64_t npix_total = (int64_t)nx1 * nx2;
    const int64_t pix_begin = (int64_t)row_begin * nx2;
    const int64_t npix = (int64_t)(row_end - row_begin) * nx2;
    if (npix == 0) return;

    // omega is a plan constant in practice (linspace(0, pi, nfo)/dv): it is uploaded when it first
    // appears or changes, from a plan-owned copy, without blocking the stream or the host
    if (pl->h_omega.size() != (size_t)pl->nfo || memcmp(pl->h_omega.data(), omega, sizeof(double) * pl->nfo) != 0) {
      pl->h_omega.assign(omega, omega + pl->nfo);
      PKM_CUDA(cudaMemcpyAsync(pl->d_omega, pl->h_omega.data(), sizeof(double) * pl->nfo, cudaMemcpyHostToDevice, st));
    }
    // frequency stride of one lane of the kernel: omega[4] - omega[0]
    const double omega_step4 = pl->nfo >= 2 ? 4.0 * (omega[1] - omega[0]) : 0.0;

    size_t per_px = sizeof(double2) * pl->nfo + sizeof(double) * n;
    if (in_host) per_px += sizeof(double) * (size_t)nt * (nz + 2);
    if (out_host && out_layout == PKM_LAYOUT_CUBE) per_px += sizeof(double) * n;
    const int tile = pixel_tile(pl, npix, per_px);
    pl->yhat.reserve(sizeof(double2) * (size_t)pl->nfo * tile);
    pl->ypx.reserve(sizeof(double) * (size_t)n * tile);
    OutSpec o{ybar_out, out_host, out_layout, npix, n};
    std::vector<double> hmu, hsg;

    for (int64_t p0 = 0; p0 < npix; p0 += tile) {
      const int np = (int)std::min<int64_t>(tile, npix - p0);
      const double *dP = P_txz, *dmu = mu_v, *dsg = sig_v;
      int64_t ms[3] = {mu_strides[0], mu_strides[1], mu_strides[2]};
      int64_t ss[3] = {sig_strides[0], sig_strides[1], sig_strides[2]};
      int64_t k_nx2 = nx2, k_npix_total = npix_total, k_pix0 = pix_begin + p0;
      if (in_host) {
        const size_t nP = (size_t)nt * np * nz, nM = (size_t)nt * np;
        pl->stage_in[0].reserve(sizeof(double) * (nP + 2 * nM));
        double* base = pl->stage_in[0].as<double>();
        PKM_CUDA(cudaMemcpy2DAsync(base, sizeof(double) * (size_t)np * nz,
                                   P_txz + (size_t)(pix_begin + p0) * nz,
                                   sizeof(double) * (size_t)npix_total * nz,
                                   sizeof(double) * (size_t)np * nz, (size_t)nt,
                                   cudaMemcpyHostToDevice, st));
        hmu.resize(nM);
        hsg.resize(nM);
        for (int t = 0; t < nt; ++t)
          for (int i = 0; i < np; ++i) {
            const int64_t xg = pix_begin + p0 + i, x1 = xg / nx2, x2 = xg % nx2;
            hmu[(size_t)t * np + i] = mu_v[t * ms[0] + x1 * ms[1] + x2 * ms[2]];
            hsg[(size_t)t * np + i] = sig_v[t * ss[0] + x1 * ss[1] + x2 * ss[2]];
          }
        PKM_CUDA(cudaMemcpyAsync(base + nP, hmu.data(), sizeof(double) * nM, cudaMemcpyHostToDevice, st));
        PKM_CUDA(cudaMemcpyAsync(base + nP + nM, hsg.data(), sizeof(double) * nM, cudaMemcpyHostToDevice, st));
        PKM_CUDA(cudaStreamSynchronize(st));  // hmu / hsg are reused by the next tile
        dP = base;
        dmu = base + nP;
        dsg = base + nP + nM;
        ms[0] = ss[0] = np; ms[1] = ss[1] = 1; ms[2] = ss[2] = 0;
        k_nx2 = 1; k_npix_total = np; k_pix0 = 0;
      }
      launch_gaussian(pl, dP, dmu, ms, dsg, ss, k_nx2, k_npix_total, k_pix0, np, tile, omega_step4,
                      pl->yhat.as<double2>(), st);
      launch_irfft_czt(pl, pl->yhat.as<double2>(), np, 1.0, pl->ypx.as<double>(), st);
      emit_tile(pl, o, pl->ypx.as<double>(), p0, np, 0, st);
    }
    if (in_host || out_host) PKM_CUDA(cudaStreamSynchronize(st));
  });
}

// ----------------------------------------------------------------------------
// small operators
// ----------------------------------------------------------------------------
// mode 0: out = sum_c w[c] * cube_c;  mode 1: out = log sum_c exp(w[c] + cube_c)
static int32_t combine_cubes(int mode, int32_t n_cmps, const double* const* ybars, const double* weights,
                             int64_t n, double* out, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(n_cmps >= 1 && ybars && weights && out && n >= 0, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool host = !pkm_is_device_pointer(out);
    for (int c = 0; c < n_cmps; ++c)
      PKM_REQUIRE(host == !pkm_is_device_pointer(ybars[c]), "all cubes must be on the same side as out");
    DevBuf tab, data;
    std::vector<const double*> ptrs(ybars, ybars + n_cmps);
    double* d_out = out;
    if (host) {
      data.reserve(sizeof(double) * (size_t)n * (n_cmps + 1));
      double* base = data.as<double>();
      for (int c = 0; c < n_cmps; ++c) {
        PKM_CUDA(cudaMemcpyAsync(base + (size_t)c * n, ybars[c], sizeof(double) * n, cudaMemcpyHostToDevice, st));
        ptrs[c] = base + (size_t)c * n;
      }
      d_out = base + (size_t)n_cmps * n;
    }
    tab.reserve(sizeof(double*) * n_cmps + sizeof(double) * n_cmps);
    const double** d_ptrs = tab.as<const double*>();
    double* d_w = reinterpret_cast<double*>(d_ptrs + n_cmps);
    PKM_CUDA(cudaMemcpyAsync(d_ptrs, ptrs.data(), sizeof(double*) * n_cmps, cudaMemcpyHostToDevice, st));
    PKM_CUDA(cudaMemcpyAsync(d_w, weights, sizeof(double) * n_cmps, cudaMemcpyHostToDevice, st));
    if (mode == 0) launch_axpy(n_cmps, d_ptrs, d_w, n, d_out, st);
    else launch_logsumexp_cubes(n_cmps, d_ptrs, d_w, n, d_out, st);
    if (host) PKM_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PKM_CUDA(cudaStreamSynchronize(st));  // the pointer table is a local
    tab.release();
    data.release();
  });
}

extern "C" int32_t pkm_axpy_cubes(int32_t n_cmps, const double* const* ybars, const double* weights,
                                  int64_t n, double* out, int32_t device, void* stream) {
  return combine_cubes(0, n_cmps, ybars, weights, n, out, device, stream);
}

extern "C" int32_t pkm_logsumexp_cubes(int32_t n_cmps, const double* const* log_ps, const double* log_weights,
                                       int64_t n, double* out, int32_t device, void* stream) {
  return combine_cubes(1, n_cmps, log_ps, log_weights, n, out, device, stream);
}

extern "C" int32_t pkm_transpose_to_cube(const double* pixel_major, int64_t npix, int64_t n,
                                         double* cube, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(pixel_major && cube, "null argument");
    PKM_REQUIRE(pkm_is_device_pointer(pixel_major) && pkm_is_device_pointer(cube), "device pointers required");
    use_device(device);
    launch_transpose(pixel_major, npix, n, cube, npix, 0, reinterpret_cast<cudaStream_t>(stream));
  });
}

static void irfft_common(pkm_plan* pl, const double* Yhat, int64_t npix, double scale, double* out,
                         void* stream, bool naive) {
  PKM_REQUIRE(pl && Yhat && out && npix >= 0, "bad argument");
  use_device(pl->d.device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool in_host = !pkm_is_device_pointer(Yhat), out_host = !pkm_is_device_pointer(out);
  const double2* dY = reinterpret_cast<const double2*>(Yhat);
  double* dO = out;
  if (in_host) {
    pl->yhat.reserve(sizeof(double2) * (size_t)pl->nfo * npix);
    PKM_CUDA(cudaMemcpyAsync(pl->yhat.p, Yhat, sizeof(double2) * (size_t)pl->nfo * npix, cudaMemcpyHostToDevice, st));
    dY = pl->yhat.as<double2>();
  }
  if (out_host) {
    pl->ypx.reserve(sizeof(double) * (size_t)pl->d.n_fft * npix);
    dO = pl->ypx.as<double>();
  }
  if (naive) launch_irfft_naive(pl, dY, npix, scale, dO, st);
  else launch_irfft_czt(pl, dY, npix, scale, dO, st);
  if (out_host) PKM_CUDA(cudaMemcpyAsync(out, dO, sizeof(double) * (size_t)pl->d.n_fft * npix, cudaMemcpyDeviceToHost, st));
  if (in_host || out_host) PKM_CUDA(cudaStreamSynchronize(st));
}

extern "C" int32_t pkm_irfft_batch(pkm_plan* pl, const double* Yhat, int64_t npix, double scale,
                                   double* out, void* stream) {
  return guarded([&] { irfft_common(pl, Yhat, npix, scale, out, stream, false); });
}
extern "C" int32_t pkm_irfft_naive(pkm_plan* pl, const double* Yhat, int64_t npix, double scale,
                                   double* out, void* stream) {
  return guarded([&] { irfft_common(pl, Yhat, npix, scale, out, stream, true); });
}

// ----------------------------------------------------------------------------
// statistics
// ----------------------------------------------------------------------------
namespace {

int64_t marginal_size(const pkm_plan* pl, int64_t npix, uint32_t mask) {
  int64_t n = 1;
  if (mask & 1) n *= pl->d.nt;
  if (mask & 2) n *= pl->d.nv;
  if (mask & 4) n *= npix;
  if (mask & 8) n *= pl->d.nz;
  return n;
}

// device pointers to the small constants of a reduction: log dv [nv] and (optional) log light weights
struct ReduceConsts {
  const double* log_dv;
  const double* log_lw;
  int uniform_dv;
};
ReduceConsts upload_reduce_consts(pkm_plan* pl, const double* log_dv, const double* light_weights,
                                  cudaStream_t st) {
  const int nt = pl->d.nt, nv = pl->d.nv, nz = pl->d.nz;
  std::vector<double> h(nv + (size_t)nt * nz);
  for (int v = 0; v < nv; ++v) h[v] = log_dv[v];
  if (light_weights)
    for (int i = 0; i < nt * nz; ++i) h[nv + i] = std::log(light_weights[i]);
  pl->small.reserve(sizeof(double) * h.size());
  PKM_CUDA(cudaMemcpyAsync(pl->small.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
  PKM_CUDA(cudaStreamSynchronize(st));   // h is a local
  // equal bins up to the rounding of diff(v_edg) (a few ulp): the reductions then fold log dv into a
  // per-thread constant instead of adding it to every element (differences <= 1e-15 in log p)
  int uniform = 1;
  for (int v = 1; v < nv; ++v)
    uniform = uniform && std::fabs(log_dv[v] - log_dv[0]) <= 1e-14 * (std::fabs(log_dv[0]) + 1.0);
  return ReduceConsts{pl->small.as<double>(), light_weights ? pl->small.as<double>() + nv : nullptr, uniform};
}

// src (device): the 5-D log density (src_mask = 0xF, weighted) or a log-probability marginal holding
// the axes of src_mask; out (device): the marginal over keep_mask (a subset of src_mask)
void reduce_device(pkm_plan* pl, const double* src, uint32_t src_mask, bool src_is_density, int64_t npix,
                   const ReduceConsts& c, bool light, uint32_t keep_mask, int density, double* out,
                   cudaStream_t st) {
  PKM_REQUIRE((keep_mask & ~src_mask) == 0, "keep_mask 0x%x is not a subset of the source axes 0x%x", keep_mask, src_mask);
  ReduceSrc rs;
  rs.nt = (src_mask & 1) ? pl->d.nt : 1;
  rs.nv = (src_mask & 2) ? pl->d.nv : 1;
  rs.npix = (src_mask & 4) ? npix : 1;
  rs.nz = (src_mask & 8) ? pl->d.nz : 1;
  rs.weighted = src_is_density ? 1 : 0;
  rs.uniform_dv = c.uniform_dv;
  if (light) PKM_REQUIRE((src_mask & 9) == 9, "light weighting needs a source that keeps t and z");
  if (keep_mask == src_mask && !src_is_density && !light) {
    // nothing to reduce: copy, then convert to a density if asked
    const int64_t n = marginal_size(pl, npix, keep_mask);
    if (out != src) PKM_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (density) launch_apply_log_dv(pl, out, rs, c.log_dv, keep_mask, -1.0, st);
    return;
  }
  // the kernels address kept axes with the SOURCE extents: a kept axis always has its full extent
  uint32_t km = keep_mask;
  launch_reduce_logp(pl, src, rs, c.log_dv, light ? c.log_lw : nullptr, km, density, out, st);
}

}  // namespace

extern "C" int32_t pkm_reduce_logp(pkm_plan* pl, const double* log_p_tvxz, int64_t npix,
                                   const double* log_dv, const double* light_weights,
                                   uint32_t keep_mask, int32_t density, double* out, void* stream) {
  return pkm_reduce_logp_multi(pl, log_p_tvxz, npix, log_dv, light_weights, 1, &keep_mask, &density, &out, stream);
}

extern "C" int32_t pkm_reduce_logp_multi(pkm_plan* pl, const double* log_p_tvxz, int64_t npix,
                                         const double* log_dv, const double* light_weights, int32_t n_out,
                                         const uint32_t* keep_masks, const int32_t* density,
                                         double* const* outs, void* stream) {
  return guarded([&] {
    PkmRange range_call("pkm_reduce_logp");
    PKM_REQUIRE(pl && log_p_tvxz && log_dv && keep_masks && density && outs && npix >= 1 && n_out >= 1 && n_out <= 16,
                "bad argument");
    use_device(pl->d.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nt = pl->d.nt, nv = pl->d.nv, nz = pl->d.nz;
    const bool host = !pkm_is_device_pointer(log_p_tvxz);
    for (int i = 0; i < n_out; ++i) {
      PKM_REQUIRE(keep_masks[i] <= 0xF, "bad keep_mask");
      PKM_REQUIRE(outs[i] && host == !pkm_is_device_pointer(outs[i]), "input and outputs must be on the same side");
    }
    const ReduceConsts c = upload_reduce_consts(pl, log_dv, light_weights, st);
    const bool light = light_weights != nullptr;
    const double* d_in = log_p_tvxz;
    if (host) {
      const size_t ntot = (size_t)nt * nv * npix * nz;
      pl->stage_in[0].reserve(sizeof(double) * ntot);
      PKM_CUDA(cudaMemcpyAsync(pl->stage_in[0].p, log_p_tvxz, sizeof(double) * ntot, cudaMemcpyHostToDevice, st));
      d_in = pl->stage_in[0].as<double>();
    }
    // Largest marginals first; every output is held as a log PROBABILITY (density = 0) until the end,
    // so that a later, smaller marginal can be reduced from the smallest already computed superset
    // instead of the 5-D array.  (The reductions are bound by one exponential per source element,
    // so only nested requests share work; disjoint ones each read the 5-D array once.)
    std::vector<int> order(n_out);
    for (int i = 0; i < n_out; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      return marginal_size(pl, npix, keep_masks[a]) > marginal_size(pl, npix, keep_masks[b]);
    });
    std::vector<double*> dev(n_out, nullptr);
    size_t stage_bytes = 0;
    if (host)
      for (int i = 0; i < n_out; ++i) stage_bytes += sizeof(double) * (size_t)marginal_size(pl, npix, keep_masks[i]);
    if (host) pl->stage_out[0].reserve(stage_bytes);
    {
      double* cur = host ? pl->stage_out[0].as<double>() : nullptr;
      for (int i = 0; i < n_out; ++i) {
        dev[i] = host ? cur : outs[i];
        if (host) cur += marginal_size(pl, npix, keep_masks[i]);
      }
    }
    for (int oi = 0; oi < n_out; ++oi) {
      const int i = order[oi];
      int best = -1;
      for (int oj = 0; oj < oi; ++oj) {
        const int j = order[oj];
        if ((keep_masks[i] & ~keep_masks[j]) != 0) continue;
        if (best < 0 || marginal_size(pl, npix, keep_masks[j]) < marginal_size(pl, npix, keep_masks[best])) best = j;
      }
      if (best >= 0)   // already light-weighted (if asked): a plain log-sum-exp of the superset
        reduce_device(pl, dev[best], keep_masks[best], false, npix, c, false, keep_masks[i], 0, dev[i], st);
      else
        reduce_device(pl, d_in, 0xF, true, npix, c, light, keep_masks[i], 0, dev[i], st);
    }
    for (int i = 0; i < n_out; ++i)
      if (density[i]) {
        ReduceSrc rs{nt, nv, nz, npix, 0};
        launch_apply_log_dv(pl, dev[i], rs, c.log_dv, keep_masks[i], -1.0, st);
      }
    if (host) {
      for (int i = 0; i < n_out; ++i)
        PKM_CUDA(cudaMemcpyAsync(outs[i], dev[i], sizeof(double) * (size_t)marginal_size(pl, npix, keep_masks[i]),
                                 cudaMemcpyDeviceToHost, st));
      PKM_CUDA(cudaStreamSynchronize(st));
    }
  });
}

extern "C" int32_t pkm_reduce_logp_from(pkm_plan* pl, const double* log_P_src, uint32_t src_mask, int64_t npix,
                                        const double* log_dv, const double* light_weights, uint32_t keep_mask,
                                        int32_t density, double* out, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(pl && log_P_src && log_dv && out && npix >= 1, "bad argument");
    PKM_REQUIRE(src_mask <= 0xF && keep_mask <= 0xF, "bad mask");
    PKM_REQUIRE(pkm_is_device_pointer(log_P_src) && pkm_is_device_pointer(out), "device buffers required");
    use_device(pl->d.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const ReduceConsts c = upload_reduce_consts(pl, log_dv, light_weights, st);
    reduce_device(pl, log_P_src, src_mask, false, npix, c, light_weights != nullptr, keep_mask, density, out, st);
  });
}

extern "C" int32_t pkm_conditional(pkm_plan* pl, const double* log_P_joint, uint32_t joint_mask,
                                   const double* log_P_given, uint32_t given_mask, int64_t npix,
                                   const double* log_dv, int32_t density, int32_t exponentiate, double* out,
                                   void* stream) {
  return guarded([&] {
    PKM_REQUIRE(pl && log_P_joint && log_dv && out && npix >= 1, "bad argument");
    PKM_REQUIRE(joint_mask <= 0xF && (given_mask & ~joint_mask) == 0, "the conditioners must be axes of the joint");
    PKM_REQUIRE(pkm_is_device_pointer(log_P_joint) && pkm_is_device_pointer(out) &&
                (!log_P_given || pkm_is_device_pointer(log_P_given)), "device buffers required");
    PKM_REQUIRE((given_mask == 0) == (log_P_given == nullptr) || given_mask == 0, "given_mask without log_P_given");
    use_device(pl->d.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const ReduceConsts c = upload_reduce_consts(pl, log_dv, nullptr, st);
    launch_conditional(pl, log_P_joint, joint_mask, log_P_given, log_P_given ? given_mask : 0, npix, c.log_dv,
                       density, exponentiate, out, st);
  });
}

extern "C" int32_t pkm_moments_logp(const double* log_joint, const double* log_given, const double* values,
                                    int64_t nouter, int64_t nvar, int64_t ninner, double* out, int32_t device,
                                    void* stream) {
  return guarded([&] {
    PKM_REQUIRE(log_joint && values && out && nouter >= 1 && nvar >= 1 && ninner >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(pkm_is_device_pointer(log_joint) && pkm_is_device_pointer(out) && pkm_is_device_pointer(values) &&
                (!log_given || pkm_is_device_pointer(log_given)), "device buffers required");
    use_device(device);
    launch_moments_logp(log_joint, log_given, values, nouter, nvar, ninner, out, reinterpret_cast<cudaStream_t>(stream));
  });
}

extern "C" int32_t pkm_moments(const double* P, const double* values, int64_t nvar, int64_t ncond,
                               double* out, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(P && values && out && nvar >= 1 && ncond >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool host = !pkm_is_device_pointer(P);
    PKM_REQUIRE(host == !pkm_is_device_pointer(out) && host == !pkm_is_device_pointer(values),
                "P, values and out must be on the same side");
    if (!host) {
      launch_moments(P, values, nvar, ncond, out, st);
      return;
    }
    DevBuf buf;
    const size_t nP = (size_t)nvar * ncond;
    buf.reserve(sizeof(double) * (nP + nvar + 4 * (size_t)ncond));
    double* dP = buf.as<double>();
    double* dV = dP + nP;
    double* dO = dV + nvar;
    PKM_CUDA(cudaMemcpyAsync(dP, P, sizeof(double) * nP, cudaMemcpyHostToDevice, st));
    PKM_CUDA(cudaMemcpyAsync(dV, values, sizeof(double) * nvar, cudaMemcpyHostToDevice, st));
    launch_moments(dP, dV, nvar, ncond, dO, st);
    PKM_CUDA(cudaMemcpyAsync(out, dO, sizeof(double) * 4 * (size_t)ncond, cudaMemcpyDeviceToHost, st));
    PKM_CUDA(cudaStreamSynchronize(st));
    buf.release();
  });
}

extern "C" int32_t pkm_max(const double* x, int64_t n, double* result_host, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(x && result_host && n >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool host = !pkm_is_device_pointer(x);
    DevBuf buf;
    buf.reserve(sizeof(double) * (2 + (host ? (size_t)n : 0)));
    double* d_res = buf.as<double>();
    const double* dx = x;
    if (host) {
      PKM_CUDA(cudaMemcpyAsync(d_res + 2, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
      dx = d_res + 2;
    }
    launch_max(dx, n, d_res, st);
    PKM_CUDA(cudaMemcpyAsync(result_host, d_res, sizeof(double), cudaMemcpyDeviceToHost, st));
    PKM_CUDA(cudaStreamSynchronize(st));
    buf.release();
  });
}

extern "C" int32_t pkm_noise(const double* ybar, int64_t n, double snr, int32_t mode, uint64_t seed,
                             uint64_t offset, double ybar_max, double* yobs, double* sigma_out,
                             int32_t device, void* stream) {
  return guarded([&] {
    PkmRange range_call("pkm_noise");
    PKM_REQUIRE(ybar && n >= 1 && (yobs || sigma_out), "bad argument");
    PKM_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (ShotNoise) or 1 (ConstantSNR)");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool host = !pkm_is_device_pointer(ybar);
    if (yobs) PKM_REQUIRE(host == !pkm_is_device_pointer(yobs), "ybar and yobs must be on the same side");
    if (sigma_out) PKM_REQUIRE(host == !pkm_is_device_pointer(sigma_out), "ybar and sigma_out must be on the same side");
    DevBuf buf;
    const size_t nn = (size_t)n;
    buf.reserve(sizeof(double) * (2 + (host ? 3 * nn : 0)));
    double* d_max = buf.as<double>();
    const double* dy = ybar;
    double *dobs = yobs, *dsig = sigma_out;
    if (host) {
      double* base = d_max + 2;
      PKM_CUDA(cudaMemcpyAsync(base, ybar, sizeof(double) * nn, cudaMemcpyHostToDevice, st));
      dy = base;
      dobs = yobs ? base + nn : nullptr;
      dsig = sigma_out ? base + 2 * nn : nullptr;
    }
    if (mode == 0) {
      // numpy: max(ybar)**0.5 - a negative or NaN maximum gives NaN everywhere (noise.py:68-70)
      if (ybar_max >= 0.0 || ybar_max != ybar_max) {
        PKM_CUDA(cudaMemcpyAsync(d_max, &ybar_max, sizeof(double), cudaMemcpyHostToDevice, st));
        PKM_CUDA(cudaStreamSynchronize(st));
      } else {
        launch_max(dy, n, d_max, st);
      }
    }
    launch_noise(dy, n, snr, mode, seed, offset, d_max, dobs, dsig, st);
    if (host) {
      if (yobs) PKM_CUDA(cudaMemcpyAsync(yobs, dobs, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
      if (sigma_out) PKM_CUDA(cudaMemcpyAsync(sigma_out, dsig, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
    }
    PKM_CUDA(cudaStreamSynchronize(st));
    buf.release();
  });
}

// ----------------------------------------------------------------------------
// parametric-model maps on the device
// ----------------------------------------------------------------------------
namespace {
// small host arrays -> one device scratch buffer; returns device pointers in order
struct Staged {
  DevBuf buf;
  std::vector<const double*> ptr;
};
void stage_small(Staged& s, std::initializer_list<std::pair<const double*, size_t>> arrays, cudaStream_t st) {
  size_t total = 0;
  for (auto& a : arrays) total += a.second;
  std::vector<double> h;
  h.reserve(total);
  for (auto& a : arrays) h.insert(h.end(), a.first, a.first + a.second);
  s.buf.reserve(sizeof(double) * std::max<size_t>(total, 1));
  PKM_CUDA(cudaMemcpyAsync(s.buf.p, h.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
  PKM_CUDA(cudaStreamSynchronize(st));   // h is a local
  size_t off = 0;
  for (auto& a : arrays) {
    s.ptr.push_back(s.buf.as<double>() + off);
    off += a.second;
  }
}
int sm_count(int device) {
  cudaDeviceProp prop;
  PKM_CUDA(cudaGetDeviceProperties(&prop, device));
  return prop.multiProcessorCount;
}
}  // namespace

extern "C" int32_t pkm_disk_maps(const double* x1, const double* x2, int32_t nx1, int32_t nx2, double cx, double cy,
                                 double rotation, const double* age_pars, int32_t nt, double scale,
                                 const double* t_dep_pars, const double* t_dep_grid, int32_t ngrid,
                                 double* log_p_x_t, double* mu_v, double* sig_v, double* t_dep, int32_t* row_idx,
                                 double log_dxdy, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(x1 && x2 && age_pars && log_p_x_t && mu_v && sig_v && nx1 >= 1 && nx2 >= 1 && nt >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(pkm_is_device_pointer(log_p_x_t) && pkm_is_device_pointer(mu_v) && pkm_is_device_pointer(sig_v),
                "outputs must be device buffers");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nsm = sm_count(device);
    Staged sm;
    stage_small(sm, {{x1, (size_t)nx1}, {x2, (size_t)nx2}, {age_pars, (size_t)10 * nt},
                     {t_dep_grid ? t_dep_grid : x1, (size_t)(t_dep_grid ? ngrid : 0)}}, st);
    ModelFrame mf{cx, cy, std::cos(rotation), std::sin(rotation), nx1, nx2};
    const int64_t npix = (int64_t)nx1 * nx2;
    launch_disk_maps(mf, sm.ptr[0], sm.ptr[1], sm.ptr[2], nt, scale, log_p_x_t, mu_v, sig_v, nsm, st);
    launch_normalise_rows(log_p_x_t, nt, npix, log_dxdy, st);
    if (t_dep && row_idx) {
      PKM_REQUIRE(t_dep_pars && t_dep_grid && ngrid >= 1, "t_dep needs its parameters and the tabulated grid");
      PKM_REQUIRE(pkm_is_device_pointer(t_dep) && pkm_is_device_pointer(row_idx), "outputs must be device buffers");
      launch_disk_t_dep(mf, sm.ptr[0], sm.ptr[1], t_dep_pars[0], t_dep_pars[1], t_dep_pars[2], t_dep_pars[3], scale,
                        sm.ptr[3], ngrid, t_dep, row_idx, nsm, st);
    }
    PKM_CUDA(cudaStreamSynchronize(st));   // the staging buffer dies with this frame
  });
}

extern "C" int32_t pkm_stream_maps(const double* x1, const double* x2, int32_t nx1, int32_t nx2, double cx, double cy,
                                   double rotation, const double* centres_x, const double* centres_y, int32_t nsmp,
                                   double sig, double half_dx, double half_dy, double theta_lo, double theta_hi,
                                   double v_lo, double v_hi, double* log_p_x, double* mu_v, double log_dxdy,
                                   int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(x1 && x2 && centres_x && centres_y && log_p_x && mu_v && nx1 >= 1 && nx2 >= 1 && nsmp >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(pkm_is_device_pointer(log_p_x) && pkm_is_device_pointer(mu_v), "outputs must be device buffers");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Staged sm;
    stage_small(sm, {{x1, (size_t)nx1}, {x2, (size_t)nx2}, {centres_x, (size_t)nsmp}, {centres_y, (size_t)nsmp}}, st);
    ModelFrame mf{cx, cy, std::cos(rotation), std::sin(rotation), nx1, nx2};
    launch_stream_maps(mf, sm.ptr[0], sm.ptr[1], sm.ptr[2], sm.ptr[3], nsmp, sig, half_dx, half_dy, theta_lo, theta_hi,
                       v_lo, v_hi, log_p_x, mu_v, sm_count(device), st);
    launch_normalise_rows(log_p_x, 1, (int64_t)nx1 * nx2, 0.0, st);   // mass normalised to one over the cube ...
    PKM_CUDA(cudaStreamSynchronize(st));
    (void)log_dxdy;
  });
}

extern "C" int32_t pkm_assemble_txz(const double* log_P_t, const double* log_dt, const double* log_p_x,
                                    int32_t px_per_age, double log_dxdy, const double* log_Pz_rows, int32_t nrow,
                                    const double* log_dz, const int32_t* row_idx, int32_t nt, int64_t npix,
                                    int32_t nz, double* P_txz, double* log_p_txz, int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(log_P_t && log_dt && log_p_x && log_Pz_rows && log_dz && (P_txz || log_p_txz) && nt >= 1 && npix >= 1 &&
                nz >= 1 && nrow >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(pkm_is_device_pointer(log_p_x) && (!row_idx || pkm_is_device_pointer(row_idx)) &&
                (!P_txz || pkm_is_device_pointer(P_txz)) && (!log_p_txz || pkm_is_device_pointer(log_p_txz)),
                "maps and outputs must be device buffers");
    PKM_REQUIRE(row_idx || nrow == 1, "several enrichment rows need a row index map");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Staged sm;
    stage_small(sm, {{log_P_t, (size_t)nt}, {log_dt, (size_t)nt}, {log_Pz_rows, (size_t)nrow * nt * nz},
                     {log_dz, (size_t)nz}}, st);
    launch_assemble_txz(sm.ptr[0], sm.ptr[1], log_p_x, px_per_age ? npix : 0, log_dxdy, sm.ptr[2], sm.ptr[3],
                        reinterpret_cast<const int*>(row_idx), nt, npix, nz, P_txz, log_p_txz, sm_count(device), st);
    PKM_CUDA(cudaStreamSynchronize(st));
  });
}

extern "C" int32_t pkm_ssp_operand(const double* coef, int32_t ncoef, const int32_t* tap_idx, const double* tap_w,
                                   int32_t n_pix, int32_t nssp, int32_t n_fft, double* Xw_out, double* FXw_out,
                                   int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(coef && tap_idx && tap_w && Xw_out && FXw_out, "null argument");
    PKM_REQUIRE(ncoef >= 4 && n_pix >= 1 && nssp >= 1 && n_fft >= n_pix, "bad shape (n_fft must be >= n_pix)");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    for (int j = 0; j < n_pix; ++j)
      PKM_REQUIRE(tap_idx[j] >= 0 && tap_idx[j] + 4 <= ncoef, "tap_idx[%d] out of range", j);
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nfo = n_fft / 2 + 1;
    // exp(-2 pi i r / n) in long double, rounded once
    const long double PI_L = 3.14159265358979323846264338327950288419716939937510L;
    std::vector<double> wn((size_t)2 * n_fft);
    for (int r = 0; r < n_fft; ++r) {
      const long double ang = -2.0L * PI_L * (long double)r / (long double)n_fft;
      wn[2 * r] = (double)cosl(ang);
      wn[2 * r + 1] = (double)sinl(ang);
    }
    const size_t n_c = (size_t)ncoef * nssp, n_x = (size_t)n_pix * nssp, n_f = (size_t)2 * nfo * nssp;
    DevBuf buf;
    buf.reserve(sizeof(double) * (n_c + 4 * (size_t)n_pix + wn.size() + n_x + n_f) + sizeof(int) * ((size_t)n_pix + 2));
    double* d_coef = buf.as<double>();
    double* d_w = d_coef + n_c;
    double* d_wn = d_w + 4 * (size_t)n_pix;
    double* d_X = d_wn + wn.size();
    double* d_F = d_X + n_x;
    int* d_idx = reinterpret_cast<int*>(d_F + n_f);
    PKM_CUDA(cudaMemcpyAsync(d_coef, coef, sizeof(double) * n_c, cudaMemcpyHostToDevice, st));
    PKM_CUDA(cudaMemcpyAsync(d_w, tap_w, sizeof(double) * 4 * n_pix, cudaMemcpyHostToDevice, st));
    PKM_CUDA(cudaMemcpyAsync(d_wn, wn.data(), sizeof(double) * wn.size(), cudaMemcpyHostToDevice, st));
    PKM_CUDA(cudaMemcpyAsync(d_idx, tap_idx, sizeof(int) * n_pix, cudaMemcpyHostToDevice, st));
    launch_ssp_operand(d_coef, d_idx, d_w, reinterpret_cast<const double2*>(d_wn), n_pix, nssp, n_fft, d_X,
                       reinterpret_cast<double2*>(d_F), sm_count(device), st);
    PKM_CUDA(cudaMemcpyAsync(Xw_out, d_X, sizeof(double) * n_x, cudaMemcpyDefault, st));
    PKM_CUDA(cudaMemcpyAsync(FXw_out, d_F, sizeof(double) * n_f, cudaMemcpyDefault, st));
    PKM_CUDA(cudaStreamSynchronize(st));
  });
}

extern "C" int32_t pkm_fp64_fma_peak(int32_t device, int32_t repeats, double* tflops_out) {
  return guarded([&] {
    PKM_REQUIRE(tflops_out && repeats >= 1, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible");
    use_device(device);
    cudaDeviceProp prop;
    PKM_CUDA(cudaGetDeviceProperties(&prop, device));
    *tflops_out = run_fma_probe(prop.multiProcessorCount, repeats);
  });
}

extern "C" int32_t pkm_synth_log_density(double* out, int32_t nt, int32_t nv, int64_t total_pix, int64_t pix0,
                                         int64_t npix, int32_t nz, uint64_t seed, double scale,
                                         int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(out && nt >= 1 && nv >= 1 && nz >= 1 && npix >= 0 && pix0 >= 0 && pix0 + npix <= total_pix,
                "bad argument");
    PKM_REQUIRE(scale > 0.0, "scale must be positive");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    PKM_REQUIRE(pkm_is_device_pointer(out), "out must be device memory");
    use_device(device);
    cudaDeviceProp prop;
    PKM_CUDA(cudaGetDeviceProperties(&prop, device));
    launch_synth_logp(out, nt, nv, total_pix, pix0, npix, nz, seed, scale, prop.multiProcessorCount,
                      reinterpret_cast<cudaStream_t>(stream));
  });
}

extern "C" int32_t pkm_histogram5d(const double* const* samples, const double* weights, int64_t n,
                                   const double* const* edges, const int32_t* nbin, double* out,
                                   int32_t device, void* stream) {
  return guarded([&] {
    PKM_REQUIRE(samples && edges && nbin && out && n >= 0, "bad argument");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaDeviceProp prop;
    PKM_CUDA(cudaGetDeviceProperties(&prop, device));
    int nb[5];
    int64_t nbins = 1;
    std::vector<double> h_edges;
    for (int a = 0; a < 5; ++a) {
      PKM_REQUIRE(nbin[a] >= 1 && samples[a] && edges[a], "bad axis %d", a);
      nb[a] = nbin[a];
      nbins *= nbin[a];
      h_edges.insert(h_edges.end(), edges[a], edges[a] + nbin[a] + 1);   // edges are host arrays
    }
    PKM_REQUIRE(h_edges.size() * sizeof(double) <= 40 * 1024, "too many bin edges");
    const bool host = !pkm_is_device_pointer(samples[0]);
    for (int a = 1; a < 5; ++a)
      PKM_REQUIRE(host == !pkm_is_device_pointer(samples[a]), "all sample arrays must be on the same side");
    if (weights) PKM_REQUIRE(host == !pkm_is_device_pointer(weights), "weights must be on the same side as the samples");
    const bool out_host = !pkm_is_device_pointer(out);
    DevBuf buf, obuf, ebuf;
    ebuf.reserve(sizeof(double) * h_edges.size());
    PKM_CUDA(cudaMemcpyAsync(ebuf.p, h_edges.data(), sizeof(double) * h_edges.size(), cudaMemcpyHostToDevice, st));
    const double* d_s[5];
    const double* d_w = weights;
    if (host) {
      buf.reserve(sizeof(double) * (size_t)std::max<int64_t>(n, 1) * 6);
      double* base = buf.as<double>();
      for (int a = 0; a < 5; ++a) {
        PKM_CUDA(cudaMemcpyAsync(base + (size_t)a * n, samples[a], sizeof(double) * n, cudaMemcpyHostToDevice, st));
        d_s[a] = base + (size_t)a * n;
      }
      if (weights) {
        PKM_CUDA(cudaMemcpyAsync(base + (size_t)5 * n, weights, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        d_w = base + (size_t)5 * n;
      }
    } else {
      for (int a = 0; a < 5; ++a) d_s[a] = samples[a];
    }
    double* d_out = out;
    if (out_host) {
      obuf.reserve(sizeof(double) * nbins);
      d_out = obuf.as<double>();
    }
    launch_hist5d(d_s, d_w, n, ebuf.as<double>(), nb, d_out, prop.multiProcessorCount, st);
    if (out_host) PKM_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * nbins, cudaMemcpyDeviceToHost, st));
    PKM_CUDA(cudaStreamSynchronize(st));
    buf.release(); obuf.release(); ebuf.release();
  });
}

extern "C" int32_t pkm_pixelate_gaussian(const double* log_w_txz, const double* mu_v, const double* sig_v,
                                         const double* v_edg, int32_t nt, int32_t nv, int64_t npix,
                                         int32_t nz, int32_t density, double* out, int32_t device,
                                         void* stream) {
  return guarded([&] {
    PkmRange range_call("pkm_pixelate_gaussian");
    PKM_REQUIRE(log_w_txz && mu_v && sig_v && v_edg && out, "bad argument");
    PKM_REQUIRE(nt >= 1 && nv >= 1 && npix >= 1 && nz >= 1, "bad shape");
    if (pkm_device_count() <= 0) PKM_THROW(PKM_ERR_CUDA, "no CUDA device visible (there is no CPU fallback)");
    use_device(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n_txz = (size_t)nt * npix * nz, n_tx = (size_t)nt * npix, n_out = n_txz * nv;
    const bool w_host = !pkm_is_device_pointer(log_w_txz), m_host = !pkm_is_device_pointer(mu_v),
               s_host = !pkm_is_device_pointer(sig_v), o_host = !pkm_is_device_pointer(out);
    std::vector<double> h_tab(2 * (size_t)nv + 1);           // v_edg [nv+1], log dv [nv]
    for (int i = 0; i <= nv; ++i) h_tab[i] = v_edg[i];
    for (int i = 0; i < nv; ++i) h_tab[nv + 1 + i] = std::log(v_edg[i + 1] - v_edg[i]);
    DevBuf buf;
    buf.reserve(sizeof(double) * (h_tab.size() + (w_host ? n_txz : 0) + (m_host ? n_tx : 0) +
                                  (s_host ? n_tx : 0) + (o_host ? n_out : 0)));
    double* cur = buf.as<double>();
    double* d_tab = cur; cur += h_tab.size();
    PKM_CUDA(cudaMemcpyAsync(d_tab, h_tab.data(), sizeof(double) * h_tab.size(), cudaMemcpyHostToDevice, st));
    auto stage = [&](const double* src, bool host, size_t n) -> const double* {
      if (!host) return src;
      double* d = cur; cur += n;
      PKM_CUDA(cudaMemcpyAsync(d, src, sizeof(double) * n, cudaMemcpyHostToDevice, st));
      return d;
    };
    const double* d_w = stage(log_w_txz, w_host, n_txz);
    const double* d_m = stage(mu_v, m_host, n_tx);
    const double* d_s = stage(sig_v, s_host, n_tx);
    double* d_out = o_host ? cur : out;
    launch_pixelate(d_w, d_m, d_s, d_tab, density ? d_tab + nv + 1 : nullptr, nt, nv, npix, nz, d_out, st);
    if (o_host) PKM_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
    PKM_CUDA(cudaStreamSynchronize(st));   // h_tab and the staging buffer die with this frame
    buf.release();
  });
}
