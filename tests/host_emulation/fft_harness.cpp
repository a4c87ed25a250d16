// Host emulation of the inverse-FFT kernels (tests/test_fft_host_emulation.py): the DEVICE code of
// popkinmocks_b200/csrc/fft_kernels.cu (the text between `namespace {` and the four-step variant,
// extracted by the test into body.inc) is compiled as plain C++; a "CTA" is a set of host threads
// (thread_local threadIdx, __syncthreads = a pthread barrier, one shared sbuf), run on the real
// tables (tables.cpp) and compared with a long-double O(n^2) inverse real DFT.  It checks
// the index algebra, twiddles and table layouts of k_irfft_czt / k_irfft_czt_h without a GPU.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <algorithm>
#include "pkm_internal.h"
#undef __device__
#undef __host__
#undef __forceinline__
#undef __global__
#undef __restrict__
#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __restrict__
#define __ldg(p) (*(p))
#include <pthread.h>
#include <thread>
#include <functional>
static pthread_barrier_t cta_barrier;
#define __syncthreads() pthread_barrier_wait(&cta_barrier)
struct Dim { int x; };
static thread_local Dim threadIdx_{0};
static Dim blockDim_{1}, blockIdx_{0}, gridDim_{1};
// run `body` as one CTA of nthr threads
static void launch(int nthr, const std::function<void()>& body) {
  blockDim_.x = nthr;
  pthread_barrier_init(&cta_barrier, nullptr, nthr);
  std::vector<std::thread> th;
  for (int t = 0; t < nthr; ++t) th.emplace_back([t, &body] { threadIdx_.x = t; body(); });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&cta_barrier);
}
#define threadIdx threadIdx_
#define blockDim blockDim_
#define blockIdx blockIdx_
#define gridDim gridDim_
static double2 sbuf[20000];
// ldg_pinned is inline PTX in the device code
#define PKM_HOST_EMULATION 1
constexpr int CZT_THREADS = 512;
using std::min;
using std::max;
namespace {
#include "body.inc"
}
static int run(int n) {
  CztTables c;
  build_czt_tables(n, c);
  const double2* tw = reinterpret_cast<const double2*>(c.twiddle.data());
  const double2* bf = reinterpret_cast<const double2*>(c.bfilt.data());
  if (c.R != 1) { printf("n=%d M=%d skipped (four-step transform)\n", n, c.M); return 0; }
  std::vector<double2> Y(c.nfo);
  srand(n);
  for (auto& v : Y) v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
  std::vector<double> ref(n);
  double mx = 0;
  for (int m = 0; m < n; ++m) {
    long double acc = 0;
    for (int k = 0; k < c.nfo; ++k) {
      double2 a = Y[k];
      if (c.alpha[k] == 1.0) a.y = 0;
      long double ang = 2.0L * M_PIl * (long double)(((long long)k * m) % n) / n;
      acc += c.alpha[k] * (a.x * cosl(ang) - a.y * sinl(ang));
    }
    ref[m] = (double)(acc / n);
    mx = std::max(mx, fabs(ref[m]));
  }
  const double s = 1.0 / ((double)c.M * n);
  auto err = [&](const std::vector<double>& o) {
    double e = 0;
    for (int m = 0; m < n; ++m) e = std::max(e, fabs(o[m] - ref[m]));
    return e / mx;
  };
  std::vector<double2> B(c.M);
  std::vector<double> o(n);
  launch(64, [&] { k_filter_spectrum(bf, B.data(), c.M, c.logM, tw, nullptr); });
  launch(64, [&] {
    k_irfft_czt(Y.data(), 1, n, c.nfo, c.M, c.logM, s, reinterpret_cast<const double2*>(c.chi.data()), c.alpha.data(),
                tw, B.data(), o.data(), nullptr);
  });
  printf("n=%d M=%d generic %.3e", n, c.M, err(o));
  if (!c.pre.empty()) {
    const double2* htw = reinterpret_cast<const double2*>(c.htw.data());
    const int nthr = std::max(32, std::min(512, c.M / 16));
    launch(nthr, [&] { k_filter_spectrum_h(bf, B.data(), c.M, c.logM, tw, htw); });
    const double2* pre = reinterpret_cast<const double2*>(c.pre.data());
    const double2* post = reinterpret_cast<const double2*>(c.post.data());
    {
      // two pixels (the same spectrum twice): the second one runs on the prefetched registers
      std::vector<double2> Y2(Y);
      Y2.insert(Y2.end(), Y.begin(), Y.end());
      std::vector<double> o2(2 * (size_t)n, 0.0);
      launch(nthr, [&] {
        const int lh = c.logM - 1;
#define CALL(L) k_irfft_czt_h<L>(Y2.data(), 2, n, c.nfo, c.M, c.logM, s, pre, post, htw, B.data(), o2.data())
        if (lh == 4) CALL(4); else if (lh == 8) CALL(8); else CALL(12);
#undef CALL
      });
      std::vector<double> oa(o2.begin(), o2.begin() + n), ob(o2.begin() + n, o2.end());
      printf(" halves %.3e %.3e", err(oa), err(ob));
    }
  }
  printf("\n");
  return 0;
}
int main(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) run(atoi(argv[i]));
  return 0;
}
