// Host-side construction of the constant tables used by the kernels.
//
// * not-a-knot cubic spline nfi -> nfo on linspace(0,1,.)  : the reference resamples the
//   velocity spectrum with scipy.interpolate.interp1d(kind="cubic")
//   (/root/reference/popkinmocks/components/base.py:846-853), which is
//   make_interp_spline(k=3) with not-a-knot end conditions: knots are the data sites
//   with the 2nd and the penultimate removed and the ends repeated 4 times.  The
//   spline is  s(x) = sum_m c_m B_m(x)  with  c = A^{-1} y,  A[j][m] = B_m(x_j).
// * folded real-DFT matrices composing roll, rfft, *dv (base.py:841-844) with A^{-1}.
// * slot layout of the contraction kernel and chirp-z tables of the inverse FFT.
//
// Everything here is plain C++ (long double accumulation, exact integer phase
// reduction) and runs without a GPU, so it is unit-tested on CPU against SciPy.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "pkm_internal.h"

namespace {

const long double PI_L = 3.14159265358979323846264338327950288419716939937510L;

// np.linspace(0, 1, n)[j]  (numpy: arange(n) * step with step = 1/(n-1); endpoint forced to 1)
inline double linspace01(int j, int n) {
  if (j == n - 1) return 1.0;
  double step = 1.0 / double(n - 1);
  return double(j) * step;
}

struct Knots {
  std::vector<long double> t;  // nfi + 4 knots
  int nfi;
};

Knots not_a_knot_knots(int nfi) {
  Knots k;
  k.nfi = nfi;
  for (int i = 0; i < 4; ++i) k.t.push_back(0.0L);
  for (int j = 2; j <= nfi - 3; ++j) k.t.push_back((long double)linspace01(j, nfi));
  for (int i = 0; i < 4; ++i) k.t.push_back(1.0L);
  return k;
}

// knot span ell (3 <= ell <= nfi-1) with t[ell] <= x < t[ell+1]; x == 1 maps to the last span
int find_span(const Knots& k, long double x) {
  int lo = 3, hi = k.nfi;  // t[nfi] == 1
  if (x >= k.t[hi]) return k.nfi - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (x >= k.t[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

// the 4 non-zero cubic B-splines B_{ell-3..ell}(x)   (Cox - de Boor, triangular scheme)
void basis4(const Knots& k, int ell, long double x, long double N[4]) {
  long double left[4], right[4];
  N[0] = 1.0L;
  for (int j = 1; j <= 3; ++j) {
    left[j] = x - k.t[ell + 1 - j];
    right[j] = k.t[ell + j] - x;
    long double saved = 0.0L;
    for (int r = 0; r < j; ++r) {
      long double tmp = N[r] / (right[r + 1] + left[j - r]);
      N[r] = saved + right[r + 1] * tmp;
      saved = left[j - r] * tmp;
    }
    N[j] = saved;
  }
}

// dense inverse by Gauss-Jordan with partial pivoting, long double
void invert(std::vector<long double>& a, int n, std::vector<long double>& inv) {
  inv.assign(size_t(n) * n, 0.0L);
  for (int i = 0; i < n; ++i) inv[size_t(i) * n + i] = 1.0L;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (fabsl(a[size_t(r) * n + col]) > fabsl(a[size_t(piv) * n + col])) piv = r;
    if (a[size_t(piv) * n + col] == 0.0L) PKM_THROW(PKM_ERR_INVALID, "singular collocation matrix");
    if (piv != col)
      for (int c = 0; c < n; ++c) {
        std::swap(a[size_t(piv) * n + c], a[size_t(col) * n + c]);
        std::swap(inv[size_t(piv) * n + c], inv[size_t(col) * n + c]);
      }
    long double d = 1.0L / a[size_t(col) * n + col];
    for (int c = 0; c < n; ++c) {
      a[size_t(col) * n + c] *= d;
      inv[size_t(col) * n + c] *= d;
    }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      long double f = a[size_t(r) * n + col];
      if (f == 0.0L) continue;
      for (int c = 0; c < n; ++c) {
        a[size_t(r) * n + c] -= f * a[size_t(col) * n + c];
        inv[size_t(r) * n + c] -= f * inv[size_t(col) * n + c];
      }
    }
  }
}

}  // namespace

void build_spline_tables(int nfi, int nfo, SplineTables& out) {
  PKM_REQUIRE(nfi >= 4, "cubic spline needs at least 4 input frequencies (nv >= 6), got nfi=%d", nfi);
  PKM_REQUIRE(nfo >= 2, "nfo=%d too small", nfo);
  out.nfi = nfi;
  out.nfo = nfo;
  Knots kn = not_a_knot_knots(nfi);
  std::vector<long double> A(size_t(nfi) * nfi, 0.0L), Ai;
  for (int j = 0; j < nfi; ++j) {
    long double x = (long double)linspace01(j, nfi);
    int ell = find_span(kn, x);
    long double N[4];
    basis4(kn, ell, x, N);
    for (int q = 0; q < 4; ++q) A[size_t(j) * nfi + (ell - 3 + q)] = N[q];
  }
  invert(A, nfi, Ai);
  out.A_inv.resize(size_t(nfi) * nfi);
  for (size_t i = 0; i < Ai.size(); ++i) out.A_inv[i] = (double)Ai[i];
  out.tap_idx.resize(nfo);
  out.tap_w.resize(size_t(nfo) * 4);
  for (int k = 0; k < nfo; ++k) {
    long double x = (long double)linspace01(k, nfo);
    int ell = find_span(kn, x);
    long double N[4];
    basis4(kn, ell, x, N);
    out.tap_idx[k] = ell - 3;
    for (int q = 0; q < 4; ++q) out.tap_w[size_t(k) * 4 + q] = (double)N[q];
  }
}

void build_fold_tables(int nv, int v0_idx, double dv, const SplineTables& sp, FoldTables& out) {
  const int nfi = nv / 2 + 1;
  PKM_REQUIRE(sp.nfi == nfi, "spline tables built for nfi=%d, need %d", sp.nfi, nfi);
  PKM_REQUIRE(v0_idx >= 0 && v0_idx < nv, "v0_idx=%d out of range", v0_idx);
  out.nv = nv;
  out.nfi = nfi;
  out.ne = nv / 2 + 1;
  out.no = (nv - 1) / 2;
  out.idx_plus.resize(out.ne);
  out.idx_minus.resize(out.ne);
  for (int u = 0; u < out.ne; ++u) {
    // rolled[u] = p[(u + v0) mod nv]  (np.roll(p, nv - v0_idx), base.py:843)
    out.idx_plus[u] = (u + v0_idx) % nv;
    out.idx_minus[u] = ((v0_idx - u) % nv + nv) % nv;
  }
  // F[j] = dv * sum_u rolled[u] exp(-2 pi i j u / nv);  with e[u] = rolled[u] + rolled[nv-u],
  // o[u] = rolled[u] - rolled[nv-u]:  Re F[j] = dv * sum_{u<ne} g_u e[u] cos(2 pi j u/nv),
  // Im F[j] = -dv * sum_{1<=u<=no} o[u] sin(2 pi j u/nv), g_u = 1/2 for the self-paired
  // positions (u = 0, and u = nv/2 when nv is even), 1 otherwise.
  std::vector<long double> C(size_t(nfi) * out.ne), S(size_t(nfi) * std::max(out.no, 1));
  for (int j = 0; j < nfi; ++j) {
    for (int u = 0; u < out.ne; ++u) {
      long long r = ((long long)j * u) % nv;
      long double ang = 2.0L * PI_L * (long double)r / (long double)nv;
      bool self = (out.idx_plus[u] == out.idx_minus[u]);
      C[size_t(j) * out.ne + u] = (long double)dv * (self ? 0.5L : 1.0L) * cosl(ang);
      if (u >= 1 && u <= out.no) S[size_t(j) * out.no + (u - 1)] = -(long double)dv * sinl(ang);
    }
  }
  out.CR.assign(size_t(nfi) * out.ne, 0.0);
  out.CI.assign(size_t(nfi) * std::max(out.no, 1), 0.0);
  for (int m = 0; m < nfi; ++m) {
    for (int u = 0; u < out.ne; ++u) {
      long double acc = 0.0L;
      for (int j = 0; j < nfi; ++j) acc += (long double)sp.A_inv[size_t(m) * nfi + j] * C[size_t(j) * out.ne + u];
      out.CR[size_t(m) * out.ne + u] = (double)acc;
    }
    for (int u = 0; u < out.no; ++u) {
      long double acc = 0.0L;
      for (int j = 0; j < nfi; ++j) acc += (long double)sp.A_inv[size_t(m) * nfi + j] * S[size_t(j) * out.no + u];
      out.CI[size_t(m) * out.no + u] = (double)acc;
    }
  }
}

void build_wseg_tables(const SplineTables& sp, int kt, WsegTables& out) {
  const int nfo = sp.nfo;
  out.kt = kt;
  out.seg.clear();
  out.w.clear();
  int k = 0;
  while (k < nfo) {
    const int span = sp.tap_idx[k];
    int cnt = 1;
    while (cnt < kt && k + cnt < nfo && sp.tap_idx[k + cnt] == span) ++cnt;
    out.seg.push_back(span);
    out.seg.push_back(k);
    out.seg.push_back(cnt);
    out.seg.push_back(0);
    for (int r = 0; r < kt; ++r)
      for (int q = 0; q < 4; ++q) out.w.push_back(r < cnt ? sp.tap_w[size_t(k + r) * 4 + q] : 0.0);
    k += cnt;
  }
  out.nwseg = (int)out.seg.size() / 4;
}

void build_czt_tables(int n, CztTables& out) {
  PKM_REQUIRE(n >= 2, "n_fft=%d too small", n);
  out.n = n;
  out.nfo = n / 2 + 1;
  int need = out.nfo + n - 1;
  int M = 1, logM = 0;
  while (M < need) { M <<= 1; ++logM; }
  if (logM < 2) { M = 4; logM = 2; }
  out.M = M;
  out.logM = logM;
  // chi[j] = exp(+i pi j^2 / n), phase reduced exactly: j^2 mod 2n
  out.chi.resize(size_t(2) * n);
  for (long long j = 0; j < n; ++j) {
    long long r = (j * j) % (2LL * n);
    long double ang = PI_L * (long double)r / (long double)n;
    out.chi[2 * j] = (double)cosl(ang);
    out.chi[2 * j + 1] = (double)sinl(ang);
  }
  // numpy.fft.irfft weights: DC and (even n) Nyquist count once and only their real
  // part is used; all other bins twice (Hermitian pair).
  out.alpha.assign(out.nfo, 2.0);
  out.alpha[0] = 1.0;
  if (n % 2 == 0) out.alpha[out.nfo - 1] = 1.0;
  // Transforms that fit shared memory (M <= 8192) run in one piece (R = 1, C = M); larger ones
  // use the four-step split M = R * C with C = 4096: R-point DFTs down the columns, twiddles
  // W_M^(n2 k1), then R shared-memory transforms of length C.
  out.C = M <= 8192 ? M : 4096;
  out.R = M / out.C;
  out.logC = 0;
  while ((1 << out.logC) < out.C) ++out.logC;
  const int C = out.C, R = out.R;
  // layout: [0, C) per-pass radix-4 twiddles of the length-C transform, contiguous in the
  // butterfly index j: the pass with quarter size q (block Nb = 4q, stride = C/Nb) reads
  // W_C^(m j stride), m = 1..3, at [C - 4q + (m-1) q + j];  [C, 2C) W_M^(n2), n2 < C;
  // [2C, 2C + R) W_R^(j).
  out.twiddle.assign(size_t(2) * (2 * size_t(C) + R), 0.0);
  for (int Nb = C; Nb >= 4; Nb >>= 2) {
    const int q = Nb >> 2, stride = C / Nb;
    for (int m = 1; m <= 3; ++m)
      for (int j = 0; j < q; ++j) {
        const long long idx = ((long long)m * j * stride) % C;
        long double ang = -2.0L * PI_L * (long double)idx / (long double)C;
        const size_t pos = size_t(C - 4 * q) + size_t(m - 1) * q + j;
        out.twiddle[2 * pos] = (double)cosl(ang);
        out.twiddle[2 * pos + 1] = (double)sinl(ang);
      }
  }
  for (int j = 0; j < C; ++j) {
    long double ang = -2.0L * PI_L * (long double)j / (long double)M;
    out.twiddle[2 * (size_t(C) + j)] = (double)cosl(ang);
    out.twiddle[2 * (size_t(C) + j) + 1] = (double)sinl(ang);
  }
  for (int j = 0; j < R; ++j) {
    long double ang = -2.0L * PI_L * (long double)j / (long double)R;
    out.twiddle[2 * (2 * size_t(C) + j)] = (double)cosl(ang);
    out.twiddle[2 * (2 * size_t(C) + j) + 1] = (double)sinl(ang);
  }
  // halves kernel tables (one-piece transforms with M = 2 * 16^p), phases reduced exactly:
  // chi_k (W_M^k)^h = exp(i pi (k^2 mod 2n)/n - 2 pi i h k/M)
  out.pre.clear();
  out.post.clear();
  out.htw.clear();
  if (R == 1 && M >= 32 && (logM - 1) % 4 == 0) {
    const int nfo = out.nfo, H = M / 2;
    for (int Nb = H; Nb > 16; Nb >>= 4)
      for (int j = 0; j < Nb / 16; ++j) {
        const long double ang = -2.0L * PI_L * (long double)j / (long double)Nb;
        out.htw.push_back((double)cosl(ang));
        out.htw.push_back((double)sinl(ang));
      }
    out.htw.resize(out.htw.size() + 2, 0.0);   // never empty (M = 32 has no such pass)
    out.pre.assign(size_t(4) * nfo, 0.0);
    for (int h = 0; h < 2; ++h)
      for (long long k = 0; k < nfo; ++k) {
        const long double ang = PI_L * (long double)((k * k) % (2LL * n)) / (long double)n -
                                2.0L * PI_L * (long double)(h * k) / (long double)M;
        out.pre[2 * (size_t(h) * nfo + k)] = out.alpha[k] * (double)cosl(ang);
        out.pre[2 * (size_t(h) * nfo + k) + 1] = out.alpha[k] * (double)sinl(ang);
      }
    out.post.assign(size_t(4) * n, 0.0);
    for (long long m = 0; m < n; ++m) {
      const long double a0 = PI_L * (long double)((m * m) % (2LL * n)) / (long double)n;
      out.post[2 * m] = (double)cosl(a0);
      out.post[2 * m + 1] = (double)sinl(a0);
      const long double a1 = a0 + 2.0L * PI_L * (long double)(m % H) / (long double)M;
      const double sgn = m < H ? 1.0 : -1.0;
      out.post[2 * (size_t(n) + m)] = sgn * (double)cosl(a1);
      out.post[2 * (size_t(n) + m) + 1] = sgn * (double)sinl(a1);
    }
  }
  // z[m] = chi[m] * sum_k (a[k] chi[k]) conj(chi[m-k]),  m-k in [-(nfo-1), n-1]
  out.bfilt.assign(size_t(2) * M, 0.0);
  for (long long j = -(long long)(out.nfo - 1); j <= (long long)(n - 1); ++j) {
    long long a = j < 0 ? -j : j;
    long long r = (a * a) % (2LL * n);
    long double ang = PI_L * (long double)r / (long double)n;
    long long pos = ((j % M) + M) % M;
    out.bfilt[2 * pos] = (double)cosl(ang);
    out.bfilt[2 * pos + 1] = -(double)sinl(ang);
  }
}
