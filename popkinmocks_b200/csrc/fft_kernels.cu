// Arbitrary-length batched inverse real FFT (numpy.fft.irfft semantics) for sm_100a.
//
// The reference calls np.fft.irfft(F_y, n_fft, axis=1) per (t,z) (base.py:858) or once per
// pixel (parametric.py:373); n_fft = len(ssps.w) is whatever np.arange yields - 1992, 4907 =
// 7*701, 1021 and 11069 (primes), ... - so a fixed-radix transform is not enough.
//
//   y[m] = (1/n) Re sum_{k<nfo} alpha_k Y'[k] e^{+2 pi i k m / n},     nfo = n/2+1
//
// (alpha = 1 for DC / Nyquist whose imaginary parts are ignored, 2 otherwise) is evaluated as
// a chirp-z transform: k m = (k^2 + m^2 - (m-k)^2)/2, i.e. a length nfo+n-1 linear convolution
// with the chirp chi[j] = e^{i pi j^2/n}, done with power-of-two FFTs of size M >= nfo+n-1.
// One CTA per pixel; the M-point transforms run in place in shared memory (M <= 8192) or in a
// CTA-private, L2-resident global scratch (larger n).  M = 2 * 16^p (8192 for the headline
// n_fft = 4907) takes the specialised "halves" kernel further down: two independent radix-16
// half-transforms, middle passes fused in registers, global operands requested ahead of use.  Forward = decimation in frequency,
// inverse = its exact stage-by-stage inverse, so no bit-reversal pass is needed: the filter
// spectrum is prepared once with the same forward routine (same permuted order).
// Chirp phases are reduced exactly (j^2 mod 2n in integers) on the host.
#include <algorithm>
#include <cstdlib>

#include "pkm_internal.h"

constexpr int CZT_THREADS = 512;   // M/16 radix-16 work items for M = 8192; 128 registers per thread

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// forward DIF: natural order in, digit-permuted order out.
// `in(idx)` supplies the inputs of the FIRST pass (so the caller's load / pre-multiply pass is
// fused into it); if `post` is non-null every output of the LAST pass is multiplied by
// post[idx] (the filter spectrum of the chirp-z convolution).
// Shared-memory index padding: one empty 16-byte slot after every 32 elements.  Without it the
// passes with small strides (radix-16 with quarter size 2 or 1) put 16 or 32 lanes of a warp on
// the same bank group; with it every pass of every transform size is conflict-free.
__device__ __forceinline__ int sp(int i) { return i + (i >> 5); }
__host__ __device__ constexpr size_t sp_elems(size_t n) { return n + (n >> 5); }

struct SmemInput {
  const double2* s;
  __device__ __forceinline__ double2 operator()(int i) const { return s[sp(i)]; }
};

// 4-point DIF butterfly (forward) and its inverse, twiddles applied by the caller
__device__ __forceinline__ void bfly4_fwd(double2& a, double2& b, double2& c, double2& d) {
  const double2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), bd = csub(b, d);
  const double2 t3 = make_double2(bd.y, -bd.x);  // -i (b - d)
  a = cadd(t0, t2);
  b = cadd(t1, t3);
  c = csub(t0, t2);
  d = csub(t1, t3);
}
__device__ __forceinline__ void bfly4_inv(double2& y0, double2& y1, double2& y2, double2& y3) {
  const double2 u0 = cadd(y0, y2), u1 = csub(y0, y2), u2 = cadd(y1, y3), yd = csub(y1, y3);
  const double2 u3 = make_double2(-yd.y, yd.x);  // +i (y1 - y3)
  y0 = cadd(u0, u2);
  y1 = cadd(u1, u3);
  y2 = csub(u0, u2);
  y3 = csub(u1, u3);
}

// Twiddles of one radix-16 step from the single table entry u = W_M^(j M/Nb):
//   level 1 (block Nb, position j + b Nb/16):  W^(m (j + b Nb/16) M/Nb) = u^m * W16^(m b)
//   level 2 (block Nb/4, position j):          W^(n j 4M/Nb)            = (u^4)^n
struct Tw16 {
  double2 u[3];   // u, u^2, u^3
  double2 v[3];   // u^4, u^8, u^12
  __device__ __forceinline__ void init(double2 w) {
    u[0] = w;
    u[1] = cmul(w, w);
    u[2] = cmul(u[1], w);
    v[0] = cmul(u[1], u[1]);
    v[1] = cmul(v[0], v[0]);
    v[2] = cmul(v[1], v[0]);
  }
  // u^m * exp(-2 pi i m b / 16), m in 1..3, b in 0..3 (compile-time after unrolling)
  __device__ __forceinline__ double2 level1(int m, int b) const {
    const double C1 = 0.92387953251128674, S1 = 0.38268343236508977, C2 = 0.70710678118654752;
    const int k = m * b;  // 0,1,2,3,4,6,9
    const double2 w = u[m - 1];
    switch (k) {
      case 0: return w;
      case 1: return cmul(w, make_double2(C1, -S1));
      case 2: return cmul(w, make_double2(C2, -C2));
      case 3: return cmul(w, make_double2(S1, -C1));
      case 4: return make_double2(w.y, -w.x);                 // * (-i)
      case 6: return cmul(w, make_double2(-C2, -C2));
      default: return cmul(w, make_double2(-C1, S1));         // k = 9
    }
  }
};

// Radix-4 passes are executed two at a time in registers (a radix-16 step: 16 loads, two levels
// of butterflies, 16 stores) - the arithmetic is exactly that of two consecutive radix-4 passes,
// with half the shared-memory traffic and barriers.  Per-pass twiddle tables: the pass with
// quarter size q reads W^(m j stride), m = 1..3, at twp[M - 4q + (m-1) q + j], contiguous in j.
template <class In>
__device__ void fft_forward(double2* s, int M, int logM, const double2* __restrict__ twp, In in,
                            const double2* __restrict__ post) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  int Nb = M;
  bool first = true;
  for (; Nb >= 16; Nb >>= 4, first = false) {
    const int q = Nb >> 2, qq = Nb >> 4;
    const bool last = post != nullptr && qq == 1;
    for (int i = tid; i < (M >> 4); i += nthr) {
      const int j = i & (qq - 1);
      const int base = (i - j) * 16 + j;
      double2 x[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int idx = base + a * q + b * qq;
          x[a][b] = first ? in(idx) : s[sp(idx)];
        }
      // All 15 twiddles of the step derive from ONE table entry u = W^(j stride): level 1 needs
      // u^m W16^(m b), level 2 (u^4)^n.  (The 128 KB table does not fit next to the 128 KB
      // transform in L1/shared memory; loading every twiddle made the kernel L2-bound.)
      Tw16 tw;
      tw.init(__ldg(twp + (M - 4 * q) + j));
      // level 1: blocks of Nb, butterflies over a, position in the quarter = j + b qq
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        bfly4_fwd(x[0][b], x[1][b], x[2][b], x[3][b]);
        x[1][b] = cmul(x[1][b], tw.level1(1, b));
        x[2][b] = cmul(x[2][b], tw.level1(2, b));
        x[3][b] = cmul(x[3][b], tw.level1(3, b));
      }
      // level 2: blocks of Nb/4 (one per a), butterflies over b, position j
      const double2 w1 = tw.v[0], w2 = tw.v[1], w3 = tw.v[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        bfly4_fwd(x[a][0], x[a][1], x[a][2], x[a][3]);
        x[a][1] = cmul(x[a][1], w1);
        x[a][2] = cmul(x[a][2], w2);
        x[a][3] = cmul(x[a][3], w3);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int idx = base + a * q + b * qq;
          s[sp(idx)] = last ? cmul(x[a][b], __ldg(post + idx)) : x[a][b];
        }
      }
    }
    __syncthreads();
  }
  for (; Nb >= 4; Nb >>= 2, first = false) {
    const int q = Nb >> 2;
    const bool last = post != nullptr && q == 1;
    for (int i = tid; i < (M >> 2); i += nthr) {
      const int j = i & (q - 1);
      const int base = (i - j) * 4 + j;
      double2 a, b, c, d;
      if (first) {
        a = in(base); b = in(base + q); c = in(base + 2 * q); d = in(base + 3 * q);
      } else {
        a = s[sp(base)]; b = s[sp(base + q)]; c = s[sp(base + 2 * q)]; d = s[sp(base + 3 * q)];
      }
      bfly4_fwd(a, b, c, d);
      const double2* tp = twp + (M - 4 * q) + j;
      b = cmul(b, __ldg(tp));
      c = cmul(c, __ldg(tp + q));
      d = cmul(d, __ldg(tp + 2 * q));
      if (last) {
        a = cmul(a, __ldg(post + base));
        b = cmul(b, __ldg(post + base + 1));
        c = cmul(c, __ldg(post + base + 2));
        d = cmul(d, __ldg(post + base + 3));
      }
      s[sp(base)] = a;
      s[sp(base + q)] = b;
      s[sp(base + 2 * q)] = c;
      s[sp(base + 3 * q)] = d;
    }
    __syncthreads();
  }
  if (Nb == 2) {
    for (int i = tid; i < (M >> 1); i += nthr) {
      double2 a = first ? in(2 * i) : s[sp(2 * i)], b = first ? in(2 * i + 1) : s[sp(2 * i + 1)];
      double2 y0 = cadd(a, b), y1 = csub(a, b);
      if (post) {
        y0 = cmul(y0, __ldg(post + 2 * i));
        y1 = cmul(y1, __ldg(post + 2 * i + 1));
      }
      s[sp(2 * i)] = y0;
      s[sp(2 * i + 1)] = y1;
    }
    __syncthreads();
  }
}

// exact inverse of fft_forward up to the factor M: permuted order in, natural order out.
// The outputs of the LAST pass go to `out(idx, value)` instead of shared memory (fuses the
// caller's post-multiply / store pass).
template <class Out>
__device__ void fft_inverse(double2* s, int M, int logM, const double2* __restrict__ twp, Out out) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  int D = 1;  // size of the blocks already transformed
  if (logM & 1) {
    const bool last = M == 2;
    for (int i = tid; i < (M >> 1); i += nthr) {
      const double2 a = s[sp(2 * i)], b = s[sp(2 * i + 1)];
      if (last) { out(2 * i, cadd(a, b)); out(2 * i + 1, csub(a, b)); }
      else { s[sp(2 * i)] = cadd(a, b); s[sp(2 * i + 1)] = csub(a, b); }
    }
    __syncthreads();
    D = 2;
  }
  if ((logM & 3) >= 2) {  // one single radix-4 pass mirrors the forward transform's remainder
    const int q = D, Nb = 4 * D;
    const bool last = Nb == M;
    for (int i = tid; i < (M >> 2); i += nthr) {
      const int j = i & (q - 1);
      const int base = (i - j) * 4 + j;
      const double2* tp = twp + (M - 4 * q) + j;
      double2 y0 = s[sp(base)];
      double2 y1 = cmulc(s[sp(base + q)], __ldg(tp));
      double2 y2 = cmulc(s[sp(base + 2 * q)], __ldg(tp + q));
      double2 y3 = cmulc(s[sp(base + 3 * q)], __ldg(tp + 2 * q));
      bfly4_inv(y0, y1, y2, y3);
      if (last) {
        out(base, y0); out(base + q, y1); out(base + 2 * q, y2); out(base + 3 * q, y3);
      } else {
        s[sp(base)] = y0; s[sp(base + q)] = y1; s[sp(base + 2 * q)] = y2; s[sp(base + 3 * q)] = y3;
      }
    }
    __syncthreads();
    D = Nb;
  }
  for (; D < M; D <<= 4) {  // fused pairs: blocks of D -> 4D -> 16D
    const int qq = D, q = 4 * D;
    const bool last = 16 * D == M;
    for (int i = tid; i < (M >> 4); i += nthr) {
      const int j = i & (qq - 1);
      const int base = (i - j) * 16 + j;
      double2 x[4][4];
      Tw16 tw;
      tw.init(__ldg(twp + (M - 4 * q) + j));
      const double2 w1 = tw.v[0], w2 = tw.v[1], w3 = tw.v[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ib = base + a * q;
        x[a][0] = s[sp(ib)];
        x[a][1] = cmulc(s[sp(ib + qq)], w1);
        x[a][2] = cmulc(s[sp(ib + 2 * qq)], w2);
        x[a][3] = cmulc(s[sp(ib + 3 * qq)], w3);
        bfly4_inv(x[a][0], x[a][1], x[a][2], x[a][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        x[1][b] = cmulc(x[1][b], tw.level1(1, b));
        x[2][b] = cmulc(x[2][b], tw.level1(2, b));
        x[3][b] = cmulc(x[3][b], tw.level1(3, b));
        bfly4_inv(x[0][b], x[1][b], x[2][b], x[3][b]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int idx = base + a * q + b * qq;
          if (last) out(idx, x[a][b]);
          else s[sp(idx)] = x[a][b];
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_filter_spectrum(const double2* __restrict__ bfilt, double2* __restrict__ Bhat, int M, int logM,
                  const double2* __restrict__ tw, double2* scratch) {
  extern __shared__ double2 sbuf[];
  double2* s = scratch ? scratch : sbuf;
  for (int i = threadIdx.x; i < M; i += blockDim.x) s[sp(i)] = bfilt[i];
  __syncthreads();
  fft_forward(s, M, logM, tw, SmemInput{s}, nullptr);
  for (int i = threadIdx.x; i < M; i += blockDim.x) Bhat[i] = s[sp(i)];
}

// chirp-z input: alpha_k Yhat[k] chi[k] for k < nfo, zero above (numpy c2r semantics for DC/Nyquist)
struct CztInput {
  const double2* yin;
  const double2* chi;
  const double* alpha;
  int nfo;
  __device__ __forceinline__ double2 operator()(int k) const {
    if (k >= nfo) return make_double2(0.0, 0.0);
    double2 a = yin[k];
    const double al = __ldg(alpha + k);
    if (al == 1.0) a.y = 0.0;  // DC / Nyquist: imaginary part ignored
    a.x *= al;
    a.y *= al;
    return cmul(a, __ldg(chi + k));
  }
};
// chirp-z output: y[m] = Re(chi[m] z[m]) * scale for m < n
struct CztOutput {
  double* yo;
  const double2* chi;
  int n;
  double scale;
  __device__ __forceinline__ void operator()(int m, double2 v) const {
    if (m < n) yo[m] = cmul(__ldg(chi + m), v).x * scale;
  }
};

__global__ void __launch_bounds__(CZT_THREADS)
k_irfft_czt(const double2* __restrict__ yhat, long long npix, int n, int nfo, int M, int logM,
            double scale, const double2* __restrict__ chi, const double* __restrict__ alpha,
            const double2* __restrict__ tw, const double2* __restrict__ Bhat,
            double* __restrict__ out, double2* scratch) {
  extern __shared__ double2 sbuf[];
  double2* s = scratch ? scratch + (size_t)blockIdx.x * M : sbuf;
  for (long long x = blockIdx.x; x < npix; x += gridDim.x) {
    // load + chirp pre-multiply fused into the first forward pass, the filter multiply into the
    // last forward pass, the chirp post-multiply + store into the last inverse pass
    fft_forward(s, M, logM, tw, CztInput{yhat + (size_t)x * nfo, chi, alpha, nfo}, Bhat);
    fft_inverse(s, M, logM, tw, CztOutput{out + (size_t)x * n, chi, n, scale});
  }
}

// ----------------------------------------------------------------------------
// "Halves" variant for M = 2 * 16^p (M = 8192 at the headline n_fft = 4907; also 512 and 32).
// The chirp-z input occupies k < nfo <= M/2, so the first radix-2 DIF stage is free:
//   half 0 = a[k],  half 1 = a[k] W_M^k   (folded into the input tables pre[0], pre[1]),
// and the two H = M/2 point transforms are pure radix-16: p passes each way.  The LAST forward
// pass and the FIRST inverse pass work on the same 16 contiguous elements, so they are fused in
// registers around the filter multiply, and the radix-2 stage that rejoins the halves is fused
// into the output pass (tables post[0], post[1] carry the chirp, the sign and conj W_M^m).
// Per pixel: 2p passes + the output sweep instead of 2p + 2 + ..., 10 instead of 14 array sweeps
// through shared memory at M = 8192, 6 CTA barriers instead of 8, and the per-pass twiddles
// (273 values) live in shared memory instead of behind an L2 round trip at the head of every pass.
// ----------------------------------------------------------------------------
__device__ __forceinline__ double2 w16_mul(double2 x, int k) {   // x * exp(-2 pi i k / 16)
  const double C1 = 0.92387953251128674, S1 = 0.38268343236508977, C2 = 0.70710678118654752;
  switch (k) {
    case 0: return x;
    case 1: return cmul(x, make_double2(C1, -S1));
    case 2: return make_double2(C2 * (x.x + x.y), C2 * (x.y - x.x));
    case 3: return cmul(x, make_double2(S1, -C1));
    case 4: return make_double2(x.y, -x.x);
    case 6: return make_double2(C2 * (x.y - x.x), -C2 * (x.x + x.y));
    default: return cmul(x, make_double2(-C1, S1));   // k = 9
  }
}
__device__ __forceinline__ double2 w16_mulc(double2 x, int k) {  // x * exp(+2 pi i k / 16)
  const double C1 = 0.92387953251128674, S1 = 0.38268343236508977, C2 = 0.70710678118654752;
  switch (k) {
    case 0: return x;
    case 1: return cmulc(x, make_double2(C1, -S1));
    case 2: return make_double2(C2 * (x.x - x.y), C2 * (x.x + x.y));
    case 3: return cmulc(x, make_double2(S1, -C1));
    case 4: return make_double2(-x.y, x.x);
    case 6: return make_double2(-C2 * (x.x + x.y), C2 * (x.x - x.y));
    default: return cmulc(x, make_double2(-C1, S1));
  }
}

// index padding of the halves kernels: one empty slot after every 16 elements - conflict-free for
// all their access patterns (contiguous runs of >= 8 elements, or one 16-element block per lane)
__device__ __forceinline__ int sq(int i) { return i + (i >> 4); }
__host__ __device__ constexpr size_t sq_elems(size_t n) { return n + (n >> 4); }

// one forward radix-16 item with block twiddle u (level 1: u^m W16^(m b), level 2: (u^4)^n)
__device__ __forceinline__ void r16_fwd(double2 (&x)[4][4], double2 u) {
  Tw16 tw;
  tw.init(u);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    bfly4_fwd(x[0][b], x[1][b], x[2][b], x[3][b]);
    x[1][b] = cmul(x[1][b], tw.level1(1, b));
    x[2][b] = cmul(x[2][b], tw.level1(2, b));
    x[3][b] = cmul(x[3][b], tw.level1(3, b));
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    bfly4_fwd(x[a][0], x[a][1], x[a][2], x[a][3]);
    x[a][1] = cmul(x[a][1], tw.v[0]);
    x[a][2] = cmul(x[a][2], tw.v[1]);
    x[a][3] = cmul(x[a][3], tw.v[2]);
  }
}
__device__ __forceinline__ void r16_inv(double2 (&x)[4][4], double2 u) {
  Tw16 tw;
  tw.init(u);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    x[a][1] = cmulc(x[a][1], tw.v[0]);
    x[a][2] = cmulc(x[a][2], tw.v[1]);
    x[a][3] = cmulc(x[a][3], tw.v[2]);
    bfly4_inv(x[a][0], x[a][1], x[a][2], x[a][3]);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    x[1][b] = cmulc(x[1][b], tw.level1(1, b));
    x[2][b] = cmulc(x[2][b], tw.level1(2, b));
    x[3][b] = cmulc(x[3][b], tw.level1(3, b));
    bfly4_inv(x[0][b], x[1][b], x[2][b], x[3][b]);
  }
}
// the same with u = 1 (blocks of 16 contiguous elements): only the W16 constants remain
__device__ __forceinline__ void r16_fwd_unit(double2 (&x)[4][4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    bfly4_fwd(x[0][b], x[1][b], x[2][b], x[3][b]);
#pragma unroll
    for (int m = 1; m < 4; ++m) x[m][b] = w16_mul(x[m][b], m * b);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) bfly4_fwd(x[a][0], x[a][1], x[a][2], x[a][3]);
}
// shared-memory twiddles of the halves transforms: htw = for Nb = H, H/16, ..., 256 the Nb/16 values
// exp(-2 pi i j / Nb) (the u of a radix-16 pass with block Nb), concatenated (CztTables::htw)
__device__ __forceinline__ void halves_stage_twiddles(double2* stw, const double2* __restrict__ htw, int M) {
  int off = 0;
  for (int Nb = M >> 1; Nb > 16; Nb >>= 4) off += Nb >> 4;
  for (int j = threadIdx.x; j < off; j += blockDim.x) stw[j] = htw[j];
}
__host__ __device__ constexpr size_t halves_twiddle_elems(size_t M) { return M / 30 + 16; }   // >= sum of Nb/16

// forward passes Nb = H ... 256 of both halves (in place in s, first pass fed by `in(h, k)`);
// returns the twiddle offset after the last pass.  The Nb = 16 pass is left to the caller.
template <class In>
__device__ __forceinline__ int halves_forward(double2* s, const double2* stw, int M, int logH, In in,
                                              bool& first, bool skip_first_pass = false) {
  const int H = M >> 1;
  int toff = 0;
  int Nb = H;
  if (skip_first_pass && H > 16) {     // the caller has done the Nb = H pass itself
    toff = H >> 4;
    Nb = H >> 4;
    first = false;
  }
  for (; Nb > 16; Nb >>= 4) {
    const int q = Nb >> 2, qq = Nb >> 4;
    for (int i = threadIdx.x; i < (M >> 4); i += blockDim.x) {
      const int h = i >> (logH - 4), li = i & ((H >> 4) - 1);
      const int j = li & (qq - 1);
      const int lbase = (li - j) * 16 + j;
      double2 x[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = lbase + a * q + b * qq;
          x[a][b] = first ? in(h, k) : s[sq(h * H + k)];
        }
      r16_fwd(x, stw[toff + j]);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[sq(h * H + lbase + a * q + b * qq)] = x[a][b];
    }
    __syncthreads();
    toff += qq;
    first = false;
  }
  return toff;
}

struct FilterHalvesInput {     // genuine radix-2 DIF stage of a full-length input
  const double2* b;
  const double2* twM;          // W_M^k
  int H;
  __device__ __forceinline__ double2 operator()(int h, int k) const {
    const double2 lo = b[k], hi = b[k + H];
    return h == 0 ? cadd(lo, hi) : cmul(csub(lo, hi), __ldg(twM + k));
  }
};
struct CztHalvesInput {        // pre[h][k] = alpha_k chi_k (W_M^k)^h; DC / Nyquist use the real part only
  const double2* yin;
  const double2* pre;
  int nfo, nyq;
  __device__ __forceinline__ double2 operator()(int h, int k) const {
    if (k >= nfo) return make_double2(0.0, 0.0);
    double2 a = yin[k];
    if (k == 0 || k == nyq) a.y = 0.0;
    return cmul(a, __ldg(pre + (size_t)h * nfo + k));
  }
};

__global__ void __launch_bounds__(512)
k_filter_spectrum_h(const double2* __restrict__ bfilt, double2* __restrict__ Bhat, int M, int logM,
                    const double2* __restrict__ tw, const double2* __restrict__ htw) {
  extern __shared__ double2 sbuf[];
  double2* s = sbuf;
  double2* stw = sbuf + sq_elems(M);
  const int H = M >> 1, logH = logM - 1;
  halves_stage_twiddles(stw, htw, M);
  __syncthreads();
  FilterHalvesInput in{bfilt, tw + M, H};
  bool first = true;
  halves_forward(s, stw, M, logH, in, first);
  for (int i = threadIdx.x; i < (M >> 4); i += blockDim.x) {
    const int base = i * 16;                   // half h = base / H, contiguous block of 16
    double2 x[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int idx = base + a * 4 + b;
        x[a][b] = first ? in(idx >= H, idx & (H - 1)) : s[sq(idx)];
      }
    r16_fwd_unit(x);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) Bhat[base + a * 4 + b] = x[a][b];
  }
}

// a load the compiler may not sink past a barrier (operands requested ahead of their use)
__device__ __forceinline__ double2 ldg_pinned(const double2* p) {
#ifdef PKM_HOST_EMULATION            // tests/host_emulation compiles this file's device code for the CPU
  return *p;
#endif
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");  // not .nc: ptxas moves non-coherent loads across BAR.SYNC
  return v;
}

// One thread owns one radix-16 item per pass (blockDim = M/16 <= 512).  Global operands never
// stall a pass (the kernel was bound by L2 round trips, ncu long-scoreboard): the NEXT pixel's
// spectrum is fetched into registers under the output sweep, the filter rows 0-1 and the first
// output weights are requested before the barrier that precedes their use.  (Staging the input
// weights `pre` in shared memory as well measured slower, 1.74 vs 1.67 ms on 14k pixels: the 78 KB
// come out of the L1 that serves the filter and the output weights.)
template <int LOGH>
__global__ void __launch_bounds__(512)
k_irfft_czt_h(const double2* __restrict__ yhat, long long npix, int n, int nfo, int /*M*/, int /*logM*/,
              double scale, const double2* __restrict__ pre, const double2* __restrict__ post,
              const double2* __restrict__ htw, const double2* __restrict__ Bhat,
              double* __restrict__ out) {
  extern __shared__ double2 sbuf[];
  double2* s = sbuf;
  double2* stw = sbuf + sq_elems(2 << LOGH);
  constexpr int logH = LOGH, H = 1 << LOGH, M = 2 * H;    // compile-time: M = 32, 512 or 8192
  halves_stage_twiddles(stw, htw, M);
  __syncthreads();
  const int nyq = (n & 1) ? -1 : nfo - 1;
  constexpr int q = H >> 2, qq = H >> 4;                  // first pass (Nb = H)
  const int item = threadIdx.x;                           // this thread's item in every pass
  const bool has_item = item < (M >> 4);
  const int h0 = item >> (logH - 4), j0 = item & (qq - 1);
  // raw spectrum values of the first pass: rows 0..2 (row 3 starts at 3H/4 >= nfo: always zero)
  double2 raw[3][4];
  auto fetch_raw = [&](long long px) {
    const double2* yin = yhat + (size_t)px * nfo;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) raw[a][b] = ldg_pinned(yin + min(j0 + a * q + b * qq, nfo - 1));
  };
  fetch_raw(min((long long)blockIdx.x, npix - 1));
  for (long long px = blockIdx.x; px < npix; px += gridDim.x) {
    CztHalvesInput in{yhat + (size_t)px * nfo, pre, nfo, nyq};
    bool first = true;
    double2 f[2][4];                                       // filter rows, two at a time
    auto request_filter = [&](int r0) {                    // rows r0, r0 + 1 of this thread's block
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) f[a][b] = ldg_pinned(Bhat + item * 16 + (r0 + a) * 4 + b);
    };
    if (H > 16) {
      if (has_item) {
        double2 x[4][4];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int k = j0 + a * q + b * qq, kk = min(k, nfo - 1);
            double2 v = raw[a][b];
            v.y = (k == 0 || k == nyq) ? 0.0 : v.y;
            v = cmul(v, __ldg(pre + (size_t)h0 * nfo + kk));
            x[a][b] = k < nfo ? v : make_double2(0.0, 0.0);
          }
#pragma unroll
        for (int b = 0; b < 4; ++b) x[3][b] = make_double2(0.0, 0.0);
        r16_fwd(x, stw[j0]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) s[sq(h0 * H + j0 + a * q + b * qq)] = x[a][b];
        if (H == 256) request_filter(0);                   // the next pass is the fused middle pass
      }
      __syncthreads();
      first = false;
    }
    // remaining forward passes Nb = H/16 ... 256
    int toff = H > 16 ? qq : 0;
#pragma unroll
    for (int Nb = H >> 4; Nb > 16; Nb >>= 4) {
      const int q2 = Nb >> 2, qq2 = Nb >> 4;
      if (has_item) {
        const int li = item & ((H >> 4) - 1), j = li & (qq2 - 1);
        const int base = h0 * H + (li - j) * 16 + j;
        double2 x[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) x[a][b] = s[sq(base + a * q2 + b * qq2)];
        r16_fwd(x, stw[toff + j]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) s[sq(base + a * q2 + b * qq2)] = x[a][b];
        if (Nb == 256) request_filter(0);
      }
      __syncthreads();
      toff += qq2;
    }
    // last forward pass (blocks of 16), filter, first inverse pass: all on the same 16 elements.
    if (has_item) {
      const int base = item * 16;
      double2 x[4][4];
      if (first) request_filter(0);                        // M = 32: nothing precedes this pass
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int idx = base + a * 4 + b;
          x[a][b] = first ? in(idx >= H, idx & (H - 1)) : s[sq(idx)];
        }
      // forward level 1 (over a), then per row: forward level 2, filter, inverse level 1
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        bfly4_fwd(x[0][b], x[1][b], x[2][b], x[3][b]);
#pragma unroll
        for (int m = 1; m < 4; ++m) x[m][b] = w16_mul(x[m][b], m * b);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        bfly4_fwd(x[a][0], x[a][1], x[a][2], x[a][3]);
#pragma unroll
        for (int b = 0; b < 4; ++b) x[a][b] = cmul(x[a][b], f[a & 1][b]);
        if (a < 2) {
#pragma unroll
          for (int b = 0; b < 4; ++b) f[a][b] = __ldg(Bhat + base + (a + 2) * 4 + b);
        }
        bfly4_inv(x[a][0], x[a][1], x[a][2], x[a][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int m = 1; m < 4; ++m) x[m][b] = w16_mulc(x[m][b], m * b);
        bfly4_inv(x[0][b], x[1][b], x[2][b], x[3][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[sq(base + a * 4 + b)] = x[a][b];
    }
    // output weights of the first batch: requested before the barrier(s) that precede the sweep
    constexpr int OB = 5;
    double2 p0[OB], p1[OB];
    auto request_weights = [&](int m0) {
#pragma unroll
      for (int r = 0; r < OB; ++r) {
        const int m = min(m0 + r * (int)blockDim.x, n - 1);
        p0[r] = ldg_pinned(post + m);
        p1[r] = ldg_pinned(post + n + m);
      }
    };
    if (H == 16) request_weights(threadIdx.x);
    __syncthreads();
    // inverse passes: blocks of D -> 16 D, D = 16 ... H/16
#pragma unroll
    for (int D = 16; 16 * D <= H; D <<= 4) {
      const int qq2 = D, q2 = 4 * D;
      toff -= qq2;
      if (has_item) {
        const int li = item & ((H >> 4) - 1), j = li & (qq2 - 1);
        const int base = h0 * H + (li - j) * 16 + j;
        double2 x[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) x[a][b] = s[sq(base + a * q2 + b * qq2)];
        r16_inv(x, stw[toff + j]);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) s[sq(base + a * q2 + b * qq2)] = x[a][b];
      }
      if (16 * D == H) request_weights(threadIdx.x);
      __syncthreads();
    }
    // the next pixel's spectrum travels under the output sweep
    fetch_raw(min(px + (long long)gridDim.x, npix - 1));   // unconditional: `raw` is dead between its use and here
    // output sweep: z[m] = z0[i] +- conj(W_M^i) z1[i] (i = m mod H), y[m] = Re(chi[m] z[m]) * scale
    //             = Re(post0[m] z0[i] + post1[m] z1[i]) * scale
    double* yo = out + (size_t)px * n;
    for (int m0 = threadIdx.x; m0 < n; m0 += OB * blockDim.x) {
      if (m0 != (int)threadIdx.x) request_weights(m0);
#pragma unroll
      for (int r = 0; r < OB; ++r) {
        const int m = m0 + r * (int)blockDim.x;
        const int i = min(m, n - 1) & (H - 1);
        const double2 z0 = s[sq(i)], z1 = s[sq(H + i)];
        const double y = (fma(p0[r].x, z0.x, -p0[r].y * z0.y) + fma(p1[r].x, z1.x, -p1[r].y * z1.y)) * scale;
        if (m < n) yo[m] = y;
      }
    }
    __syncthreads();   // the next pixel's first pass overwrites s
  }
}

// ----------------------------------------------------------------------------
// four-step variant for M = R * C > shared memory (C = 4096 in smem, R <= 16 rows in a CTA-private,
// L2-resident global scratch g[R][C]):
//   forward: column DFTs of length R (direct, R^2 complex MACs per column) with the twiddle
//            W_M^(n2 k1), then R shared-memory transforms of length C (output row k1 holds the
//            frequencies k1 + R k2 in the digit-permuted order of fft_forward);
//   inverse: the exact reverse.  4 passes over the scratch instead of ~2 log4(M).
// ----------------------------------------------------------------------------
struct RowInput {
  const double2* row;
  __device__ __forceinline__ double2 operator()(int i) const { return row[i]; }
};
struct RowOutput {
  double2* row;
  __device__ __forceinline__ void operator()(int i, double2 v) const { row[i] = v; }
};

template <int R, class In>
__device__ void fft4_forward(double2* g, double2* s, int C, int logC, const double2* __restrict__ tw,
                             In in, const double2* __restrict__ post) {
  constexpr int CZT4_MAXR = R;
  const double2* twM = tw + C;       // W_M^(n2)
  const double2* twR = tw + 2 * C;   // W_R^(j)
  for (int n2 = threadIdx.x; n2 < C; n2 += blockDim.x) {
    double2 x[CZT4_MAXR];
#pragma unroll
    for (int n1 = 0; n1 < CZT4_MAXR; ++n1)
      if (n1 < R) x[n1] = in(n1 * C + n2);
    const double2 w1 = __ldg(twM + n2);
    double2 wk = make_double2(1.0, 0.0);            // W_M^(n2 k1)
    for (int k1 = 0; k1 < R; ++k1) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int n1 = 0; n1 < CZT4_MAXR; ++n1)
        if (n1 < R) {
          const double2 w = __ldg(twR + ((n1 * k1) & (R - 1)));
          acc.x = fma(x[n1].x, w.x, fma(-x[n1].y, w.y, acc.x));
          acc.y = fma(x[n1].x, w.y, fma(x[n1].y, w.x, acc.y));
        }
      g[(size_t)k1 * C + n2] = cmul(acc, wk);
      wk = cmul(wk, w1);
    }
  }
  __syncthreads();
  for (int k1 = 0; k1 < R; ++k1) {
    fft_forward(s, C, logC, tw, RowInput{g + (size_t)k1 * C}, post ? post + (size_t)k1 * C : nullptr);
    for (int i = threadIdx.x; i < C; i += blockDim.x) g[(size_t)k1 * C + i] = s[sp(i)];
    __syncthreads();
  }
}

template <int R, class Out>
__device__ void fft4_inverse(double2* g, double2* s, int C, int logC, const double2* __restrict__ tw,
                             Out out) {
  constexpr int CZT4_MAXR = R;
  const double2* twM = tw + C;
  const double2* twR = tw + 2 * C;
  for (int k1 = 0; k1 < R; ++k1) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s[sp(i)] = g[(size_t)k1 * C + i];
    __syncthreads();
    fft_inverse(s, C, logC, tw, RowOutput{g + (size_t)k1 * C});
  }
  for (int n2 = threadIdx.x; n2 < C; n2 += blockDim.x) {
    double2 y[CZT4_MAXR];
    const double2 w1 = __ldg(twM + n2);
    double2 wk = make_double2(1.0, 0.0);
#pragma unroll
    for (int k1 = 0; k1 < CZT4_MAXR; ++k1)
      if (k1 < R) {
        y[k1] = cmulc(g[(size_t)k1 * C + n2], wk);   // undo W_M^(n2 k1)
        wk = cmul(wk, w1);
      }
    for (int n1 = 0; n1 < R; ++n1) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int k1 = 0; k1 < CZT4_MAXR; ++k1)
        if (k1 < R) {
          const double2 w = __ldg(twR + ((n1 * k1) & (R - 1)));   // conjugate applied below
          acc.x = fma(y[k1].x, w.x, fma(y[k1].y, w.y, acc.x));
          acc.y = fma(y[k1].y, w.x, fma(-y[k1].x, w.y, acc.y));
        }
      out(n1 * C + n2, acc);
    }
  }
  __syncthreads();
}

template <int R>
__global__ void __launch_bounds__(512)
k_filter_spectrum4(const double2* __restrict__ bfilt, double2* __restrict__ Bhat, int C, int logC,
                   const double2* __restrict__ tw) {
  extern __shared__ double2 sbuf[];
  fft4_forward<R>(Bhat, sbuf, C, logC, tw, RowInput{bfilt}, nullptr);   // Bhat doubles as the scratch
}

template <int R>
__global__ void __launch_bounds__(256, 2)
k_irfft_czt4(const double2* __restrict__ yhat, long long npix, int n, int nfo, int C, int logC,
             double scale, const double2* __restrict__ chi, const double* __restrict__ alpha,
             const double2* __restrict__ tw, const double2* __restrict__ Bhat,
             double* __restrict__ out, double2* scratch) {
  extern __shared__ double2 sbuf[];
  double2* g = scratch + (size_t)blockIdx.x * R * C;
  for (long long x = blockIdx.x; x < npix; x += gridDim.x) {
    fft4_forward<R>(g, sbuf, C, logC, tw, CztInput{yhat + (size_t)x * nfo, chi, alpha, nfo}, Bhat);
    fft4_inverse<R>(g, sbuf, C, logC, tw, CztOutput{out + (size_t)x * n, chi, n, scale});
  }
}

// O(n * nfo) reference transform with exact integer phase reduction (tests only)
__global__ void __launch_bounds__(256)
k_irfft_naive(const double2* __restrict__ yhat, int n, int nfo, double scale,
              const double* __restrict__ alpha, const double2* __restrict__ wn,
              double* __restrict__ out) {
  const long long x = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const double2* yin = yhat + (size_t)x * nfo;
  double acc = 0.0;
  int r = 0;  // (k*m) mod n
  for (int k = 0; k < nfo; ++k) {
    double2 a = yin[k];
    const double al = alpha[k];
    if (al == 1.0) a.y = 0.0;
    double2 w = __ldg(wn + r);  // e^{+2 pi i r / n}
    acc += al * (a.x * w.x - a.y * w.y);
    r += m;
    if (r >= n) r -= n;
  }
  out[(size_t)x * n + m] = acc * scale;
}

__global__ void k_transpose(const double* __restrict__ src, long long npix, long long n,
                            double* __restrict__ dst, long long row_stride, long long col0) {
  __shared__ double tile[32][33];
  const long long m0 = (long long)blockIdx.x * 32, x0 = (long long)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    long long x = x0 + r, m = m0 + threadIdx.x;
    if (x < npix && m < n) tile[r][threadIdx.x] = src[x * n + m];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    long long m = m0 + r, x = x0 + threadIdx.x;
    if (x < npix && m < n) dst[m * row_stride + col0 + x] = tile[threadIdx.x][r];
  }
}

__global__ void k_axpy(int n_cmps, const double* const* __restrict__ ptrs,
                       const double* __restrict__ w, long long n, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < n_cmps; ++c) acc += w[c] * ptrs[c][i];  // same order as mixture.py:45-47
    out[i] = acc;
  }
}

// out[i] = log sum_c exp(log_w[c] + ptrs[c][i]): the component-collapsed log density of a Mixture
// (scipy.special.logsumexp over the component axis, mixture.py:50-80), max-shifted
__global__ void k_logsumexp_cubes(int n_cmps, const double* const* __restrict__ ptrs,
                                  const double* __restrict__ log_w, long long n, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double m = -INFINITY;
    for (int c = 0; c < n_cmps; ++c) m = fmax(m, log_w[c] + ptrs[c][i]);
    double acc = 0.0;
    if (m > -INFINITY)
      for (int c = 0; c < n_cmps; ++c) acc += exp(log_w[c] + ptrs[c][i] - m);
    out[i] = m > -INFINITY ? m + log(acc) : -INFINITY;
  }
}

}  // namespace

template <int R>
static void filter4_t(pkm_plan* pl, const double2* d_b) {
  const CztTables& c = pl->czt;
  const size_t smem = sizeof(double2) * sp_elems(c.C);
  PKM_CUDA(cudaFuncSetAttribute(k_filter_spectrum4<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_filter_spectrum4<R><<<1, 512, smem>>>(d_b, pl->d_Bhat, c.C, c.logC, pl->d_twiddle);
}

void czt_prepare_filter(pkm_plan* pl) {
  const CztTables& c = pl->czt;
  double2* d_b = nullptr;
  PKM_CUDA(cudaMalloc(&d_b, sizeof(double2) * c.M));
  PKM_CUDA(cudaMemcpy(d_b, c.bfilt.data(), sizeof(double2) * c.M, cudaMemcpyHostToDevice));
  if (c.R == 1) {
    const size_t smem = sizeof(double2) * sp_elems(c.M);
    PKM_CUDA(cudaFuncSetAttribute(k_filter_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_filter_spectrum<<<1, 256, smem>>>(d_b, pl->d_Bhat, c.M, c.logM, pl->d_twiddle, nullptr);
  } else {
    switch (c.R) {
      case 2: filter4_t<2>(pl, d_b); break;
      case 4: filter4_t<4>(pl, d_b); break;
      case 8: filter4_t<8>(pl, d_b); break;
      case 16: filter4_t<16>(pl, d_b); break;
      default: PKM_THROW(PKM_ERR_INVALID, "n_fft=%d too long for the inverse FFT (M=%d)", c.n, c.M);
    }
  }
  PKM_CUDA(cudaGetLastError());
  if (pl->d_Bhat_h) {
    const size_t smem = sizeof(double2) * (sq_elems(c.M) + halves_twiddle_elems(c.M));
    PKM_CUDA(cudaFuncSetAttribute(k_filter_spectrum_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_filter_spectrum_h<<<1, std::max(32, std::min(512, c.M / 16)), smem>>>(d_b, pl->d_Bhat_h, c.M, c.logM, pl->d_twiddle, pl->d_htw);
    PKM_CUDA(cudaGetLastError());
  }
  PKM_CUDA(cudaDeviceSynchronize());
  cudaFree(d_b);
}

// M = 2 * 16^p, one-piece transform: the halves kernel applies (PKM_IRFFT_HALVES=0 turns it off)
bool czt_halves_ok(const CztTables& c) {
  if (c.R != 1 || c.M < 32 || c.M > 8192 || (c.logM - 1) % 4 != 0) return false;
  const char* e = getenv("PKM_IRFFT_HALVES");
  return !(e && atoi(e) == 0);
}

template <int R>
static void irfft4_t(pkm_plan* pl, const double2* yhat, int64_t npix, double s, double* out, cudaStream_t st) {
  const CztTables& c = pl->czt;
  const size_t smem = sizeof(double2) * sp_elems(c.C);
  PKM_CUDA(cudaFuncSetAttribute(k_irfft_czt4<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(npix, (int64_t)pl->num_sms * 2);
  pl->misc.reserve(sizeof(double2) * (size_t)c.M * grid);
  k_irfft_czt4<R><<<grid, 256, smem, st>>>(yhat, npix, c.n, c.nfo, c.C, c.logC, s, pl->d_chi, pl->d_alpha,
                                          pl->d_twiddle, pl->d_Bhat, out, pl->misc.as<double2>());
}

void launch_irfft_czt(pkm_plan* pl, const double2* yhat, int64_t npix, double scale, double* out,
                      cudaStream_t st) {
  if (npix <= 0) return;
  const CztTables& c = pl->czt;
  const double s = scale / ((double)c.M * (double)c.n);
  KernelTimer tm(pl, PKM_K_IRFFT, st);
  if (pl->d_Bhat_h && czt_halves_ok(c)) {
    const size_t smem = sizeof(double2) * (sq_elems(c.M) + halves_twiddle_elems(c.M));
    auto kern = k_irfft_czt_h<12>;
    if (c.logM - 1 == 4) kern = k_irfft_czt_h<4>;
    if (c.logM - 1 == 8) kern = k_irfft_czt_h<8>;
    PKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(220 * 1024) / smem));
    const int grid = (int)std::min<int64_t>(npix, (int64_t)pl->num_sms * per_sm);
    kern<<<grid, std::max(32, std::min(512, c.M / 16)), smem, st>>>(yhat, npix, c.n, c.nfo, c.M, c.logM, s, pl->d_pre,
                                                                   pl->d_post, pl->d_htw, pl->d_Bhat_h, out);
  } else if (c.R == 1) {
    const size_t smem = sizeof(double2) * sp_elems(c.M);
    PKM_CUDA(cudaFuncSetAttribute(k_irfft_czt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(220 * 1024) / smem));
    const int grid = (int)std::min<int64_t>(npix, (int64_t)pl->num_sms * per_sm);
    k_irfft_czt<<<grid, c.M >= 4096 ? CZT_THREADS : 256, smem, st>>>(yhat, npix, c.n, c.nfo, c.M, c.logM, s, pl->d_chi,
                                                                    pl->d_alpha, pl->d_twiddle, pl->d_Bhat, out, nullptr);
  } else {
    switch (c.R) {
      case 2: irfft4_t<2>(pl, yhat, npix, s, out, st); break;
      case 4: irfft4_t<4>(pl, yhat, npix, s, out, st); break;
      case 8: irfft4_t<8>(pl, yhat, npix, s, out, st); break;
      case 16: irfft4_t<16>(pl, yhat, npix, s, out, st); break;
      default: PKM_THROW(PKM_ERR_INVALID, "n_fft=%d too long for the inverse FFT (M=%d)", c.n, c.M);
    }
  }
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}

void launch_irfft_naive(pkm_plan* pl, const double2* yhat, int64_t npix, double scale, double* out,
                        cudaStream_t st) {
  if (npix <= 0) return;
  const CztTables& c = pl->czt;
  dim3 grid((c.n + 255) / 256, (unsigned)npix);
  k_irfft_naive<<<grid, 256, 0, st>>>(yhat, c.n, c.nfo, scale / (double)c.n, pl->d_alpha, pl->d_wn, out);
  PKM_CUDA(cudaGetLastError());
  pl->launches++;
}

void launch_transpose(const double* src, int64_t npix, int64_t n, double* dst, int64_t row_stride,
                      int64_t col0, cudaStream_t st) {
  if (npix <= 0 || n <= 0) return;
  dim3 block(32, 8);
  // grid.y is limited to 65535 blocks of 32 pixels (2M pixels) - split if ever needed
  for (int64_t x0 = 0; x0 < npix; x0 += 65535LL * 32) {
    int64_t np = std::min<int64_t>(npix - x0, 65535LL * 32);
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((np + 31) / 32));
    k_transpose<<<grid, block, 0, st>>>(src + x0 * n, np, n, dst, row_stride, col0 + x0);
    PKM_CUDA(cudaGetLastError());
  }
}

void launch_axpy(int n_cmps, const double* const* d_ptrs_dev, const double* d_w, int64_t n,
                 double* out, cudaStream_t st) {
  if (n <= 0) return;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  k_axpy<<<grid, 256, 0, st>>>(n_cmps, d_ptrs_dev, d_w, n, out);
  PKM_CUDA(cudaGetLastError());
}

void launch_logsumexp_cubes(int n_cmps, const double* const* d_ptrs_dev, const double* d_log_w, int64_t n,
                            double* out, cudaStream_t st) {
  if (n <= 0) return;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  k_logsumexp_cubes<<<grid, 256, 0, st>>>(n_cmps, d_ptrs_dev, d_log_w, n, out);
  PKM_CUDA(cudaGetLastError());
}
