// Satellites of the datacube path: log-marginals of the pixelated density, moments, noise.
//
//  * pkm_reduce_logp: every Component._get_log_p_* marginal
//    (/root/reference/popkinmocks/components/base.py:884-1174) is
//    logsumexp_{dropped axes}( log p + log dV [+ log L(t,z) - log norm] ) (base.py:1017-1041).
//    One generic kernel does it in ONE streaming pass over the 5-D array (online log-sum-exp:
//    running maximum + rescaled sum per output) instead of the reference's 3+ full-array
//    temporaries per call.
//  * pkm_moments: mean and central moments 2..4 of P[nvar][ncond] along axis 0 (base.py:208-260).
//  * pkm_noise: ShotNoise / ConstantSNR sigma maps and Gaussian sampling (noise.py:29-71),
//    Philox4x32-10 counter RNG + Box-Muller in FP64.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "exp_tab.cuh"
#include "pkm_internal.h"

namespace {

__device__ __forceinline__ void atomic_max_double(double* addr, double val) {
  // IEEE-754 ordering trick: non-negative doubles order like signed ints, negative ones like
  // reversed unsigned ints.  NaN is never passed in.
  if (val >= 0.0)
    atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(val));
  else
    atomicMin(reinterpret_cast<unsigned long long*>(addr),
              static_cast<unsigned long long>(__double_as_longlong(val)));
}

struct ReduceGeom {
  int nt, nv, nz;
  long long npix;
  int kt, kv, kx, kz;  // 1 = axis kept
  double log_dxdy;     // 0 when the source already holds volume-weighted log probabilities
  int wt;              // 1 = add the log volume elements of (t, v, z) to the source (a log DENSITY)
  double out_log_dxdy; // log(dx dy): what a density OUTPUT divides by when x is kept
};

__device__ __forceinline__ long long out_index(const ReduceGeom& g, int t, int v, long long x, int z) {
  long long o = g.kt ? t : 0;
  o = o * (g.kv ? g.nv : 1) + (g.kv ? v : 0);
  o = o * (g.kx ? g.npix : 1) + (g.kx ? x : 0);
  o = o * (g.kz ? g.nz : 1) + (g.kz ? z : 0);
  return o;
}

// log-sum-exp kept as a pair (m, s): value = m + log(s), m = running maximum, s = sum of exp(a - m)
struct Lse {
  double m, s;
};
__device__ __forceinline__ Lse lse_merge(Lse a, Lse b) {
  const double M = fmax(a.m, b.m);
  if (!(M > -INFINITY)) return Lse{-INFINITY, a.s + b.s};   // 0, or NaN if either side carries one
  return Lse{M, a.s * exp(a.m - M) + b.s * exp(b.m - M)};
}

// exp(x) for -708 <= x <= 0 (or NaN) with the shared 256-entry table (exp_tab.cuh), WITHOUT a range
// test: the caller skips the term when x < -708.  NaN propagates through the polynomial and the
// exponent patch (the low word of a NaN `t` is 0).  9 FP64 + 6 integer instructions.
__device__ __forceinline__ double exp_core_tab(double x, const double* __restrict__ tab) {
  const double t = fma(x, 369.32993046757463228, 6755399441055744.0);
  const int k = __double2loint(t);
  const double fk = t - 6755399441055744.0;
  double r = fma(fk, -2.70760617331688990816e-03, x);
  r = fma(fk, -7.45396456746323320321e-13, r);
  double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  const double p = fma(r * r, q, r);
  const double T = tab[k & (PKM_EXP_TAB - 1)];
  const double v = fma(T, p, T);
  return __hiloint2double(__double2hiint(v) + ((k >> 8) << 20), __double2loint(v));
}
__device__ __forceinline__ double exp_neg_tab(double x, const double* __restrict__ tab) {
  return x < -708.0 ? 0.0 : exp_core_tab(x, tab);
}

// The two per-element tests of the streaming loop, written as PTX so that they stay what they are.
// Left to the compiler, "does any element of the batch exceed the running maximum" became a
// NaN-aware running maximum (DSETP.MAX + FSEL + SEL + LOP3 + four moves per element) and the
// conditional add a DADD with two selects: 56 instructions per element, issue bound at half the
// HBM rate (ncu, profiles/r2c_reduce_summary.md).
__device__ __forceinline__ bool any_greater8(const double (&a)[8], double m) {   // NaN compares false
  unsigned r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f64 p, %1, %9;\n\t"
      "setp.gt.or.f64 p, %2, %9, p;\n\t"
      "setp.gt.or.f64 p, %3, %9, p;\n\t"
      "setp.gt.or.f64 p, %4, %9, p;\n\t"
      "setp.gt.or.f64 p, %5, %9, p;\n\t"
      "setp.gt.or.f64 p, %6, %9, p;\n\t"
      "setp.gt.or.f64 p, %7, %9, p;\n\t"
      "setp.gt.or.f64 p, %8, %9, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(m));
  return r != 0;
}
__device__ __forceinline__ bool any_less8(const double (&a)[8], double lim) {    // NaN compares false
  unsigned r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %1, %9;\n\t"
      "setp.lt.or.f64 p, %2, %9, p;\n\t"
      "setp.lt.or.f64 p, %3, %9, p;\n\t"
      "setp.lt.or.f64 p, %4, %9, p;\n\t"
      "setp.lt.or.f64 p, %5, %9, p;\n\t"
      "setp.lt.or.f64 p, %6, %9, p;\n\t"
      "setp.lt.or.f64 p, %7, %9, p;\n\t"
      "setp.lt.or.f64 p, %8, %9, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(lim));
  return r != 0;
}
// s += y unless x < -708 (y underflows); a NaN x is unordered, passes the test and poisons the sum
__device__ __forceinline__ void add_unless_underflow(double& s, double x, double y) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.geu.f64 p, %1, 0dC086200000000000;\n\t"
      "@p add.f64 %0, %0, %2;\n\t}"
      : "+d"(s)
      : "d"(x), "d"(y));
}

// ONE streaming pass over the 5-D array ("online" log-sum-exp).
// Thread = one (pixel-in-chunk, z) element of a chunk of PB whole pixels, contiguous in memory for
// fixed (t, v).  CTA = (chunk or chunk group, kept t, block of VPC kept v's); for every kept v of
// its block the CTA runs the reduction over the dropped axes and stores one result, so the
// per-CTA set-up (tables, barriers) is shared by VPC outputs.
// x kept:    gridDim.x = number of chunks, every output cell has exactly one owner -> the pair is
//            stored straight into (mx, out).
// x dropped: gridDim.x = G chunk groups, CTA c walks chunks c, c+G, ...; the per-CTA pairs go to
//            (mx, out)[o * G + c] and k_reduce_combine folds the G partials of every output.
// Elements are consumed in batches of U (one maximum, then U+1 independent table exponentials) with
// TWO batches of loads in flight.  The kernel was bound by instruction issue (65 instructions per
// element, ncu profiles/r2_*): full batches take an unpredicated path with pointer increments,
// equal velocity bins are folded into the per-(t,z) constant, the exponential has one range test.
constexpr int RED_U = 8;      // any_greater8() is written for this batch size
constexpr int RED_VPC = 16;    // kept velocity bins per CTA
// MINB = resident CTAs per SM the register budget is set for: 4 (64 registers) is best when the
// velocity axis is kept (batches along t: 7.7 vs 8.2 ms per pass of the 201x201 array), 3 (80) when
// it is summed (batches along v: 5.2 vs 5.5 ms)
template <int MINB, bool CONTIG>
__global__ void __launch_bounds__(256, MINB)
k_reduce_online(const double* __restrict__ logp, ReduceGeom g, const double* __restrict__ log_dt,
                const double* __restrict__ log_dv, const double* __restrict__ log_dz,
                const double* __restrict__ log_lw, const double* __restrict__ exp2_tab,
                double* __restrict__ mx, double* __restrict__ out, double* __restrict__ gmax, int PB,
                long long nchunk, int uniform_dv, int per_contig) {
  extern __shared__ double s_tab[];          // log_dt [nt] | log_dv [nv] | log_lw [nt][nz]
  __shared__ double red_m[256], red_s[256];
  __shared__ double s_exp[PKM_EXP_TAB];      // 2^(j/256) for exp_neg_tab()
  const int tid = threadIdx.x;
  const int nz_ = g.nz;
  double* s_dt = s_tab;
  double* s_dv = s_tab + g.nt;
  double* s_lw = s_dv + g.nv;
  // with equal velocity bins log dv is part of the per-thread constant, not of every element
  const double dv_all = (g.wt && uniform_dv) ? log_dv[0] : 0.0;
  for (int i = tid; i < g.nt; i += blockDim.x) s_dt[i] = g.wt ? log_dt[i] : 0.0;
  for (int i = tid; i < g.nv; i += blockDim.x) s_dv[i] = (g.wt && !uniform_dv) ? log_dv[i] : 0.0;
  if (log_lw) for (int i = tid; i < g.nt * nz_; i += blockDim.x) s_lw[i] = log_lw[i];
  for (int i = tid; i < PKM_EXP_TAB; i += blockDim.x) s_exp[i] = exp2_tab[i];
  __syncthreads();

  const int nact = PB * nz_;
  const int xl = tid / nz_, z = tid - xl * nz_;
  const long long G = gridDim.x;
  const long long row = g.npix * nz_;        // elements per (t, v)
  const bool active = tid < nact;
  int t0 = 0, t1 = g.nt;
  if (g.kt) { t0 = blockIdx.y; t1 = t0 + 1; }
  const int vblk0 = g.kv ? blockIdx.z * RED_VPC : 0;
  const int vblk1 = g.kv ? min(g.nv, vblk0 + RED_VPC) : 1;
  const double cz = (g.wt ? log_dz[min(z, nz_ - 1)] : 0.0) + g.log_dxdy + dv_all;
  const long long cstep = g.kx ? nchunk : G;            // x kept: exactly one chunk per CTA
  const long long part = g.kx ? 1 : G, pidx = g.kx ? 0 : blockIdx.x;
  const bool per_v = !uniform_dv && g.wt;               // per-element log dv

  for (int vo = vblk0; vo < vblk1; ++vo) {
    int v0 = 0, v1 = g.nv;
    if (g.kv) { v0 = vo; v1 = vo + 1; }
    const int nT = t1 - t0, nV = v1 - v0;
    Lse acc{-INFINITY, 0.0};
    // Consume a batch.  The running maximum M only ever grows, so the common case is "no element
    // exceeds M": one compare per element (no selects), then RED_U independent table exponentials
    // of a - M <= 0, each added unless it underflows (a predicated add instead of a select).  A
    // batch that raises M takes the slow path once (rescale the sum).  NaN compares false, passes
    // the underflow test and poisons the sum - scipy's logsumexp gives NaN as well.
    auto flush = [&acc](const double (&a)[RED_U], const double* tab) {
      const bool grow = any_greater8(a, acc.m);
      if (grow) {
        double bm = a[0];
#pragma unroll
        for (int k = 1; k < RED_U; ++k) bm = fmax(bm, a[k]);
        acc.s = acc.m > -INFINITY ? acc.s * exp_neg_tab(acc.m - bm, tab) : acc.s;   // bm > acc.m
        acc.m = bm;
      }
      if (acc.m > -INFINITY) {
        double sum = acc.s;
        if (!any_less8(a, acc.m - 708.0)) {
          // common case: no term underflows (none is -inf either), so every exponential is added
          // unconditionally; NaN compares false, lands here and poisons the sum
#pragma unroll
          for (int k = 0; k < RED_U; ++k) sum += exp_core_tab(a[k] - acc.m, tab);
        } else {
#pragma unroll
          for (int k = 0; k < RED_U; ++k) {
            const double x = a[k] - acc.m;
            add_unless_underflow(sum, x, exp_core_tab(x, tab));
          }
        }
        acc.s = sum;
      } else {                      // nothing finite seen so far: -inf elements add nothing, NaN poisons
        bool nan = false;
#pragma unroll
        for (int k = 0; k < RED_U; ++k) nan = nan || a[k] != a[k];
        if (nan) acc.s = NAN;
      }
    };
    if (active) {
      if (CONTIG && (nV > 1 || nT > 1)) {
        // x summed, many chunks per CTA: the CTA owns a CONTIGUOUS range of per_contig chunks and
        // the batches run along it, so that for every (t,v) row it streams one contiguous piece of
        // memory (8 chunks = 16 KB per batch) instead of 2 KB pieces megabytes apart (4.4 vs 5.0 ms
        // per pass of the 201x201 array).  per_contig is a multiple of RED_U: a batch never
        // straddles two rows.
        const long long c0 = (long long)blockIdx.x * per_contig;
        const int nc = (int)max(0LL, min(nchunk, c0 + per_contig) - c0);
        // chunks of this range in which this thread's pixel exists (only the array's last chunk is ragged)
        const int ncv = (nc > 0 && (c0 + nc - 1) * PB + xl >= g.npix) ? nc - 1 : nc;
        // `nrows` rows `rstride` elements apart starting at p; row r carries the constant crow(r)
        auto stream_rows = [&](const double* p, int nrows, long long rstride, auto crow) {
          int lr = 0, lc = 0;                    // position of the next batch to load: (row, chunk)
          auto load = [&](double (&a)[RED_U]) {
            const double cr = crow(min(lr, nrows - 1));
            const double* q = p + (long long)lr * rstride + (long long)lc * nact;
            if (lr < nrows && lc + RED_U <= ncv) {
#pragma unroll
              for (int k = 0; k < RED_U; ++k) a[k] = __ldg(q + k * nact) + cr;
            } else {
#pragma unroll
              for (int k = 0; k < RED_U; ++k) {
                const bool ok = lr < nrows && lc + k < ncv;
                a[k] = ok ? __ldg(q + k * nact) + cr : -INFINITY;
              }
            }
            lc += RED_U;
            if (lc >= per_contig) {
              lc = 0;
              ++lr;
            }
          };
          const int nbatch = nrows * (per_contig / RED_U);
          double a[RED_U], b[RED_U];
          load(a);
          for (int ib = 0; ib < nbatch; ib += 2) {
            load(b);
            flush(a, s_exp);
            load(a);
            if (ib + 1 < nbatch) flush(b, s_exp);   // a batch wholly past the end holds only -inf
          }
        };
        if (nV > 1) {                            // rows = velocity bins, once per age
          for (int t = t0; t < t1; ++t) {
            const double ct = cz + s_dt[t] + (log_lw ? s_lw[t * nz_ + z] : 0.0);
            stream_rows(logp + ((long long)t * g.nv + v0) * row + c0 * nact + tid, nV, row,
                        [&](int r) { return per_v ? ct + s_dv[v0 + r] : ct; });
          }
        } else {                                 // v kept, t summed: rows = ages
          const double cv = cz + s_dv[v0];
          stream_rows(logp + ((long long)t0 * g.nv + v0) * row + c0 * nact + tid, nT, (long long)g.nv * row,
                      [&](int r) { return cv + s_dt[t0 + r] + (log_lw ? s_lw[(t0 + r) * nz_ + z] : 0.0); });
        }
      } else if (nV > 1) {
        // batches along v (stride = one (t,v) row)
        for (long long chunk = blockIdx.x; chunk < nchunk; chunk += cstep) {
          if (chunk * PB + xl >= g.npix) continue;
          for (int t = t0; t < t1; ++t) {
            const double ct = cz + s_dt[t] + (log_lw ? s_lw[t * nz_ + z] : 0.0);
            const double* p = logp + ((long long)t * g.nv + v0) * row + chunk * nact + tid;
            auto load = [&](double (&a)[RED_U], int vb) {
              const double* q = p + (long long)vb * row;
              if (vb + RED_U <= nV) {
#pragma unroll
                for (int k = 0; k < RED_U; ++k) a[k] = __ldg(q + k * row) + (per_v ? ct + s_dv[v0 + vb + k] : ct);
              } else {
#pragma unroll
                for (int k = 0; k < RED_U; ++k) {
                  const bool ok = vb + k < nV;
                  a[k] = ok ? __ldg(q + k * row) + ct + s_dv[v0 + vb + k] : -INFINITY;
                }
              }
            };
            double a[RED_U], b[RED_U];
            load(a, 0);
            for (int vb = 0; vb < nV; vb += 2 * RED_U) {
              load(b, vb + RED_U);
              flush(a, s_exp);
              load(a, vb + 2 * RED_U);
              if (vb + RED_U < nV) flush(b, s_exp);
            }
          }
        }
      } else if (nT > 1) {
        // v kept, t summed: batches along t (stride = nv rows)
        const double cv = cz + s_dv[v0];
        const long long tstride = (long long)g.nv * row;
        for (long long chunk = blockIdx.x; chunk < nchunk; chunk += cstep) {
          if (chunk * PB + xl >= g.npix) continue;
          const double* p = logp + (long long)v0 * row + chunk * nact + tid;
          auto load = [&](double (&a)[RED_U], int tb) {
            const double* q = p + (long long)tb * tstride;
            if (tb + RED_U <= nT) {
#pragma unroll
              for (int k = 0; k < RED_U; ++k)
                a[k] = __ldg(q + k * tstride) + (cv + s_dt[tb + k] + (log_lw ? s_lw[(tb + k) * nz_ + z] : 0.0));
            } else {
#pragma unroll
              for (int k = 0; k < RED_U; ++k) {
                const bool ok = tb + k < nT;
                const int tt = ok ? tb + k : 0;
                a[k] = ok ? __ldg(q + k * tstride) + (cv + s_dt[tt] + (log_lw ? s_lw[tt * nz_ + z] : 0.0)) : -INFINITY;
              }
            }
          };
          double a[RED_U], b[RED_U];
          load(a, 0);
          for (int tb = 0; tb < nT; tb += 2 * RED_U) {
            load(b, tb + RED_U);
            flush(a, s_exp);
            load(a, tb + 2 * RED_U);
            if (tb + RED_U < nT) flush(b, s_exp);
          }
        }
      } else {
        // t and v kept: batches along this CTA's pixel chunks (a single element when x is kept too)
        const double c = cz + s_dt[t0] + (log_lw ? s_lw[t0 * nz_ + z] : 0.0) + s_dv[v0];
        const double* p = logp + ((long long)t0 * g.nv + v0) * row + tid;
        if (g.kx) {
          const long long chunk = blockIdx.x;
          if (chunk * PB + xl < g.npix) {
            const double a0 = __ldg(p + chunk * nact) + c;
            acc.m = a0 > -INFINITY ? a0 : -INFINITY;
            acc.s = a0 > -INFINITY ? 1.0 : (a0 != a0 ? NAN : 0.0);
          }
        } else {
          // per_contig > 0: this CTA's chunks are the contiguous range [cfirst, clast), else the
          // strided set blockIdx.x, blockIdx.x + G, ...
          const long long cst = per_contig > 0 ? 1 : cstep;
          const long long cfirst = per_contig > 0 ? (long long)blockIdx.x * per_contig : blockIdx.x;
          const long long clast = per_contig > 0 ? min(nchunk, cfirst + per_contig) : nchunk;
          // chunks below cfull exist for this thread's pixel (only the array's last chunk is ragged)
          const long long cfull = (nchunk - 1) * PB + xl < g.npix ? clast : min(clast, nchunk - 1);
          auto load = [&](double (&a)[RED_U], long long cb) {
            if (cb + (RED_U - 1) * cst < cfull) {
              const double* q = p + cb * nact;
#pragma unroll
              for (int k = 0; k < RED_U; ++k) a[k] = __ldg(q + k * cst * nact) + c;
            } else {
#pragma unroll
              for (int k = 0; k < RED_U; ++k) {
                const long long chunk = cb + k * cst;
                a[k] = chunk < cfull ? __ldg(p + chunk * nact) + c : -INFINITY;
              }
            }
          };
          double a[RED_U], b[RED_U];
          load(a, cfirst);
          for (long long cb = cfirst; cb < clast; cb += 2 * cst * RED_U) {
            load(b, cb + cst * RED_U);
            flush(a, s_exp);
            load(a, cb + 2 * cst * RED_U);
            if (cb + cst * RED_U < clast) flush(b, s_exp);
          }
        }
      }
    }

    if (gmax) {                                 // global maximum (light-weighted normalisation)
      double r = acc.m;
      for (int off = 16; off > 0; off >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, off));
      if ((tid & 31) == 0 && r > -INFINITY) atomic_max_double(gmax, r);
    }

    // combine over the dropped in-CTA axes and store the pair(s)
    if (g.kx && g.kz) {
      const long long x = (long long)blockIdx.x * PB + xl;
      if (active && x < g.npix) {
        const long long o = out_index(g, t0, v0, x, z);
        mx[o] = acc.m;
        out[o] = acc.s;
      }
      continue;
    }
    // Two-phase so that every thread evaluates ONE table exponential: (1) group maximum M from
    // shared memory, (2) s * exp(m - M) summed per group.  (A serial chain of pair merges by one
    // thread per group cost as much as the streaming loop.)  NaN sums are kept: 0 * NaN = NaN.
    red_m[tid] = active ? acc.m : -INFINITY;
    __syncthreads();
    if (g.kx) {            // groups = pixels: the nz elements of pixel xl are contiguous
      double M = -INFINITY;
      if (active)
        for (int zz = 0; zz < nz_; ++zz) M = fmax(M, red_m[xl * nz_ + zz]);
      red_s[tid] = active ? (M > -INFINITY ? acc.s * exp_neg_tab(acc.m - M, s_exp) : acc.s) : 0.0;
      __syncthreads();
      const long long x = (long long)blockIdx.x * PB + xl;
      if (active && z == 0 && x < g.npix) {
        double sum = 0.0;
        for (int zz = 0; zz < nz_; ++zz) sum += red_s[xl * nz_ + zz];
        const long long o = out_index(g, t0, v0, x, 0);
        mx[o] = M;
        out[o] = sum;
      }
    } else if (g.kz) {     // groups = metallicities: the PB pixels of z are strided by nz
      double M = -INFINITY;
      if (active)
        for (int xx = 0; xx < PB; ++xx) M = fmax(M, red_m[xx * nz_ + z]);
      red_s[tid] = active ? (M > -INFINITY ? acc.s * exp_neg_tab(acc.m - M, s_exp) : acc.s) : 0.0;
      __syncthreads();
      if (active && xl == 0) {
        double sum = 0.0;
        for (int xx = 0; xx < PB; ++xx) sum += red_s[xx * nz_ + z];
        const long long o = out_index(g, t0, v0, 0, z) * part + pidx;
        mx[o] = M;
        out[o] = sum;
      }
    } else {               // one group: the whole CTA
      for (int st = 128; st > 0; st >>= 1) {
        if (tid < st) red_m[tid] = fmax(red_m[tid], red_m[tid + st]);
        __syncthreads();
      }
      const double M = red_m[0];
      red_s[tid] = active ? (M > -INFINITY ? acc.s * exp_neg_tab(acc.m - M, s_exp) : acc.s) : 0.0;
      __syncthreads();
      for (int st = 128; st > 0; st >>= 1) {
        if (tid < st) red_s[tid] += red_s[tid + st];
        __syncthreads();
      }
      if (tid == 0) {
        const long long o = out_index(g, t0, v0, 0, 0) * part + pidx;
        mx[o] = M;
        out[o] = red_s[0];
      }
    }
    __syncthreads();       // red_m / red_s are reused by the next kept velocity bin
  }
}

// fold the G partial pairs of every output: (pm, ps)[o * G + c] -> (mx, out)[o]
__global__ void k_reduce_combine(const double* __restrict__ pm, const double* __restrict__ ps,
                                 long long nout, int G, double* __restrict__ mx,
                                 double* __restrict__ out) {
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= nout) return;
  const int lane = threadIdx.x & 31;
  Lse r{-INFINITY, 0.0};
  for (int c = lane; c < G; c += 32) r = lse_merge(r, Lse{pm[o * G + c], ps[o * G + c]});
  for (int off = 16; off > 0; off >>= 1) {
    Lse other{__shfl_xor_sync(0xffffffffu, r.m, off), __shfl_xor_sync(0xffffffffu, r.s, off)};
    r = lse_merge(r, other);
  }
  if (lane == 0) {
    mx[o] = r.m;
    out[o] = r.s;
  }
}

// total[0] += sum_o out[o] * exp(mx[o] - gmax)   (log normalisation = gmax + log total)
__global__ void k_reduce_total(const double* __restrict__ mx, const double* __restrict__ out,
                               long long nout, const double* __restrict__ gmax, double* total) {
  double acc = 0.0;
  const double gm = *gmax;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nout;
       i += (long long)gridDim.x * blockDim.x) {
    double m = mx[i];
    if (m > -INFINITY) acc += out[i] * exp(m - gm);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(total, acc);
}

__global__ void k_reduce_finalize(double* __restrict__ out, const double* __restrict__ mx,
                                  long long nout, ReduceGeom g, const double* __restrict__ log_dt,
                                  const double* __restrict__ log_dv, const double* __restrict__ log_dz,
                                  const double* gmax, const double* total, int density) {
  double lognorm = 0.0;
  if (total) lognorm = *gmax + log(*total);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nout;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    int z = 0, t = 0, v = 0;
    if (g.kz) { z = (int)(r % g.nz); r /= g.nz; }
    if (g.kx) { r /= g.npix; }
    if (g.kv) { v = (int)(r % g.nv); r /= g.nv; }
    if (g.kt) { t = (int)r; }
    double m = mx[i];
    const double sm = out[i];
    double val = (m > -INFINITY || sm != sm) ? m + log(sm) : -INFINITY;
    val -= lognorm;
    if (density) {
      double ldv = 0.0;
      if (g.kt) ldv += log_dt[t];
      if (g.kv) ldv += log_dv[v];
      if (g.kx) ldv += g.out_log_dxdy;
      if (g.kz) ldv += log_dz[z];
      val -= ldv;
    }
    out[i] = val;
  }
}

// all axes kept: out = log p + log dV (+ log lw) ; also feeds the global max for normalisation
__global__ void k_reduce_identity(const double* __restrict__ logp, ReduceGeom g,
                                  const double* __restrict__ log_dt, const double* __restrict__ log_dv,
                                  const double* __restrict__ log_dz, const double* __restrict__ log_lw,
                                  double* __restrict__ out, long long ntot) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ntot;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    int z = (int)(r % g.nz); r /= g.nz;
    r /= g.npix;
    int v = (int)(r % g.nv); r /= g.nv;
    int t = (int)r;
    const double ldV = g.wt ? log_dt[t] + log_dv[v] + g.log_dxdy + log_dz[z] : 0.0;
    out[i] = logp[i] + ldV + (log_lw ? log_lw[t * g.nz + z] : 0.0);
  }
}

// out[i] -= lognorm + (density ? log dV(t,v,x,z) : 0) for the all-axes-kept case
__global__ void k_sub_density(double* __restrict__ out, long long ntot, ReduceGeom g,
                              const double* __restrict__ log_dt, const double* __restrict__ log_dv,
                              const double* __restrict__ log_dz, const double* lognorm, int density) {
  const double ln = lognorm ? *lognorm : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ntot;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    int z = (int)(r % g.nz); r /= g.nz;
    r /= g.npix;
    int v = (int)(r % g.nv); r /= g.nv;
    int t = (int)r;
    double val = out[i] - ln;
    if (density) val -= log_dt[t] + log_dv[v] + g.out_log_dxdy + log_dz[z];
    out[i] = val;
  }
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_moments(const double* __restrict__ P, const double* __restrict__ values,
                          long long nvar, long long ncond, double* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncond) return;
  double mean = 0.0;
  for (long long i = 0; i < nvar; ++i) mean += values[i] * P[i * ncond + c];
  double m2 = 0.0, m3 = 0.0, m4 = 0.0;
  for (long long i = 0; i < nvar; ++i) {
    const double d = values[i] - mean, p = P[i * ncond + c];
    const double d2 = d * d;
    m2 += d2 * p;
    m3 += d2 * d * p;
    m4 += d2 * d2 * p;
  }
  out[c] = mean;
  out[ncond + c] = m2;
  out[2 * ncond + c] = m3;
  out[3 * ncond + c] = m4;
}

// Moments of a conditional distribution straight from the log-probability marginals
// (base.py:130-177, 208-260): P_i = exp(log_joint[a][i][b] - log_given[a][b]), i = variable index,
// (a, b) = conditioner index c = a * ninner + b.  log_given may be NULL (unconditional, P = exp(joint)).
// out[0][c] = mean, out[1..3][c] = central moments 2..4.  NaN where P(given) = 0, as in the reference.
__global__ void k_moments_logp(const double* __restrict__ lj, const double* __restrict__ lg,
                               const double* __restrict__ values, long long nouter, long long nvar,
                               long long ninner, double* __restrict__ out) {
  const long long ncond = nouter * ninner;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncond) return;
  const long long a = c / ninner, b = c - a * ninner;
  const double* col = lj + a * nvar * ninner + b;
  const double g = lg ? lg[c] : 0.0;
  double mean = 0.0;
  for (long long i = 0; i < nvar; ++i) mean += values[i] * exp(col[i * ninner] - g);
  double m2 = 0.0, m3 = 0.0, m4 = 0.0;
  for (long long i = 0; i < nvar; ++i) {
    const double d = values[i] - mean, p = exp(col[i * ninner] - g);
    const double d2 = d * d;
    m2 += d2 * p;
    m3 += d2 * d * p;
    m4 += d2 * d2 * p;
  }
  out[c] = mean;
  out[ncond + c] = m2;
  out[2 * ncond + c] = m3;
  out[3 * ncond + c] = m4;
}

// out[i] += sign * log dV(kept axes of i)   (log probability <-> log density of a marginal)
__global__ void k_apply_log_dv(double* __restrict__ out, long long nout, ReduceGeom g,
                               const double* __restrict__ log_dt, const double* __restrict__ log_dv,
                               const double* __restrict__ log_dz, double sign) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nout;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    int z = 0, t = 0, v = 0;
    if (g.kz) { z = (int)(r % g.nz); r /= g.nz; }
    if (g.kx) { r /= g.npix; }
    if (g.kv) { v = (int)(r % g.nv); r /= g.nv; }
    if (g.kt) { t = (int)r; }
    double ldv = 0.0;
    if (g.kt) ldv += log_dt[t];
    if (g.kv) ldv += log_dv[v];
    if (g.kx) ldv += g.out_log_dxdy;
    if (g.kz) ldv += log_dz[z];
    out[i] += sign * ldv;
  }
}

// Conditional / plain distribution from device-resident log-probability marginals (base.py:130-177):
//   out[dep axes..., given axes...] = log P(joint) - log P(given) - (density ? log dV(dep) : 0), optionally
//   exponentiated.  Axes inside each group are in (t, v, x, z) order; `e*` are the extents of the
//   joint (1 for an absent axis), `g*` flags mark the conditioning axes.  log_given may be NULL.
struct CondGeom {
  int et, ev, ez;
  long long ex;
  int gt, gv, gx, gz;      // 1 = axis belongs to the conditioners
  int density, do_exp;
  double log_dxdy;
};
__global__ void k_conditional(const double* __restrict__ lj, const double* __restrict__ lg, CondGeom c,
                              const double* __restrict__ log_dt, const double* __restrict__ log_dv,
                              const double* __restrict__ log_dz, long long n, double* __restrict__ out) {
  // output dimension list: dependents first, then conditioners, each in t, v, x, z order
  long long ext[8];
  int axis[8], nd = 0;
  const long long e4[4] = {c.et, c.ev, c.ex, c.ez};
  const int g4[4] = {c.gt, c.gv, c.gx, c.gz};
  for (int pass = 0; pass < 2; ++pass)
    for (int a = 0; a < 4; ++a)
      if (g4[a] == pass && e4[a] > 1) { ext[nd] = e4[a]; axis[nd] = a; ++nd; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long idx[4] = {0, 0, 0, 0};
    long long r = i;
    for (int d = nd - 1; d >= 0; --d) {
      idx[axis[d]] = r % ext[d];
      r /= ext[d];
    }
    const long long ij = ((idx[0] * c.ev + idx[1]) * c.ex + idx[2]) * c.ez + idx[3];
    double val = lj[ij];
    if (lg) {
      long long ig = 0;
      for (int a = 0; a < 4; ++a)
        if (g4[a]) ig = ig * e4[a] + idx[a];
      val -= lg[ig];                               // (-inf) - (-inf) = NaN where P(given) = 0, as in the reference
    }
    if (c.density) {
      double ldv = 0.0;
      if (!c.gt && c.et > 1) ldv += log_dt[idx[0]];
      if (!c.gv && c.ev > 1) ldv += log_dv[idx[1]];
      if (!c.gx && c.ex > 1) ldv += c.log_dxdy;
      if (!c.gz && c.ez > 1) ldv += log_dz[idx[3]];
      val -= ldv;
    }
    out[i] = c.do_exp ? exp(val) : val;
  }
}

// numpy.max semantics: NaN if any element is NaN
__global__ void k_max(const double* __restrict__ x, long long n, double* result, int* nan_flag) {
  double m = -INFINITY;
  bool has_nan = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double v = x[i];
    if (v != v) has_nan = true; else m = fmax(m, v);
  }
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (__any_sync(0xffffffffu, has_nan) && (threadIdx.x & 31) == 0) atomicExch(nan_flag, 1);
  if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_double(result, m);
}
__global__ void k_max_finish(double* result, const int* nan_flag) {
  if (*nan_flag) *result = NAN;
}

// Philox4x32-10 (Salmon et al. 2011), counter = (offset+i, stream), key = seed
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// One Philox call, one log / sqrt / sincospi per PAIR of voxels (both Box-Muller outputs are used):
// voxel with global index gi = offset + i takes the cosine (gi even) or sine (gi odd) deviate of the
// pair gi >> 1, so the noise of a voxel depends on (seed, global index) only - not on how a cube
// is cut into calls.  Thread = one aligned pair.
__global__ void k_noise(const double* __restrict__ ybar, long long n, double snr, int mode,
                        unsigned long long seed, unsigned long long offset,
                        const double* __restrict__ d_max, double* __restrict__ yobs,
                        double* __restrict__ sigma_out) {
  double K = 0.0;
  if (mode == 0) K = sqrt(*d_max) / snr;
  const unsigned long long odd = offset & 1ull;          // the first voxel is the second of its pair
  const long long npair = (n + (long long)odd + 1) >> 1;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < npair;
       j += (long long)gridDim.x * blockDim.x) {
    const long long i0 = 2 * j - (long long)odd, i1 = i0 + 1;   // local indices of the pair
    const bool h0 = i0 >= 0, h1 = i1 < n;
    const double y0 = h0 ? ybar[i0] : 0.0, y1 = h1 ? ybar[i1] : 0.0;
    const double s0 = (mode == 0) ? K * sqrt(y0) : y0 / snr;
    const double s1 = (mode == 0) ? K * sqrt(y1) : y1 / snr;
    if (sigma_out) {
      if (h0) sigma_out[i0] = s0;
      if (h1) sigma_out[i1] = s1;
    }
    if (yobs) {
      const unsigned long long ctr = (offset + (unsigned long long)(i0 + (long long)odd)) >> 1;
      uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x706b6d62u, 0u};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      // two 53-bit uniforms: u1 in (0,1], u2 in [0,1)
      const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1];
      const unsigned long long b = ((unsigned long long)c[2] << 32) | c[3];
      const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
      const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      const double r = sqrt(-2.0 * log(u1));
      if (h0) yobs[i0] = y0 + s0 * (r * cs);
      if (h1) yobs[i1] = y1 + s1 * (r * sn);
    }
  }
}

}  // namespace

// number of pixel-chunk groups (CTAs along x) when x is summed over: enough CTAs to fill the GPU
// several times over, never more than there are chunks
static int reduce_groups(const pkm_plan* pl, const ReduceGeom& g, long long nchunk) {
  const long long n_outer = (long long)(g.kt ? g.nt : 1) * (g.kv ? (g.nv + RED_VPC - 1) / RED_VPC : 1);
  long long G = ((long long)pl->num_sms * 32 + n_outer - 1) / n_outer;
  G = std::max<long long>(1, std::min<long long>(G, nchunk));
  return (int)G;
}

// register budget (resident CTAs per SM) of the instantiation to launch; PKM_RED_MINB=3|4 forces one
static int reduce_minb(const ReduceGeom& g) {
  if (const char* e = getenv("PKM_RED_MINB")) return atoi(e) == 4 ? 4 : 3;
  return g.kv ? 4 : 3;
}

// one online pass + (x dropped) the fold of the chunk-group partials; leaves the pairs in (mx, sum)
static int reduce_online(pkm_plan* pl, const double* logp, const ReduceGeom& g, const double* d_log_dv,
                         const double* d_log_lw, double* d_mx, double* d_sum, double* d_part,
                         long long nout, double* d_gmax, int uniform_dv, cudaStream_t st) {
  const int PB = 256 / g.nz;
  const long long nchunk = (g.npix + PB - 1) / PB;
  const size_t smem = sizeof(double) * ((size_t)g.nt + g.nv + (size_t)g.nt * g.nz);
  PKM_REQUIRE(smem <= 40 * 1024, "age/velocity/metallicity grid too large for the reduction kernel");
  if (g.kx) {
    dim3 grid((unsigned)nchunk, g.kt ? g.nt : 1, g.kv ? (g.nv + RED_VPC - 1) / RED_VPC : 1);
    auto kern = reduce_minb(g) == 4 ? k_reduce_online<4, false> : k_reduce_online<3, false>;
    kern<<<grid, 256, smem, st>>>(logp, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, d_log_lw, pl->d_exp_tab, d_mx, d_sum,
                                 d_gmax, PB, nchunk, uniform_dv, 0);
    return 1;
  }
  int G = reduce_groups(pl, g, nchunk);
  // enough chunks per CTA: contiguous ranges of a whole number of batches (see the kernel)
  int per_contig = 0;
  {
    // fewer groups are fine as long as the grid still fills the GPU several times over
    const long long n_outer = (long long)(g.kt ? g.nt : 1) * (g.kv ? (g.nv + RED_VPC - 1) / RED_VPC : 1);
    const long long Gc = std::min<long long>(G, std::max<long long>(1, nchunk / RED_U));
    if (!getenv("PKM_RED_STRIDED") && nchunk / Gc >= RED_U && n_outer * Gc >= (long long)pl->num_sms * 8) {
      per_contig = (int)(((nchunk + Gc - 1) / Gc + RED_U - 1) / RED_U) * RED_U;
      G = (int)((nchunk + per_contig - 1) / per_contig);
    }
  }
  double* pm = d_part;
  double* ps = d_part + nout * G;
  dim3 grid((unsigned)G, g.kt ? g.nt : 1, g.kv ? (g.nv + RED_VPC - 1) / RED_VPC : 1);
  const bool m4 = reduce_minb(g) == 4;
  auto kern = per_contig > 0 ? (m4 ? k_reduce_online<4, true> : k_reduce_online<3, true>)
                             : (m4 ? k_reduce_online<4, false> : k_reduce_online<3, false>);
  kern<<<grid, 256, smem, st>>>(logp, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, d_log_lw, pl->d_exp_tab, pm, ps, d_gmax,
                               PB, nchunk, uniform_dv, per_contig);
  k_reduce_combine<<<(unsigned)((nout + 7) / 8), 256, 0, st>>>(pm, ps, nout, G, d_mx, d_sum);
  return 2;
}

void launch_reduce_logp(pkm_plan* pl, const double* logp, const ReduceSrc& src, const double* d_log_dv,
                        const double* d_log_lw, uint32_t keep_mask, int density, double* out,
                        cudaStream_t st) {
  const int64_t npix = src.npix;
  ReduceGeom g;
  g.nt = src.nt; g.nv = src.nv; g.nz = src.nz; g.npix = npix;
  g.kt = keep_mask & 1; g.kv = (keep_mask >> 1) & 1; g.kx = (keep_mask >> 2) & 1; g.kz = (keep_mask >> 3) & 1;
  g.wt = src.weighted ? 1 : 0;
  g.log_dxdy = g.wt ? std::log(pl->d.dx) + std::log(pl->d.dy) : 0.0;
  g.out_log_dxdy = std::log(pl->d.dx) + std::log(pl->d.dy);
  PKM_REQUIRE(g.nz <= 256, "nz=%d > 256 unsupported by the reduction kernel", g.nz);
  const long long ntot = (long long)g.nt * g.nv * npix * g.nz;
  long long nout = 1;
  if (g.kt) nout *= g.nt;
  if (g.kv) nout *= g.nv;
  if (g.kx) nout *= npix;
  if (g.kz) nout *= g.nz;
  const int PB = 256 / g.nz;
  const long long nchunk = (npix + PB - 1) / PB;
  PKM_REQUIRE(nchunk < (1LL << 31), "too many pixels");
  const bool lw = d_log_lw != nullptr;
  const int fgrid = (int)std::min<long long>((nout + 255) / 256, (long long)pl->num_sms * 16);
  const bool all_kept = keep_mask == 0xF;

  // scratch: [gmax, total, mx0, sum0] + mx[nout] + partial pairs (x dropped)
  ReduceGeom g0 = g;
  g0.kt = g0.kv = g0.kx = g0.kz = 0;
  const size_t n_mx = all_kept ? 0 : (size_t)nout;
  const size_t n_part = all_kept ? 2 * (size_t)reduce_groups(pl, g0, nchunk)
                                 : (g.kx ? 0 : 2 * (size_t)nout * reduce_groups(pl, g, nchunk));
  pl->misc.reserve(sizeof(double) * (4 + n_mx + n_part));
  double* d_gmax = pl->misc.as<double>();
  double* d_total = d_gmax + 1;
  double* d_mx0 = d_gmax + 2;
  double* d_sum0 = d_gmax + 3;
  double* d_mx = d_gmax + 4;
  double* d_part = d_mx + n_mx;
  k_fill<<<1, 32, 0, st>>>(d_gmax, 1, -INFINITY);
  k_fill<<<1, 32, 0, st>>>(d_total, 1, 0.0);
  pl->launches += 2;

  if (all_kept) {
    const int grid = (int)std::min<long long>((ntot + 255) / 256, (long long)pl->num_sms * 16);
    k_reduce_identity<<<grid, 256, 0, st>>>(logp, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, d_log_lw, out, ntot);
    pl->launches += 1;
    if (lw) {
      // normalisation: scalar log-sum-exp of the weighted array (one more pass)
      pl->launches += reduce_online(pl, logp, g0, d_log_dv, d_log_lw, d_mx0, d_sum0, d_part, 1, nullptr, src.uniform_dv, st);
      // d_sum0 <- mx0 + log(sum0) = log normalisation (finalize on the 1-element array)
      k_reduce_finalize<<<1, 32, 0, st>>>(d_sum0, d_mx0, 1, g0, pl->d_log_dt, d_log_dv, pl->d_log_dz, nullptr, nullptr, 0);
      k_sub_density<<<grid, 256, 0, st>>>(out, ntot, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, d_sum0, density);
      pl->launches += 2;
    } else if (density) {
      k_sub_density<<<grid, 256, 0, st>>>(out, ntot, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, nullptr, density);
      pl->launches += 1;
    }
    PKM_CUDA(cudaGetLastError());
    return;
  }

  pl->launches += reduce_online(pl, logp, g, d_log_dv, d_log_lw, d_mx, out, d_part, nout, lw ? d_gmax : nullptr, src.uniform_dv, st);
  if (lw) k_reduce_total<<<fgrid, 256, 0, st>>>(d_mx, out, nout, d_gmax, d_total);
  k_reduce_finalize<<<fgrid, 256, 0, st>>>(out, d_mx, nout, g, pl->d_log_dt, d_log_dv, pl->d_log_dz,
                                           lw ? d_gmax : nullptr, lw ? d_total : nullptr, density);
  PKM_CUDA(cudaGetLastError());
  pl->launches += 1 + (lw ? 1 : 0);
}

void launch_moments(const double* P, const double* values, int64_t nvar, int64_t ncond, double* out,
                    cudaStream_t st) {
  if (ncond <= 0) return;
  k_moments<<<(unsigned)((ncond + 127) / 128), 128, 0, st>>>(P, values, nvar, ncond, out);
  PKM_CUDA(cudaGetLastError());
}

void launch_moments_logp(const double* log_joint, const double* log_given, const double* values,
                         int64_t nouter, int64_t nvar, int64_t ninner, double* out, cudaStream_t st) {
  const int64_t ncond = nouter * ninner;
  if (ncond <= 0) return;
  k_moments_logp<<<(unsigned)((ncond + 127) / 128), 128, 0, st>>>(log_joint, log_given, values, nouter, nvar,
                                                                 ninner, out);
  PKM_CUDA(cudaGetLastError());
}

// out (a marginal with the axes of keep_mask, extents taken from src) += sign * log dV
void launch_apply_log_dv(pkm_plan* pl, double* out, const ReduceSrc& src, const double* d_log_dv,
                         uint32_t keep_mask, double sign, cudaStream_t st) {
  ReduceGeom g;
  g.nt = src.nt; g.nv = src.nv; g.nz = src.nz; g.npix = src.npix;
  g.kt = keep_mask & 1; g.kv = (keep_mask >> 1) & 1; g.kx = (keep_mask >> 2) & 1; g.kz = (keep_mask >> 3) & 1;
  g.wt = 0; g.log_dxdy = 0.0;
  g.out_log_dxdy = std::log(pl->d.dx) + std::log(pl->d.dy);
  long long nout = 1;
  if (g.kt) nout *= g.nt;
  if (g.kv) nout *= g.nv;
  if (g.kx) nout *= g.npix;
  if (g.kz) nout *= g.nz;
  const int grid = (int)std::min<long long>((nout + 255) / 256, (long long)pl->num_sms * 16);
  k_apply_log_dv<<<grid, 256, 0, st>>>(out, nout, g, pl->d_log_dt, d_log_dv, pl->d_log_dz, sign);
  PKM_CUDA(cudaGetLastError());
  pl->launches += 1;
}

void launch_conditional(pkm_plan* pl, const double* log_joint, uint32_t joint_mask, const double* log_given,
                        uint32_t given_mask, int64_t npix, const double* d_log_dv, int density, int do_exp,
                        double* out, cudaStream_t st) {
  CondGeom c;
  c.et = (joint_mask & 1) ? pl->d.nt : 1;
  c.ev = (joint_mask & 2) ? pl->d.nv : 1;
  c.ex = (joint_mask & 4) ? npix : 1;
  c.ez = (joint_mask & 8) ? pl->d.nz : 1;
  c.gt = given_mask & 1; c.gv = (given_mask >> 1) & 1; c.gx = (given_mask >> 2) & 1; c.gz = (given_mask >> 3) & 1;
  c.density = density; c.do_exp = do_exp;
  c.log_dxdy = std::log(pl->d.dx) + std::log(pl->d.dy);
  const long long n = (long long)c.et * c.ev * c.ex * c.ez;
  const int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)pl->num_sms * 16));
  k_conditional<<<grid, 256, 0, st>>>(log_joint, log_given, c, pl->d_log_dt, d_log_dv, pl->d_log_dz, n, out);
  PKM_CUDA(cudaGetLastError());
  pl->launches += 1;
}

// d_result: 2 doubles of device scratch (value, then an int flag in the second slot)
void launch_max(const double* x, int64_t n, double* d_result, cudaStream_t st) {
  int* flag = reinterpret_cast<int*>(d_result + 1);
  k_fill<<<1, 32, 0, st>>>(d_result, 1, -INFINITY);
  PKM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8));
  k_max<<<grid, 256, 0, st>>>(x, n, d_result, flag);
  k_max_finish<<<1, 1, 0, st>>>(d_result, flag);
  PKM_CUDA(cudaGetLastError());
}

void launch_noise(const double* ybar, int64_t n, double snr, int mode, uint64_t seed, uint64_t offset,
                  const double* d_max, double* yobs, double* sigma, cudaStream_t st) {
  if (n <= 0) return;
  int grid = (int)std::min<int64_t>((n / 2 + 256) / 256, 148 * 16);
  k_noise<<<grid, 256, 0, st>>>(ybar, n, snr, mode, seed, offset, d_max, yobs, sigma);
  PKM_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------------
// 5-D particle histogram with numpy.histogramdd bin semantics (FromParticle,
// /root/reference/popkinmocks/components/particle.py:52-54): per axis the bin is
// searchsorted(edges, x, side="right") - 1, a sample equal to the last edge belongs to the last
// bin, anything outside (or NaN) is dropped.  Comparisons are the same float64 comparisons NumPy
// makes, so bin indices are bit-exact.  Unweighted counts use 64-bit integer atomics (exact);
// weighted sums use float64 atomics (order of additions differs from NumPy's bincount).
// ----------------------------------------------------------------------------
namespace {
struct HistGeom {
  int n[5];
  int off[5];   // offset of each axis' edges in the shared edge table
};

__device__ __forceinline__ int bin_of(const double* e, int nbin, double x) {
  // number of edges <= x, minus one (side="right"); x == last edge -> last bin
  if (!(x >= e[0]) || !(x <= e[nbin])) return -1;   // also rejects NaN
  if (x == e[nbin]) return nbin - 1;
  int lo = 0, hi = nbin;                             // invariant: e[lo] <= x < e[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
k_hist5d(const double* __restrict__ s0, const double* __restrict__ s1, const double* __restrict__ s2,
         const double* __restrict__ s3, const double* __restrict__ s4, const double* __restrict__ wts,
         long long n, const double* __restrict__ edges, HistGeom g, int nedge_total,
         unsigned long long* __restrict__ counts, double* __restrict__ sums) {
  extern __shared__ double s_edges[];
  for (int i = threadIdx.x; i < nedge_total; i += blockDim.x) s_edges[i] = edges[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double xs[5] = {s0[i], s1[i], s2[i], s3[i], s4[i]};
    long long flat = 0;
    bool inside = true;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const int b = bin_of(s_edges + g.off[a], g.n[a], xs[a]);
      inside = inside && b >= 0;
      flat = flat * g.n[a] + (b >= 0 ? b : 0);
    }
    if (!inside) continue;
    if (wts) atomicAdd(sums + flat, wts[i]);
    else atomicAdd(counts + flat, 1ull);
  }
}

// converts in place (in == out): no __restrict__
__global__ void k_u64_to_f64(const unsigned long long* in, double* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
}  // namespace

// out (device, nbins doubles) receives counts or weighted sums; samples/weights/edges on device
void launch_hist5d(const double* const samples[5], const double* weights, int64_t n, const double* d_edges,
                   const int nbin[5], double* out, int num_sms, cudaStream_t st) {
  HistGeom g;
  int off = 0;
  long long nbins = 1;
  for (int a = 0; a < 5; ++a) {
    g.n[a] = nbin[a];
    g.off[a] = off;
    off += nbin[a] + 1;
    nbins *= nbin[a];
  }
  PKM_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * nbins, st));
  if (n > 0) {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms * 16);
    // unweighted: count in place as u64 (same size as the float64 output), convert afterwards
    k_hist5d<<<grid, 256, sizeof(double) * off, st>>>(samples[0], samples[1], samples[2], samples[3], samples[4],
                                                     weights, n, d_edges, g, off,
                                                     reinterpret_cast<unsigned long long*>(out), out);
    PKM_CUDA(cudaGetLastError());
  }
  if (!weights) {
    const int grid = (int)std::min<int64_t>((nbins + 255) / 256, (int64_t)num_sms * 16);
    k_u64_to_f64<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(out), out, nbins);
    PKM_CUDA(cudaGetLastError());
  }
}

// ----------------------------------------------------------------------------
// Synthetic white-noise log density for the benchmark workload: a
// counter-based generator indexed by the GLOBAL element index of the [nt][nv][total_pix][nz]
// array, so any pixel range of the cube - a rank's shard, a spot pixel checked on the host -
// holds the same values whatever the sharding.  u = splitmix64(seed + index) mapped to (0,1);
// out = log(u * scale).  The host replica is popkinmocks_b200.engine.synth_log_density_host.
// ----------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_synth_logp(double* __restrict__ out, int nt, int nv, long long total_pix, long long pix0,
                             long long npix, int nz, unsigned long long seed, double scale) {
  const long long row = npix * nz;
  const long long n = (long long)nt * nv * row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tv = i / row, r = i - tv * row;          // r = x_local * nz + z
    const unsigned long long gidx = (unsigned long long)(tv * total_pix * nz + pix0 * nz + r);
    const unsigned long long h = splitmix64(seed + gidx);
    const double u = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    out[i] = log(u * scale);
  }
}
}  // namespace

void launch_synth_logp(double* out, int nt, int nv, int64_t total_pix, int64_t pix0, int64_t npix, int nz,
                       uint64_t seed, double scale, int num_sms, cudaStream_t st) {
  const long long n = (long long)nt * nv * npix * nz;
  if (n <= 0) return;
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms * 32);
  k_synth_logp<<<grid, 256, 0, st>>>(out, nt, nv, total_pix, pix0, npix, nz, seed, scale);
  PKM_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------------
// FP64 FMA throughput probe (roofline denominator of the FP64-pipe-bound kernels; the
// driver-written MEASURED_PEAKS.json has no FP64 figure).  8 independent DFMA chains per
// thread, 256 threads x 8 CTAs per SM.
// ----------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_fma_probe(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) sink[0] = s;  // never true; keeps the chain alive
}
}  // namespace

double run_fma_probe(int num_sms, int repeats) {
  double* sink = nullptr;
  PKM_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  PKM_CUDA(cudaEventCreate(&e0));
  PKM_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, grid = num_sms * 8;
  double best = 0.0;
  for (int r = 0; r < repeats + 1; ++r) {
    PKM_CUDA(cudaEventRecord(e0));
    k_fma_probe<<<grid, 256>>>(sink, iters, 0.999999, 1e-9);
    PKM_CUDA(cudaEventRecord(e1));
    PKM_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    PKM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
    if (r > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return best;
}

// ----------------------------------------------------------------------------
// Analytic pixelation of a Gaussian-LOSVD component on the device:
//   out[t,v,x,z] = log_w[t,x,z] + log( Phi((e[v+1]-mu[t,x])/sig[t,x]) - Phi((e[v]-mu)/sig) ) - ldv[v]
// i.e. ParametricComponent._get_log_p_tvxz (/root/reference/popkinmocks/components/
// parametric.py:567-596) with the velocity factor of _get_log_p_v_tx (parametric.py:180-211):
// normal log-cdf at the nv+1 velocity edges and a signed log-sum-exp of neighbouring edges.
// The 5-D array (20.8 GB at 201x201 px) never exists on the host; the kernel is bound by the
// HBM write of `out`.  One CTA = (age t, 32 consecutive pixels): the (nv+1) x 32 edge log-cdfs are
// evaluated once into shared memory, turned in place into the nv x 32 bin terms, and then every
// (t,v) row segment of 32*nz contiguous doubles is streamed out with evict-first stores.
// ----------------------------------------------------------------------------
namespace {
constexpr int PIX_PXC = 32;

// log of the standard normal cdf, accurate in both tails (erfcx form below -1, log1p form above 0)
__device__ __forceinline__ double log_ndtr_dev(double x) {
  const double rs2 = 0.70710678118654752440;
  if (x < -1.0) return log(0.5 * erfcx(-x * rs2)) - 0.5 * x * x;
  if (x < 0.0) return log(0.5 * erfc(-x * rs2));
  return log1p(-0.5 * erfc(x * rs2));
}

__global__ void __launch_bounds__(256, 4)
k_pixelate(const double* __restrict__ log_w, const double* __restrict__ mu,
           const double* __restrict__ sig, const double* __restrict__ v_edg,
           const double* __restrict__ log_dv, int nt, int nv, long long npix, int nz,
           double* __restrict__ out) {
  extern __shared__ double s_lc[];            // [PIX_PXC][nv + 1] edge log-cdfs, then bin terms
  const int ne = nv + 1;
  const int t = blockIdx.y;
  const long long x0 = (long long)blockIdx.x * PIX_PXC;
  const int npx = (int)min((long long)PIX_PXC, npix - x0);
  const int tid = threadIdx.x;

  for (int i = tid; i < npx * ne; i += blockDim.x) {
    const int px = i / ne, e = i - px * ne;
    const double m = mu[(long long)t * npix + x0 + px], s = sig[(long long)t * npix + x0 + px];
    s_lc[px * ne + e] = log_ndtr_dev((v_edg[e] - m) / s);
  }
  __syncthreads();
  // bin terms: each thread first reads all its (lo, hi) pairs, then the block overwrites in place
  constexpr int MAXI = 16;
  const int nitem = npx * nv;
  for (int base = 0; base < nitem; base += MAXI * blockDim.x) {
    double val[MAXI];
#pragma unroll
    for (int k = 0; k < MAXI; ++k) {
      const int i = base + k * blockDim.x + tid;
      val[k] = 0.0;
      if (i < nitem) {
        const int px = i / nv, v = i - px * nv;
        const double lo = s_lc[px * ne + v], hi = s_lc[px * ne + v + 1];
        // signed logsumexp([hi, lo], b=[1,-1]): hi + log(1 - exp(lo - hi)); an empty bin stays -inf
        double r = hi;
        if (hi != -INFINITY) r = hi + log1p(-exp(lo - hi));
        val[k] = r - (log_dv ? log_dv[v] : 0.0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXI; ++k) {
      const int i = base + k * blockDim.x + tid;
      if (i < nitem) {
        const int px = i / nv, v = i - px * nv;
        s_lc[px * ne + v] = val[k];            // slot (px, v) was read only by this thread or earlier
      }
    }
    __syncthreads();
  }
  // stream the rows: for fixed (t, v) the npx*nz outputs are contiguous
  const int nrow = npx * nz;
  for (int r0 = 0; r0 < nrow; r0 += 2 * blockDim.x) {
    const int i0 = r0 + tid, i1 = r0 + blockDim.x + tid;
    const bool ok0 = i0 < nrow, ok1 = i1 < nrow;
    const int p0 = ok0 ? i0 / nz : 0, p1 = ok1 ? i1 / nz : 0;
    const double w0 = ok0 ? log_w[((long long)t * npix + x0) * nz + i0] : 0.0;
    const double w1 = ok1 ? log_w[((long long)t * npix + x0) * nz + i1] : 0.0;
    double* o = out + (((long long)t * nv) * npix + x0) * nz;
    for (int v = 0; v < nv; ++v) {
      if (ok0) __stcs(o + i0, w0 + s_lc[p0 * ne + v]);
      if (ok1) __stcs(o + i1, w1 + s_lc[p1 * ne + v]);
      o += npix * nz;
    }
  }
}
}  // namespace

void launch_pixelate(const double* log_w, const double* mu, const double* sig, const double* d_v_edg,
                     const double* d_log_dv, int nt, int nv, int64_t npix, int nz, double* out,
                     cudaStream_t st) {
  if (npix <= 0) return;
  const size_t smem = sizeof(double) * PIX_PXC * (size_t)(nv + 1);
  PKM_REQUIRE(smem <= 200 * 1024, "nv=%d too large for the pixelation kernel", nv);
  PKM_REQUIRE(nt <= 65535, "nt=%d too large for the pixelation kernel", nt);
  if (smem > 48 * 1024)
    PKM_CUDA(cudaFuncSetAttribute(k_pixelate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((npix + PIX_PXC - 1) / PIX_PXC), (unsigned)nt);
  k_pixelate<<<grid, 256, smem, st>>>(log_w, mu, sig, d_v_edg, d_log_dv, nt, nv, (long long)npix, nz, out);
  PKM_CUDA(cudaGetLastError());
}
