// Table-based exp for the FP64-pipe-bound kernels (the fold of the coefficient kernel, the online
// log-sum-exp reductions), where the library exp (~26 FP64 instructions) dominated:
//   k = rint(x 256/ln 2),  exp(x) = 2^(k>>8) * T[k&255] * e^r,  T[j] = 2^(j/256) in shared memory,
//   degree-4 polynomial on |r| <= ln2/512 (truncation 4e-17) - 9 FP64 instructions, error < 2 ulp.
// Below -708 the result is flushed to 0 (the library would return a denormal < 3e-308), above 709
// it is +inf; NaN propagates.  (A conflict-free 16-entry table with a degree-8 polynomial measured
// slower: these kernels are bound by FP64 issue, not by the lookups.)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

constexpr int PKM_EXP_TAB = 256;

__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  const double t = fma(x, 369.32993046757463228, 6755399441055744.0);   // 256/ln2, 1.5 * 2^52
  const int k = __double2loint(t);
  const double fk = t - 6755399441055744.0;
  double r = fma(fk, -2.70760617331688990816e-03, x);                    // ln2/256 high part
  r = fma(fk, -7.45396456746323320321e-13, r);                           // ln2/256 low part
  double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  const double p = fma(r * r, q, r);                                     // e^r - 1
  const double T = tab[k & (PKM_EXP_TAB - 1)];
  const double v = fma(T, p, T);                                         // in [1, 2.01)
  const double y = __hiloint2double(__double2hiint(v) + ((k >> 8) << 20), __double2loint(v));
  // !(x >= -708) is true for NaN: keep it (x != x) instead of flushing
  return x < -708.0 ? 0.0 : (x > 709.0 ? INFINITY : (x != x ? x : y));
}

// host side: the 256 table entries, rounded from long double
inline void pkm_fill_exp_tab(double* tab) {
  for (int j = 0; j < PKM_EXP_TAB; ++j) tab[j] = (double)exp2l((long double)j / (long double)PKM_EXP_TAB);
}
