// rbvfit_b200 -- device-side building blocks of the Voigt likelihood kernels (sm_100a).
//
// The arithmetic restates the reference's forward model (src/rbvfit/core/voigt_model.py:100-230,
// Appendix A of SURVEY.md); only the evaluation strategy is new:
//
//   x      = A_l * (1/lambda_p) - B_l           one FMA per (line, pixel); A_l, B_l per (walker, line)
//   d      = x^2 + a_l^2 = |z|^2
//   tier   = warp-uniform choice from min d over the warp's 128 pixels (integer min of the high words)
//     d >= 4e4 : H = (a/sqrt(pi)) rho (q1 + q2 rho + q3 rho^2),            rho = 1/d      [3 coefficients]
//     d >= 576 : same series with 6 coefficients
//     d >= 64  : same series with 13 coefficients
//     d <  64  : a <= A_FAST: exp(a^2-x^2) cos(2ax) + a sum_k a^(2k) g_k(x) with tabulated g_0..g_3
//                a >  A_FAST: Weideman N=40 rational approximation
//   The q_p(a^2) are polynomials in a^2 (exact truncated asymptotic series of w(z), any a), scaled by
//   kappa = N f K a / sqrt(pi) once per (walker, line) so that the per-pixel work is a Horner chain in rho.
//
// Measured accuracy against scipy.special.wofz / mpmath (tests/test_faddeeva.py): relative error of H
// <= 3e-13 over a in [1e-8, 20], |x| in [0, 3e4].
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "faddeeva_tables.h"

namespace rbv {

#ifndef RBV_MIN_CTAS
#define RBV_MIN_CTAS 2
#endif
#ifndef RBV_THREADS
#define RBV_THREADS 256
#endif
constexpr int kThreads = RBV_THREADS;
constexpr int kMaxHalo = 8192;                   // K - 1 <= kMaxHalo (tiles have at most 256 * 256 flux slots)
constexpr int kSuperPix = 1024;                  // super-chunk: unit of tier classification + far field
constexpr int kWantWaves = 1;                    // biggest tiles that still fill every CTA slot of the GPU once
constexpr int kSmallChunkLimit = 17;             // tiles with fewer 256-px chunks than this use 64-px chunks

// line-constant record (doubles)
constexpr int LC_A = 0, LC_B = 1, LC_A2 = 2, LC_a = 3, LC_Q = 4 /* Q1..Q13 */, LC_COEF = 17,
              LC_F32B = 18 /* float2: Q3, unused */, LC_AUX = 19 /* kappa */,
              LC_F32A = 20 /* float4: A, a^2, Q1, Q2 in FP32 for the gated far-wing path */, LC_STRIDE = 22;

constexpr double kAFast = 0.05;               // table path valid for a <= kAFast
constexpr double kABig = 1.0;                 // rho-polynomial series valid for a <= kABig; beyond: complex series
constexpr double kDCore = 64.0;               // |z|^2 below which the core evaluation is used
constexpr double kDNear = 576.0;              // 13-coefficient series below this
constexpr double kDFar = 40000.0;             // 6-coefficient series below this, 3 above
constexpr int kHiCore = 0x40500000;           // high words of the three thresholds (low words are zero)
constexpr int kHiNear = 0x40820000;
constexpr int kHiFar = 0x40E38800;
constexpr int kNQFar = 3, kNQMid = 6, kNQNear = 13;

constexpr double kInvSqrtPi = 0.56418958354775628695;
constexpr double kSqrtPi = 1.7724538509055160273;
constexpr double kFourPi = 12.566370614359172;   // 4*np.pi as numpy evaluates it
constexpr double kCFreq = 2.99792458e18;         // voigt_model.py:131
constexpr double kAtomicConst = 4.48898479507e3; // voigt_model.py:132
constexpr double kCkms = 299792.458;             // voigt_model.py:197

__constant__ double c_ctab[(RBV_ASYM_PMAX + 1) * (RBV_ASYM_MMAX + 1)];
__constant__ double c_weid[RBV_WEID_N];
__constant__ double c_ff_nodes[RBV_FF_M];              // Chebyshev nodes of the far-field interpolant
__constant__ double c_ff_minv[RBV_FF_M * RBV_FF_M];    // node values -> monomial coefficients, [power][node]

// ---------------------------------------------------------------------------------------------- helpers
// Full-precision reciprocal of a positive, normal double: MUFU seed + one cubically convergent step.
// rcp.approx.ftz.f64 returns >= 20 good bits (checked on the device by rbv_selftest), so e^3 < 2^-60.
__device__ __forceinline__ double rcp_pos(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  double t = fma(e, e, e);
  return fma(r, t, r);
}

// q_p(a^2) * kappa for p = 1..nq  -> Q[0..nq)
__device__ __forceinline__ double asym_coef(int p, double a2, double kappa) {
  const double* row = c_ctab + p * (RBV_ASYM_MMAX + 1);
  double q = row[RBV_ASYM_MMAX];
#pragma unroll
  for (int m = RBV_ASYM_MMAX - 1; m >= 0; --m) q = fma(q, a2, row[m]);
  return q * kappa;
}

// sum_{p=1..NQ} Q_p rho^p for one pixel
template <int NQ>
__device__ __forceinline__ double asym_series(const double* __restrict__ Q, double d) {
  double rho = rcp_pos(d);
  double s = Q[NQ - 1];
#pragma unroll
  for (int p = NQ - 2; p >= 0; --p) s = fma(s, rho, Q[p]);
  return s * rho;
}

// Truncated asymptotic series of w(z) in complex arithmetic (k = 0..12), for lines with a > kABig where
// a^2 rho is not small and the rho-polynomial rearrangement above loses accuracy.  |z|^2 >= 64.
//   w(z) ~ (i/sqrt(pi)) (1/z) sum_k c_k (1/z^2)^k,  c_k = (2k-1)!!/2^k   =>   Re w = -(1/sqrt(pi)) Im[(1/z) S]
__device__ __noinline__ double asym_complex(double x, double a, double d) {
  double rho = 1.0 / d;
  double rr = x * rho, ri = -a * rho;                 // 1/z
  double sr = fma(rr, rr, -ri * ri), si = 2.0 * rr * ri;   // 1/z^2
  double c[13];
  c[0] = 1.0;
#pragma unroll
  for (int k = 1; k <= 12; ++k) c[k] = c[k - 1] * (2 * k - 1) * 0.5;
  double Sr = c[12], Si = 0.0;
#pragma unroll
  for (int k = 11; k >= 0; --k) {
    double tr = fma(Sr, sr, fma(-Si, si, c[k]));
    double ti = fma(Sr, si, Si * sr);
    Sr = tr;
    Si = ti;
  }
  return -kInvSqrtPi * fma(rr, Si, ri * Sr);
}

// exp(x) for the flux: Cody-Waite reduction x = k ln2 + r, |r| <= 0.347, degree-13 Taylor polynomial
// (truncation 4e-18), scaling by an exponent-field add.  Coefficients sit in constant memory so that every
// DFMA takes its constant as a c[][] operand (CUDA's exp() spends 23 UMOVs per call on them).
// Results below 2^-1020 flush to 0 (np.exp would return a denormal < 1e-307: irrelevant at |dflux| <= 1e-10).
__constant__ double c_exp[14] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};

__device__ __forceinline__ double exp_flux(double x) {
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);   // round(x log2 e) in the low word
  const int k = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = c_exp[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, c_exp[i]);
  double y = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  if (!(x >= -707.0)) y = (x != x) ? x : 0.0;     // underflow (and NaN passes through)
  if (x > 709.0) y = CUDART_INF;
  return y;
}

// Core evaluation of H(a,x) for |z|^2 < 64, a <= kAFast, from the g_k tables (global memory, L1-resident).
__device__ __noinline__ double core_H_table(double x, double a, double a2, const double* __restrict__ tab) {
  double ax = fabs(x);
  int j = (int)(ax * RBV_CORE_INV_H);
  j = max(0, min(j, RBV_CORE_NINT - 1));
  double t = fma(ax, 2.0 * RBV_CORE_INV_H, -(double)(2 * j + 1));
  const double* p0 = tab + RBV_CORE_OFF0 + j;
  const double* p1 = tab + RBV_CORE_OFF1 + j;
  const double* p2 = tab + RBV_CORE_OFF2 + j;
  const double* p3 = tab + RBV_CORE_OFF3 + j;
  double g0 = __ldg(p0 + RBV_CORE_DEG0 * RBV_CORE_NINT);
#pragma unroll
  for (int k = RBV_CORE_DEG0 - 1; k >= 0; --k) g0 = fma(g0, t, __ldg(p0 + k * RBV_CORE_NINT));
  double g1 = __ldg(p1 + RBV_CORE_DEG1 * RBV_CORE_NINT);
#pragma unroll
  for (int k = RBV_CORE_DEG1 - 1; k >= 0; --k) g1 = fma(g1, t, __ldg(p1 + k * RBV_CORE_NINT));
  // the a^5 g2 and a^7 g3 terms matter only for strong damping: relative to a g0 they are a^4 and a^6 (times ratios
  // below one), i.e. <= 1e-14 once a^2 <= 1e-7 / 1e-5 -- true for every physical line at b >= 1 km/s in the far UV
  // (a ~ 1e-4); skipping them saves 12 of the 32 table loads (the branch is uniform: a is per line)
  double G = fma(g1, a2, g0);
  if (a2 > 1e-7) {
    double g2 = __ldg(p2 + RBV_CORE_DEG2 * RBV_CORE_NINT);
#pragma unroll
    for (int k = RBV_CORE_DEG2 - 1; k >= 0; --k) g2 = fma(g2, t, __ldg(p2 + k * RBV_CORE_NINT));
    double g3 = 0.0;
    if (a2 > 1e-5) {
      g3 = __ldg(p3 + RBV_CORE_DEG3 * RBV_CORE_NINT);
#pragma unroll
      for (int k = RBV_CORE_DEG3 - 1; k >= 0; --k) g3 = fma(g3, t, __ldg(p3 + k * RBV_CORE_NINT));
    }
    G = fma(fma(fma(g3, a2, g2), a2, g1), a2, g0);
  }
  double E = exp_flux(fma(-x, x, a2));            // exp(a^2 - x^2), argument in (-64, 0.0025]
  double ax_ = a * x;
  double y = ax_ * ax_;                           // cos(2 a x) = sum_k (-4 y)^k / (2k)!
  double c = -4.0 / 14175.0;
  c = fma(c, y, 2.0 / 315.0);
  c = fma(c, y, -4.0 / 45.0);
  c = fma(c, y, 2.0 / 3.0);
  c = fma(c, y, -2.0);
  c = fma(c, y, 1.0);
  return fma(E, c, a * G);
}

// Weideman (1994) N = 40: Re w(x + i a) for a > kAFast, |z|^2 < 64 (relative error <= 2e-13 there).
__device__ __noinline__ double core_H_weideman(double x, double a) {
  const double L = RBV_WEID_L;
  double dr = L + a, di = -x;   // L - i z
  double nr = L - a, ni = x;    // L + i z
  double inv = 1.0 / fma(dr, dr, di * di);
  double Zr = fma(nr, dr, ni * di) * inv;
  double Zi = fma(ni, dr, -nr * di) * inv;
  double pr = c_weid[0], pi = 0.0;
#pragma unroll 4
  for (int k = 1; k < RBV_WEID_N; ++k) {
    double tr = fma(pr, Zr, fma(-pi, Zi, c_weid[k]));
    double ti = fma(pr, Zi, pi * Zr);
    pr = tr;
    pi = ti;
  }
  double ir = dr * inv, ii = -di * inv;                 // 1/(L - i z)
  double i2r = fma(ir, ir, -ii * ii), i2i = 2.0 * ir * ii;
  return fma(2.0, fma(pr, i2r, -pi * i2i), kInvSqrtPi * ir);
}

// H(a,x) for any x when a > kABig (per-line general path; unphysical damping, correctness only)
__device__ __forceinline__ double general_H(double x, double a, double d) {
  return (d < kDCore) ? core_H_weideman(x, a) : asym_complex(x, a, d);
}

// H(a,x) for |z|^2 < 64 (any a > 0)
__device__ __forceinline__ double core_H(double x, double a, double a2, const double* __restrict__ tab) {
  return (a <= kAFast) ? core_H_table(x, a, a2, tab) : core_H_weideman(x, a);
}

// Tepper-Garcia variant exactly as the reference writes it (voigt_approx.py:69-86).
__device__ __forceinline__ double tg_H(double x, double a_over_sqrtpi, double eps, double core_fac) {
  double x2 = x * x;
  double G = exp(-x2);
  double safe = fmax(x2, eps);
  double numer = G * (4.0 * (safe * safe) + 7.0 * safe + 4.0) - 1.5;
  double sp1 = safe + 1.0;
  double denom = safe * (sp1 * sp1);
  double Htg = G - a_over_sqrtpi * numer / denom;
  double Hcore = G * core_fac;
  return (x2 < eps) ? Hcore : Htg;
}

// exp(x) for -2^-6 < x <= 0: degree-6 Taylor polynomial, truncation |x|^7/7! <= 4.5e-17 (less than half an ulp of
// the result, which lies in (0.98, 1])
__device__ __forceinline__ double exp_small(double x) {
  double p = c_exp[6];
#pragma unroll
  for (int i = 5; i >= 0; --i) p = fma(p, x, c_exp[i]);
  return p;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rbv
