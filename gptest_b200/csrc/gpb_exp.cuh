// gpb_exp.cuh - exp(x) for the covariance kernels, x <= ~0 (x = -r^2/2).
//
// ncu on the first version of se_build_kernel: 55 FP64 instructions per matrix element, 37 of them
// inside the library exp(); the kernel was FP64-issue bound at 60 % of the HBM roofline.  This version
// needs 10:  k = rint(64 x / ln2);  x = k ln2/64 + r (two-constant Cody-Waite, |r| <= ln2/128);
//            exp(x) = 2^(k>>6) * T[k&63] * (1 + r p(r)),  p of degree 4,  T[j] = 2^(j/64) (shared memory).
// Maximum error 1.01 ulp over [-60, 0] (tools: the derivation script is quoted in DESIGN.md), i.e. the same
// class as the library exp (1 ulp).  Results below 2^-1020 are flushed to zero.
#pragma once
#include <cuda_runtime.h>

namespace gpb {

static __device__ __constant__ double EXP2_64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// every thread of the CTA calls this once (blockDim.x >= 64), followed by __syncthreads()
__device__ __forceinline__ void exp_table_to_smem(double* tab) {
  if (threadIdx.x < 64) tab[threadIdx.x] = EXP2_64[threadIdx.x];
}

// exp(x) for x <= 700; tab = the 64-entry table in shared memory.  Below x = -708 (results under DBL_MIN, where
// numpy's exp returns subnormals and, from -745.13 on, exactly 0) the result is flushed to 0: an absolute error
// below 3.4e-308, and exact zeros wherever the reference has them.  The select also covers |x| > 2.3e7 (a scaled
// distance above ~6800: short length scales, line-search excursions), where k no longer fits the low word and the
// scale would be garbage.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  const double kd0 = fma(x, 0x1.71547652b82fep+6, 6755399441055744.0);   // 64/ln2, 2^52 + 2^51: rint in the low word
  const int k = __double2loint(kd0);
  const double kd = kd0 - 6755399441055744.0;
  double r = fma(kd, -0x1.62e42fee00000p-7, x);                          // ln2/64, high 32 bits (k * hi is exact)
  r = fma(kd, -0x1.a39ef35793c76p-39, r);                                // ln2/64, low part
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double t = tab[k & 63];
  const double s = fma(t, r * p, t);                                     // T * (1 + r p(r)), one rounding
  const int m = max(k >> 6, -1022);
  const double v = s * __hiloint2double((m + 1023) << 20, 0);            // exact scaling by 2^m
  return x < -708.0 ? 0.0 : v;
}

// Radial part of the stationary covariance functions, from x = -r^2/2 (r = scaled distance):
//   KIND 0  squared exponential   exp(-r^2/2)                                  (GPr.py:102)
//   KIND 1  Matern 3/2            (1 + a) exp(-a),            a = sqrt(3) r
//   KIND 2  Matern 5/2            (1 + a + a^2/3) exp(-a),    a = sqrt(5) r
// radial_dl is g(r) in  dk/dlog l_k = sf2 * g(r) * (x_ik - x_jk)^2 / l_k^2  (the ARD length-scale derivative):
//   SE: exp(-r^2/2)     Matern 3/2: 3 exp(-a)     Matern 5/2: (5/3) (1 + a) exp(-a)
template <int KIND>
__device__ __forceinline__ double radial(double x, const double* __restrict__ tab) {
  if (KIND == 0) return exp_tab(x, tab);
  const double r = sqrt(fmax(-2.0 * x, 0.0));
  if (KIND == 1) {
    const double a = 1.7320508075688772 * r;
    return (1.0 + a) * exp_tab(-a, tab);
  }
  const double a = 2.23606797749979 * r;
  return fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0) * exp_tab(-a, tab);
}
template <int KIND>
__device__ __forceinline__ void radial_and_dl(double x, const double* __restrict__ tab, double& k, double& g) {
  if (KIND == 0) { k = exp_tab(x, tab); g = k; return; }
  const double r = sqrt(fmax(-2.0 * x, 0.0));
  if (KIND == 1) {
    const double a = 1.7320508075688772 * r, e = exp_tab(-a, tab);
    k = (1.0 + a) * e; g = 3.0 * e; return;
  }
  const double a = 2.23606797749979 * r, e = exp_tab(-a, tab);
  k = fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0) * e;
  g = (5.0 / 3.0) * (1.0 + a) * e;
}

}  // namespace gpb
