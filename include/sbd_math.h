/*
 * sbd_math.h -- "pinned" transcendental functions for the sea-breeze trigger path.
 *
 * The reference (generic/sea_breeze_diag.f90, python_wrapper/seabreezediag (both .f90 files)) calls the
 * Fortran intrinsics exp, atan2, sin, cos on default REAL (binary32); gfortran lowers them to the
 * platform libm (expf, atan2f, sinf, cosf).  libm results are neither correctly rounded nor the
 * same across glibc versions / CPUs (x86-64 glibc selects FMA variants by ifunc), so "bit-exact
 * against the reference" is only meaningful once the transcendental algorithms are pinned.
 *
 * Every function below is a fixed sequence of IEEE-754 basic operations (+,-,*,/ in binary32 or
 * binary64, conversions, integer ops) with NO fused multiply-add anywhere.  The same source is
 * compiled by g++ (-ffp-contract=off) for the CPU oracle and by nvcc (-fmad=false) for the sm_100a
 * kernels, so both sides execute identical roundings and agree bit-for-bit by construction.
 *
 * The algorithms follow the published algorithms glibc 2.39 uses for the same functions
 * (expf/sinf/cosf: double-precision polynomial kernels from ARM optimized-routines;
 * atanf/atan2f: the fdlibm binary32 kernels), so the pinned values equal this container's libm in
 * the overwhelming majority of inputs; tests/test_math_pinned.py measures the residual rate
 * per function and tests/test_libm_census.py on the outputs of the whole path (0 differing bits on c2 / c4).
 *
 * Attribution.  No code was copied; the functions below are restatements of published algorithms and use
 * their published constants (table entries, polynomial coefficients, argument-reduction constants):
 *   - expf, sinf, cosf: Arm Optimized Routines (math/expf.c, math/sincosf.h), Copyright (c) 2017-2019 Arm
 *     Limited, distributed under MIT OR Apache-2.0 WITH LLVM-exception; the same routines were
 *     contributed to glibc (sysdeps/ieee754/flt-32/e_expf.c, s_sincosf.h) under the LGPL-2.1-or-later.
 *   - atanf, atan2f: FreeBSD msun / fdlibm (s_atanf.c, e_atan2f.c), Copyright (C) 1993 by Sun Microsystems,
 *     Inc.  "Developed at SunPro, a Sun Microsystems, Inc. business.  Permission to use, copy, modify, and
 *     distribute this software is freely granted, provided that this notice is preserved."
 *
 * Host+device header: no CUDA runtime dependency.
 */
#ifndef SBD_MATH_H
#define SBD_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SBD_HD __host__ __device__ __forceinline__
#else
#define SBD_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SBD_F2U(x) __float_as_uint(x)
#define SBD_U2F(x) __uint_as_float(x)
#define SBD_D2U(x) ((uint64_t)__double_as_longlong(x))
#define SBD_U2D(x) __longlong_as_double((long long)(x))
#else
SBD_HD uint32_t sbd_f2u_(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SBD_HD float sbd_u2f_(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
SBD_HD uint64_t sbd_d2u_(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
SBD_HD double sbd_u2d_(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
#define SBD_F2U(x) sbd_f2u_(x)
#define SBD_U2F(x) sbd_u2f_(x)
#define SBD_D2U(x) sbd_d2u_(x)
#define SBD_U2D(x) sbd_u2d_(x)
#endif

SBD_HD float sbd_fabsf(float x) { return SBD_U2F(SBD_F2U(x) & 0x7fffffffu); }

/* ------------------------------------------------------------------------------------------ */
/* expf: N=32 table + degree-3 polynomial in binary64, result rounded once to binary32.        */
/* ------------------------------------------------------------------------------------------ */
#define SBD_EXP2F_N 32
/* T[i] = bits(2^(i/32)) - (i << 47) */
#define SBD_EXP2F_TAB                                                                            \
  {                                                                                              \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull  \
  }
#if defined(__CUDACC__)
static __device__ const uint64_t sbd_exp2f_tab_dev[SBD_EXP2F_N] = SBD_EXP2F_TAB;
#endif
static const uint64_t sbd_exp2f_tab_host[SBD_EXP2F_N] = SBD_EXP2F_TAB;

SBD_HD float sbd_expf(float x)
{
  const uint32_t ux = SBD_F2U(x);
  const uint32_t abstop = (ux >> 20) & 0x7ffu;
  if (abstop >= 0x42bu) { /* |x| >= 88 or NaN/Inf */
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return x + x;
    if (x > 88.72283172607421875f) return SBD_U2F(0x7f800000u); /* overflow -> +Inf */
    if (x < -103.972076416015625f) return 0.0f;                  /* underflow -> +0 */
  }
  const double invln2n = 0x1.71547652b82fep+0 * SBD_EXP2F_N;
  const double shift = 0x1.8p+52;
  const double c0 = 0x1.c6af84b912394p-5 / SBD_EXP2F_N / SBD_EXP2F_N / SBD_EXP2F_N;
  const double c1 = 0x1.ebfce50fac4f3p-3 / SBD_EXP2F_N / SBD_EXP2F_N;
  const double c2 = 0x1.62e42ff0c52d6p-1 / SBD_EXP2F_N;
  const double xd = (double)x;
  double z = invln2n * xd;
  double kd = z + shift;
  const uint64_t ki = SBD_D2U(kd);
  kd = kd - shift;
  const double r = z - kd;
#if defined(__CUDA_ARCH__)
  uint64_t t = sbd_exp2f_tab_dev[ki % SBD_EXP2F_N];
#else
  uint64_t t = sbd_exp2f_tab_host[ki % SBD_EXP2F_N];
#endif
  t += ki << (52 - 5);
  const double s = SBD_U2D(t);
  z = c0 * r;
  z = z + c1;
  const double r2 = r * r;
  double y = c2 * r;
  y = y + 1.0;
  double zr2 = z * r2;
  y = zr2 + y;
  y = y * s;
  return (float)y;
}

/* ------------------------------------------------------------------------------------------ */
/* sinf / cosf: binary64 polynomial kernels with fast pi/2 reduction (|x| < 120).              */
/* Beyond 120 the same reduction is applied with a rounded quotient (outside the path's domain: */
/* arguments are latitudes/longitudes in radians and half-differences of them).                 */
/* ------------------------------------------------------------------------------------------ */
SBD_HD double sbd_sincos_poly_(double x, double x2, int n, int neg)
{
  /* neg selects the sign-flipped cosine coefficient set (second table row in the published
     algorithm); the sine coefficients are shared. */
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
         C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
  if (neg) { C0 = -C0; C1 = -C1; C2 = -C2; C3 = -C3; C4 = -C4; }
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double s1 = S2 + x2 * S3;
    const double x7 = x3 * x2;
    const double s = x + x3 * S1;
    return s + x7 * s1;
  } else {
    const double x4 = x2 * x2;
    const double c2 = C3 + x2 * C4;
    const double c1 = C0 + x2 * C1;
    const double x6 = x4 * x2;
    const double c = c1 + x4 * C2;
    return c + x6 * c2;
  }
}

SBD_HD double sbd_reduce_(double x, int* np)
{
  const double hpi_inv = 0x1.45F306DC9C883p+23; /* 2/pi * 2^24 */
  const double hpi = 0x1.921FB54442D18p0;
  const double r = x * hpi_inv;
  int n;
  if (r < 2147483000.0 && r > -2147483000.0) {
    n = ((int32_t)r + 0x800000) >> 24;
  } else {
    n = 0; /* out of domain; caller guards with |x| < 120 */
  }
  *np = n;
  return x - (double)n * hpi;
}

SBD_HD float sbd_sinf(float y)
{
  const uint32_t at = (SBD_F2U(y) >> 20) & 0x7ffu;
  double x = (double)y;
  if (at < 0x3f4u) { /* |y| < pi/4 */
    const double s = x * x;
    if (at < 0x398u) return y; /* |y| < 2^-12 */
    return (float)sbd_sincos_poly_(x, s, 0, 0);
  }
  if (at >= 0x7f8u) return y - y; /* Inf/NaN -> NaN */
  int n;
  if (at >= 0x42fu) { /* |y| >= 120: coarse reduction, outside the path's domain */
    const double q = x * 0x1.45F306DC9C883p-1;
    const double k = (double)(long long)(q + (q < 0 ? -0.5 : 0.5));
    x = x - k * 0x1.921FB54442D18p0;
    n = (int)((long long)k & 3);
  } else {
    x = sbd_reduce_(x, &n);
  }
  const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return (float)sbd_sincos_poly_(x * sg, x * x, n, (n & 2) != 0);
}

SBD_HD float sbd_cosf(float y)
{
  const uint32_t at = (SBD_F2U(y) >> 20) & 0x7ffu;
  double x = (double)y;
  if (at < 0x3f4u) {
    const double x2 = x * x;
    if (at < 0x398u) return 1.0f;
    return (float)sbd_sincos_poly_(x, x2, 1, 0);
  }
  if (at >= 0x7f8u) return y - y;
  int n;
  if (at >= 0x42fu) {
    const double q = x * 0x1.45F306DC9C883p-1;
    const double k = (double)(long long)(q + (q < 0 ? -0.5 : 0.5));
    x = x - k * 0x1.921FB54442D18p0;
    n = (int)((long long)k & 3);
  } else {
    x = sbd_reduce_(x, &n);
  }
  const int m = (n + 1) & 3;
  const double sg = (m == 1 || m == 2) ? -1.0 : 1.0;
  return (float)sbd_sincos_poly_(x * sg, x * x, n ^ 1, (m & 2) != 0);
}

/* ------------------------------------------------------------------------------------------ */
/* atanf / atan2f: fdlibm binary32 kernels (all arithmetic in binary32).                        */
/* ------------------------------------------------------------------------------------------ */
SBD_HD float sbd_atanf(float x)
{
  const float atanhi0 = 4.6364760399e-01f, atanhi1 = 7.8539812565e-01f,
              atanhi2 = 9.8279368877e-01f, atanhi3 = 1.5707962513e+00f;
  const float atanlo0 = 5.0121582440e-09f, atanlo1 = 3.7748947079e-08f,
              atanlo2 = 3.4473217170e-08f, atanlo3 = 7.5497894159e-08f;
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
              aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
              aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
              aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const uint32_t hx = SBD_F2U(x);
  const uint32_t ix = hx & 0x7fffffffu;
  const int negx = (hx >> 31) != 0;
  float hi = 0.0f, lo = 0.0f;
  int id;
  if (ix >= 0x4c000000u) { /* |x| >= 2^25 */
    if (ix > 0x7f800000u) return x + x;
    float r = atanhi3 + atanlo3;
    return negx ? -r : r;
  }
  if (ix < 0x3ee00000u) { /* |x| < 0.4375 */
    if (ix < 0x31000000u) return x; /* |x| < 2^-29 */
    id = -1;
  } else {
    x = sbd_fabsf(x);
    if (ix < 0x3f980000u) {   /* |x| < 1.1875 */
      if (ix < 0x3f300000u) { /* 7/16 <= |x| < 11/16 */
        id = 0; hi = atanhi0; lo = atanlo0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {
        id = 1; hi = atanhi1; lo = atanlo1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000u) { /* |x| < 2.4375 */
        id = 2; hi = atanhi2; lo = atanlo2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else {
        id = 3; hi = atanhi3; lo = atanlo3;
        x = -1.0f / x;
      }
    }
  }
  const float z = x * x;
  const float w = z * z;
  float s1 = w * aT10; s1 = aT8 + s1; s1 = w * s1; s1 = aT6 + s1; s1 = w * s1; s1 = aT4 + s1;
  s1 = w * s1; s1 = aT2 + s1; s1 = w * s1; s1 = aT0 + s1; s1 = z * s1;
  float s2 = w * aT9; s2 = aT7 + s2; s2 = w * s2; s2 = aT5 + s2; s2 = w * s2; s2 = aT3 + s2;
  s2 = w * s2; s2 = aT1 + s2; s2 = w * s2;
  const float ss = s1 + s2;
  const float xs = x * ss;
  if (id < 0) return x - xs;
  float t = xs - lo;
  t = t - x;
  t = hi - t;
  return negx ? -t : t;
}

SBD_HD float sbd_atan2f(float y, float x)
{
  const float tiny = 1.0e-30f;
  const float pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const uint32_t hx = SBD_F2U(x), hy = SBD_F2U(y);
  const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
  if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
  if (hx == 0x3f800000u) return sbd_atanf(y);
  const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000u) {
    if (iy == 0x7f800000u) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = ((int)iy - (int)ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if ((hx >> 31) && k < -60) z = 0.0f;
  else z = sbd_atanf(sbd_fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

/* Fortran MODULO(a, p) for reals: result has the sign of p (gfortran: fmod + fix-up). */
#if defined(__CUDA_ARCH__)
#define SBD_FMODF(a, p) fmodf((a), (p))
#else
#include <math.h>
#define SBD_FMODF(a, p) fmodf((a), (p))
#endif
SBD_HD float sbd_modulo(float a, float p)
{
  float r = SBD_FMODF(a, p);
  if (r != 0.0f && ((r < 0.0f) != (p < 0.0f))) r = r + p;
  return r;
}

#endif /* SBD_MATH_H */
