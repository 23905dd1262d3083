// k_trig.cuh — strict-parity sine and cosine for the kinematics / dynamics point functions.
//
// batotp evaluates sin/cos with the host libm (robot.cpp:130-136 fwdKinKuka, 196-199 fwdKinRR, 408-419 dynRR;
// util.cpp:544-549 aa2q).  CUDA's own FP64 sin/cos are good to 1-2 ulp but not bit-identical to it, and the BA
// amplifies a 1-ulp difference in a kinematic table into a different integer step count.  So
// that the trig-bearing point functions can run on the device AND reproduce the reference bit for bit, this file
// restates the algorithm of the libm the reference links against in this image — glibc 2.39, x86-64, the IBM
// Accurate Mathematical Library routines (upstream sysdeps/ieee754/dbl-64/s_sin.c: __sin, __cos, do_sin, do_cos,
// reduce_sincos, TAYLOR_SIN; table sincostab.c) — operation by operation:
//
//   |x| < 2^-26 (sin) / 2^-27 (cos)   x / 1
//   |x| < 0.855469                    do_sin(x, 0) / do_cos(x, 0): x = xk + r with xk = k/128 from the `big` trick,
//                                     sin/cos(xk) from the 440-entry hi/lo table, short polynomials in r
//   |x| < 2.426265                    sin: copysign(do_cos(pi/2 - |x|, lo(pi/2)), x); cos: do_sin(pi/2 - |x| ...)
//   |x| < 105414350                   reduce_sincos (x - n*pi/2 in three + two pieces), then do_sin / do_cos by n
//   larger, inf, nan                  outside the port (no joint angle gets there): CUDA's sin / cos
//
// glibc selects one of two ARITHMETICS at load time (sysdeps/x86_64/fpu/multiarch): on a CPU with FMA + AVX2 the
// routines were compiled with -mfma and GCC contracted a*b+c into fused multiply-adds; otherwise every product
// and sum is rounded on its own.  Both are here (template parameter FMA); which products are fused was read off
// the instruction stream of the libm in this image (the expression trees are those of the C source).  The host
// layer picks the variant of the machine it runs on and batotp_cuda_selftest_trig compares the device functions
// with the host's sin / cos on as many arguments as asked (10^9 in the GPU test suite: 0 mismatches required).
#pragma once
#include "emu.h"

namespace strig {

#if defined(__CUDA_ARCH__)
#define STRIG_TAB_QUAL __device__
#else
#define STRIG_TAB_QUAL
#endif
// {sin hi, sin lo, cos hi, cos lo} of k/128 (see scripts/gen_sincostab.py).  Two copies: nvcc needs the
// __device__ one for device code, the host passes (and the host emulation) read the plain one.
#ifndef BATOTP_HOST_EMU
static __device__ const double g_tab_dev[440] = {
#include "sincostab.inc"
};
#endif
static const double g_tab_host[440] = {
#include "sincostab.inc"
};
__host__ __device__ __forceinline__ double tab(int i) {
#if defined(__CUDA_ARCH__)
  return g_tab_dev[i];
#else
  return g_tab_host[i];
#endif
}

__host__ __device__ __forceinline__ double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
// a*b + c, c - a*b, a*b - c: fused, or with the product rounded first (the translation unit is compiled
// without contraction: -fmad=false / -ffp-contract=off)
template <bool FMA>
__host__ __device__ __forceinline__ double mad(double a, double b, double c) {
  if (FMA) return fma_(a, b, c);
  const double p = a * b;
  return p + c;
}
template <bool FMA>
__host__ __device__ __forceinline__ double nmad(double a, double b, double c) {  // c - a*b
  if (FMA) return fma_(-a, b, c);
  const double p = a * b;
  return c - p;
}
template <bool FMA>
__host__ __device__ __forceinline__ double msub(double a, double b, double c) {  // a*b - c
  if (FMA) return fma_(a, b, -c);
  const double p = a * b;
  return p - c;
}
__host__ __device__ __forceinline__ double copysign_(double mag, double sgn) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((__double2hiint(mag) & 0x7fffffff) | (__double2hiint(sgn) & 0x80000000), __double2loint(mag));
#else
  return __builtin_copysign(mag, sgn);
#endif
}
__host__ __device__ __forceinline__ int lo_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  long long b;
  memcpy(&b, &x, 8);
  return (int)(unsigned)(b & 0xffffffffll);
#endif
}
__host__ __device__ __forceinline__ int hi_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  long long b;
  memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}

// constants of s_sin.c / usncs.h
#define STRIG_BIG 0x1.8p45                      // 1.5 * 2^45: |x| + big rounds |x| to a multiple of 1/128
#define STRIG_SN3 -0x1.5555555555515p-3
#define STRIG_SN5 0x1.11110e829872fp-7
#define STRIG_CS2 0x1.0p-1
#define STRIG_CS4 -0x1.5555555555535p-5
#define STRIG_CS6 0x1.6c16bedd9e239p-10
#define STRIG_S1 -0x1.5555555555555p-3          // TAYLOR_SIN
#define STRIG_S2 0x1.1111111110ecep-7
#define STRIG_S3 -0x1.a01a019db08b8p-13
#define STRIG_S4 0x1.71de27b9a7ed9p-19
#define STRIG_S5 -0x1.addffc2fcdf59p-26
#define STRIG_HP0 0x1.921fb54442d18p+0          // pi/2, high and low part
#define STRIG_HP1 0x1.1a62633145c07p-54
#define STRIG_HPINV 0x1.45f306dc9c883p-1        // 2/pi
#define STRIG_TOINT 0x1.8p52
#define STRIG_MP1 0x1.921fb58p+0                // pi/2 in pieces for reduce_sincos
#define STRIG_MP2 -0x1.dde973cp-27
#define STRIG_PP3 -0x1.cb3b398p-55
#define STRIG_PP4 -0x1.d747f23e32ed7p-83

// x + ((POLYNOMIAL(xx)*x - 0.5*dx)*xx + dx), |x| < 0.126
template <bool FMA>
__host__ __device__ __forceinline__ double taylor_sin(double x, double dx) {
  const double xx = x * x;
  double p = mad<FMA>(STRIG_S5, xx, STRIG_S4);
  p = mad<FMA>(p, xx, STRIG_S3);
  p = mad<FMA>(p, xx, STRIG_S2);
  p = mad<FMA>(p, xx, STRIG_S1);  // ((((s5*xx + s4)*xx + s3)*xx + s2)*xx) + s1
  const double h = 0.5 * dx;
  const double t = mad<FMA>(msub<FMA>(p, x, h), xx, dx);
  return x + t;
}

// sin(x + dx) for |x| < 0.855469 (do_sin)
template <bool FMA>
__host__ __device__ __forceinline__ double do_sin(double x, double dx) {
  const double ax = fabs(x);
  if (ax < 0.126) return taylor_sin<FMA>(x, dx);
  if (x <= 0) dx = -dx;
  const double u = STRIG_BIG + ax;
  const double r = ax - (u - STRIG_BIG);
  const int k = lo_word(u) * 4;
  const double xx = r * r;
  const double s = r + mad<FMA>(r * xx, mad<FMA>(STRIG_SN5, xx, STRIG_SN3), dx);  // x + (dx + x*xx*(sn3 + xx*sn5))
  const double pc = mad<FMA>(mad<FMA>(STRIG_CS6, xx, STRIG_CS4), xx, STRIG_CS2);
  const double c = mad<FMA>(r, dx, xx * pc);  // x*dx + xx*(cs2 + xx*(cs4 + xx*cs6))
  const double sn = tab(k), ssn = tab(k + 1), cs = tab(k + 2), ccs = tab(k + 3);
  const double cor = mad<FMA>(s, cs, nmad<FMA>(c, sn, mad<FMA>(s, ccs, ssn)));  // (ssn + s*ccs - sn*c) + cs*s
  return copysign_(sn + cor, x);
}

// cos(x + dx) for |x| < 0.855469 (do_cos)
template <bool FMA>
__host__ __device__ __forceinline__ double do_cos(double x, double dx) {
  if (x < 0) dx = -dx;
  const double ax = fabs(x);
  const double u = STRIG_BIG + ax;
  const double r = (ax - (u - STRIG_BIG)) + dx;
  const int k = lo_word(u) * 4;
  const double xx = r * r;
  const double s = mad<FMA>(r * xx, mad<FMA>(STRIG_SN5, xx, STRIG_SN3), r);  // x + x*xx*(sn3 + xx*sn5)
  const double c = xx * mad<FMA>(mad<FMA>(STRIG_CS6, xx, STRIG_CS4), xx, STRIG_CS2);
  const double sn = tab(k), ssn = tab(k + 1), cs = tab(k + 2), ccs = tab(k + 3);
  const double cor = nmad<FMA>(s, sn, nmad<FMA>(c, cs, nmad<FMA>(s, ssn, ccs)));  // (ccs - s*ssn - cs*c) - sn*s
  return cs + cor;
}

// x - n*pi/2 as a + da, |x| < 105414350; returns n & 3 (reduce_sincos)
template <bool FMA>
__host__ __device__ __forceinline__ int reduce_sincos(double x, double &a, double &da) {
  const double t = mad<FMA>(x, STRIG_HPINV, STRIG_TOINT);
  const double xn = t - STRIG_TOINT;
  const double y = nmad<FMA>(xn, STRIG_MP2, nmad<FMA>(xn, STRIG_MP1, x));  // (x - xn*mp1) - xn*mp2
  const int n = lo_word(t) & 3;
  double b, db;
  if (FMA) {
    // GCC fused every use of the products xn*pp3 and xn*pp4 on its own: t1 is never rounded
    const double t2 = fma_(-xn, STRIG_PP3, y);
    db = fma_(-xn, STRIG_PP3, y - t2);
    b = fma_(-xn, STRIG_PP4, t2);
    db = db + fma_(-xn, STRIG_PP4, t2 - b);
  } else {
    double t1 = xn * STRIG_PP3;
    const double t2 = y - t1;
    db = (y - t2) - t1;
    t1 = xn * STRIG_PP4;
    b = t2 - t1;
    db += (t2 - b) - t1;
  }
  a = b;
  da = db;
  return n;
}

template <bool FMA>
__host__ __device__ __forceinline__ double do_sincos(double a, double da, int n) {
  const double r = (n & 1) ? do_cos<FMA>(a, da) : do_sin<FMA>(a, da);
  return (n & 2) ? -r : r;
}

// __sin
template <bool FMA>
__host__ __device__ inline double sin_libm(double x) {
  const int k = hi_word(x) & 0x7fffffff;
  if (k < 0x3e500000) return x;
  if (k < 0x3feb6000) return do_sin<FMA>(x, 0.0);
  if (k < 0x400368fd) {
    const double t = STRIG_HP0 - fabs(x);
    return copysign_(do_cos<FMA>(t, STRIG_HP1), x);
  }
  if (k < 0x419921FB) {
    double a, da;
    const int n = reduce_sincos<FMA>(x, a, da);
    return do_sincos<FMA>(a, da, n);
  }
  return sin(x);  // outside the port (|x| >= 105414350, inf, nan)
}

// __cos
template <bool FMA>
__host__ __device__ inline double cos_libm(double x) {
  const int k = hi_word(x) & 0x7fffffff;
  if (k < 0x3e400000) return 1.0;
  if (k < 0x3feb6000) return do_cos<FMA>(x, 0.0);
  if (k < 0x400368fd) {
    const double y = STRIG_HP0 - fabs(x);
    const double a = y + STRIG_HP1;
    const double da = (y - a) + STRIG_HP1;
    return do_sin<FMA>(a, da);
  }
  if (k < 0x419921FB) {
    double a, da;
    const int n = reduce_sincos<FMA>(x, a, da);
    return do_sincos<FMA>(a, da, n + 1);
  }
  return cos(x);
}

}  // namespace strig

// Trigonometry of the point functions by mode (cfg.trig_mode):
//   0  CUDA's sin / cos (1-2 ulp; results within the bisection tolerance, not bit-identical)
//   1  the host libm's algorithm without fused multiply-adds    } strict: bit-identical to the reference on a host
//   3  the host libm's algorithm with fused multiply-adds       } whose libm is the one ported (see above)
// (the C-ABI's trig_mode 1 is mapped to 1 or 3 by the host layer, from the CPU it runs on; 2 = evaluated by the
// host itself, host_strict.inl)
struct Trig {
  int mode;
  __host__ __device__ __forceinline__ double s(double x) const {
    if (mode == 3) return strig::sin_libm<true>(x);
    if (mode == 1) return strig::sin_libm<false>(x);
    return sin(x);
  }
  __host__ __device__ __forceinline__ double c(double x) const {
    if (mode == 3) return strig::cos_libm<true>(x);
    if (mode == 1) return strig::cos_libm<false>(x);
    return cos(x);
  }
};
