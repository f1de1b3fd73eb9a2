// exp() and log() with the results of the libm the reference runs on: a transcription, operation by operation
// (same fused multiply-adds, same order), of the x86-64 FMA variants of exp and log in glibc 2.39 (the
// optimized-routines algorithms: table of 2^(j/128) / of (1/c, log c), short polynomials).  IEEE-754 operations are
// deterministic, so the device reproduces the host's bits; tools/check_glibc_math.c compares this file, compiled
// for the host, with libm on hundreds of millions of arguments, and tests/test_gpu_post_sw.py the device with the
// host.  post_sw's forward-backward (sw-post.c) sums exp() terms whose exact values tie in exact arithmetic (two
// equally likely colour errors), so the last bit decides base calls: hence bit-exact, not merely accurate.
#pragma once
#include <stdint.h>
#include "glibc_tables.inc"

#ifdef __CUDACC__
#define GM_HD __host__ __device__ __forceinline__
#define GM_SLOW __host__ __device__ __noinline__
#else
#define GM_SLOW static inline
#include <math.h>
#include <string.h>
#define GM_HD static inline
#endif

namespace glibc_math {

// The scalar constants: operands straight from the constant bank in device code (a 64-bit immediate costs two moves,
// a global load a round trip), literals on the host
#ifdef __CUDACC__
static __constant__ unsigned long long GM_DEV_EXP[8] = {GLIBC_EXP_CONST_0, GLIBC_EXP_CONST_1, GLIBC_EXP_CONST_2, GLIBC_EXP_CONST_3,
                                                        GLIBC_EXP_CONST_4, GLIBC_EXP_CONST_5, GLIBC_EXP_CONST_6, GLIBC_EXP_CONST_7};
static __constant__ unsigned long long GM_DEV_LOG[18] = {
    GLIBC_LOG_CONST_0,  GLIBC_LOG_CONST_1,  GLIBC_LOG_CONST_2,  GLIBC_LOG_CONST_3,  GLIBC_LOG_CONST_4,  GLIBC_LOG_CONST_5,
    GLIBC_LOG_CONST_6,  GLIBC_LOG_CONST_7,  GLIBC_LOG_CONST_8,  GLIBC_LOG_CONST_9,  GLIBC_LOG_CONST_10, GLIBC_LOG_CONST_11,
    GLIBC_LOG_CONST_12, GLIBC_LOG_CONST_13, GLIBC_LOG_CONST_14, GLIBC_LOG_CONST_15, GLIBC_LOG_CONST_16, GLIBC_LOG_CONST_17};
#endif
#ifdef __CUDA_ARCH__
#define GM_EC(i) gm_asdouble(GM_DEV_EXP[i])
#define GM_LC(i) gm_asdouble(GM_DEV_LOG[i])
#else
#define GM_EC(i) gm_asdouble(GLIBC_EXP_CONST_##i)
#define GM_LC(i) gm_asdouble(GLIBC_LOG_CONST_##i)
#endif

GM_HD double gm_asdouble(unsigned long long u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
GM_HD unsigned long long gm_asuint64(double d) {
#ifdef __CUDA_ARCH__
  return (unsigned long long)__double_as_longlong(d);
#else
  unsigned long long u; memcpy(&u, &d, 8); return u;
#endif
}

// The tables are passed in (device: pointers to global copies; host: the static arrays)
struct Tables {
  const unsigned long long *exp_const, *exp_tab, *log_const, *log_tab;
};

GM_HD double exp_special(double tmp, unsigned long long sbits, unsigned long long ki) {   // specialcase(), e_exp.c
  double scale, y;
  if ((ki & 0x80000000ull) == 0) {   // k > 0: the exponent of scale might have overflowed by <= 460
    sbits -= 1009ull << 52;
    scale = gm_asdouble(sbits);
    y = 0x1p1009 * fma(scale, tmp, scale);
    return y;
  }
  sbits += 1022ull << 52;            // k < 0: take care in the subnormal range
  scale = gm_asdouble(sbits);
  const double st = scale * tmp;
  y = scale + st;
  if (y < 1.0) {
    double hi, lo;
    lo = scale - y;
    lo = lo + st;
    hi = 1.0 + y;
    lo = ((1.0 - hi) + y) + lo;
    y = (hi + lo) - 1.0;
    if (y == 0.0) y = 0.0;
  }
  return 0x1p-1022 * y;
}

GM_SLOW double exp_glibc_full(double x, const Tables &T) {
  const unsigned long long ix = gm_asuint64(x);
  unsigned int abstop = (unsigned int)(ix >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int)(abstop - 0x3c9u) < 0) return 1.0 + x;   // |x| < 2^-54
    if (abstop > 0x408u) {                            // |x| >= 1024, inf, nan
      if (ix == 0xfff0000000000000ull) return 0.0;
      if (abstop == 0x7ffu) return 1.0 + x;
      if (ix >> 63) return 0x1p-767 * 0x1p-767;        // underflow
      return 0x1p769 * 0x1p769;                        // overflow
    }
    abstop = 0;   // 512 <= |x| < 1024: handled below with the special scaling
  }
  // only the 2^(j/128) table is read through the pointer
  const double InvLn2N = GM_EC(0), Shift = GM_EC(1), NegLn2hiN = GM_EC(2), NegLn2loN = GM_EC(3);
  const double C2 = GM_EC(4), C3 = GM_EC(5), C4 = GM_EC(6), C5 = GM_EC(7);
  double kd = fma(x, InvLn2N, Shift);
  const unsigned long long ki = gm_asuint64(kd);
  kd = kd - Shift;
  double r = fma(kd, NegLn2hiN, x);
  r = fma(kd, NegLn2loN, r);
  const unsigned int idx = 2u * (unsigned int)(ki & 0x7fu);
  const unsigned long long top = ki << 45;
  const double p1 = fma(r, C3, C2);
  const double tr = r + gm_asdouble(T.exp_tab[idx]);
  const unsigned long long sbits = T.exp_tab[idx + 1] + top;
  const double r2 = r * r;
  const double p2 = fma(r, C5, C4);
  const double t = fma(p1, r2, tr);
  const double r4 = r2 * r2;
  const double tmp = fma(r4, p2, t);
  if (abstop == 0) return exp_special(tmp, sbits, ki);
  const double scale = gm_asdouble(sbits);
  return fma(scale, tmp, scale);
}

GM_SLOW double log_glibc_full(double x, const Tables &T) {
  unsigned long long ix = gm_asuint64(x);
  if (ix + 0xc012000000000000ull <= 0x308ffffffffffull) {   // 1 - 0x1p-4 <= x < 1 + 0x1.09p-4
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double B0 = GM_LC(7), B1 = GM_LC(8), B2 = GM_LC(9), B3 = GM_LC(10), B4 = GM_LC(11), B5 = GM_LC(12),
                 B6 = GM_LC(13), B7 = GM_LC(14), B8 = GM_LC(15), B9 = GM_LC(16), B10 = GM_LC(17);
    const double r = x - 1.0;
    double q1 = fma(r, B2, B1);
    double q2 = fma(r, B5, B4);
    const double r2 = r * r;
    const double q3 = fma(r, B8, B7);
    q1 = fma(r2, B3, q1);
    q2 = fma(r2, B6, q2);
    const double r3 = r * r2;
    double q4 = fma(r2, B9, q3);
    q4 = fma(r3, B10, q4);
    const double q5 = fma(q4, r3, q2);
    const double q6 = fma(q5, r3, q1);
    const double t1 = fma(r, 0x1p27, r);
    const double rhi = fma(-0x1p27, r, t1);
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = fma(rhi2, B0, r);
    const double d = r - hi;
    const double rs = r + rhi;
    double lo = fma(rhi2, B0, d);
    const double b0rlo = B0 * rlo;
    lo = fma(b0rlo, rs, lo);
    const double y = fma(q6, r3, lo);
    return y + hi;
  }
  unsigned int top = (unsigned int)(ix >> 48);
  if (top - 0x10u > 0x7fdfu) {   // x < 0x1p-1022 or inf or nan
    if (ix * 2 == 0) return -gm_asdouble(0x7ff0000000000000ull);             // log(+-0) = -inf
    if (ix == 0x7ff0000000000000ull) return x;                                // log(inf) = inf
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return (x - x) / 0.0; // negative or nan
    ix = gm_asuint64(x * 0x1p52);   // subnormal: normalise
    ix -= 52ull << 52;
  }
  const double ln2hi = GM_LC(0), ln2lo = GM_LC(1), A0 = GM_LC(2), A1 = GM_LC(3), A2 = GM_LC(4), A3 = GM_LC(5), A4 = GM_LC(6);
  const unsigned long long tmp = ix + 0xc01a000000000000ull;   // ix - 0x3fe6000000000000
  const unsigned int i = (unsigned int)(tmp >> 45) & 0x7fu;
  const int k = (int)((long long)tmp >> 52);
  const unsigned long long iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = gm_asdouble(T.log_tab[2 * i]), logc = gm_asdouble(T.log_tab[2 * i + 1]);
  const double z = gm_asdouble(iz);
  const double kd = (double)k;
  const double w = fma(kd, ln2hi, logc);
  const double r = fma(z, invc, -1.0);
  const double pA = fma(r, A2, A1);
  const double hi = r + w;
  const double r2 = r * r;
  double lo = w - hi;
  lo = lo + r;
  lo = fma(kd, ln2lo, lo);
  const double r3 = r * r2;
  const double pB = fma(r, A4, A3);
  lo = fma(r2, A0, lo);
  const double pC = fma(pB, r2, pA);
  const double y0 = fma(r3, pC, lo);
  return y0 + hi;
}

// ---- the entry points post_sw calls ---------------------------------------------------------------------------------
// The transcriptions above (exp_glibc_full / log_glibc_full) branch where libm does: exp returns 1 + x for |x| < 2^-54,
// log returns 0 for x = 1 and takes a different polynomial near 1.  In post_sw those branches diverge in EVERY warp and
// every column (the node that carries the scale has forwards = 0, so one lane computes exp(-0); the sums 1 + tiny sit in
// log's near-1 interval, the others do not).  The entry points below give the same bits without diverging:
//  * exp: |x| < 2^-54 goes through the main path, which returns 1.0 there as libm's shortcut 1.0 + x does (round to
//    nearest: kd = Shift exactly, r = x, tmp = x (1 + O(x)), and fma(1, tmp, 1) rounds to 1.0 like 1 + x);
//    |x| >= 512, infinities and NaN (the +infinity forwards of the nodes that cannot start a read) leave through one
//    rarely taken call of the full transcription;
//  * log: both polynomials are evaluated and the result selected; x = 1 gives +0 through the near-1 polynomial (every
//    product is a zero and the final sums are +0 + +-0 = +0); zero, subnormal, negative, infinite and NaN arguments
//    leave through the full transcription.
// tools/check_glibc_math.c compares THESE with the host's libm.
GM_HD double exp_glibc(double x, const Tables &T) {
  const unsigned long long ix = gm_asuint64(x);
  const unsigned int abstop = (unsigned int)(ix >> 52) & 0x7ffu;
  if (abstop > 0x407u) return exp_glibc_full(x, T);   // |x| >= 512, inf, nan
  const double InvLn2N = GM_EC(0), Shift = GM_EC(1), NegLn2hiN = GM_EC(2), NegLn2loN = GM_EC(3);
  const double C2 = GM_EC(4), C3 = GM_EC(5), C4 = GM_EC(6), C5 = GM_EC(7);
  double kd = fma(x, InvLn2N, Shift);
  const unsigned long long ki = gm_asuint64(kd);
  kd = kd - Shift;
  double r = fma(kd, NegLn2hiN, x);
  r = fma(kd, NegLn2loN, r);
  const unsigned int idx = 2u * (unsigned int)(ki & 0x7fu);
  const unsigned long long top = ki << 45;
  const double p1 = fma(r, C3, C2);
  const double tr = r + gm_asdouble(T.exp_tab[idx]);
  const unsigned long long sbits = T.exp_tab[idx + 1] + top;
  const double r2 = r * r;
  const double p2 = fma(r, C5, C4);
  const double t = fma(p1, r2, tr);
  const double r4 = r2 * r2;
  const double tmp = fma(r4, p2, t);
  const double scale = gm_asdouble(sbits);
  return fma(scale, tmp, scale);
}

GM_HD double log_glibc(double x, const Tables &T) {
  const unsigned long long ix = gm_asuint64(x);
  const unsigned int top = (unsigned int)(ix >> 48);
  if (top - 0x10u > 0x7fdfu) return log_glibc_full(x, T);   // x < 0x1p-1022 (zero, subnormal, negative), inf, nan
  // near 1
  double near;
  {
    const double B0 = GM_LC(7), B1 = GM_LC(8), B2 = GM_LC(9), B3 = GM_LC(10), B4 = GM_LC(11), B5 = GM_LC(12),
                 B6 = GM_LC(13), B7 = GM_LC(14), B8 = GM_LC(15), B9 = GM_LC(16), B10 = GM_LC(17);
    const double r = x - 1.0;
    double q1 = fma(r, B2, B1);
    double q2 = fma(r, B5, B4);
    const double r2 = r * r;
    const double q3 = fma(r, B8, B7);
    q1 = fma(r2, B3, q1);
    q2 = fma(r2, B6, q2);
    const double r3 = r * r2;
    double q4 = fma(r2, B9, q3);
    q4 = fma(r3, B10, q4);
    const double q5 = fma(q4, r3, q2);
    const double q6 = fma(q5, r3, q1);
    const double t1 = fma(r, 0x1p27, r);
    const double rhi = fma(-0x1p27, r, t1);
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = fma(rhi2, B0, r);
    const double d = r - hi;
    const double rs = r + rhi;
    double lo = fma(rhi2, B0, d);
    const double b0rlo = B0 * rlo;
    lo = fma(b0rlo, rs, lo);
    const double y = fma(q6, r3, lo);
    near = y + hi;
  }
  double far;
  {
    const double ln2hi = GM_LC(0), ln2lo = GM_LC(1), A0 = GM_LC(2), A1 = GM_LC(3), A2 = GM_LC(4), A3 = GM_LC(5), A4 = GM_LC(6);
    const unsigned long long tmp = ix + 0xc01a000000000000ull;   // ix - 0x3fe6000000000000
    const unsigned int i = (unsigned int)(tmp >> 45) & 0x7fu;
    const int k = (int)((long long)tmp >> 52);
    const unsigned long long iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = gm_asdouble(T.log_tab[2 * i]), logc = gm_asdouble(T.log_tab[2 * i + 1]);
    const double z = gm_asdouble(iz);
    const double kd = (double)k;
    const double w = fma(kd, ln2hi, logc);
    const double r = fma(z, invc, -1.0);
    const double pA = fma(r, A2, A1);
    const double hi = r + w;
    const double r2 = r * r;
    double lo = w - hi;
    lo = lo + r;
    lo = fma(kd, ln2lo, lo);
    const double r3 = r * r2;
    const double pB = fma(r, A4, A3);
    lo = fma(r2, A0, lo);
    const double pC = fma(pB, r2, pA);
    const double y0 = fma(r3, pC, lo);
    far = y0 + hi;
  }
  return (ix + 0xc012000000000000ull <= 0x308ffffffffffull) ? near : far;   // 1 - 0x1p-4 <= x < 1 + 0x1.09p-4
}

}  // namespace glibc_math
