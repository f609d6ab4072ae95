/* pf_math.h -- deterministic scalar building blocks shared by the CUDA kernels
 * and by the CPU oracle: counter-based Philox4x32-10, the u32 -> double
 * conversions, Box-Muller normals, and exp/log/log1p/sin-cos(2*pi*u) written
 * only with IEEE-754 +, *, /, sqrt and explicit fma().
 *
 * Why this exists: the reference draws its random numbers
 * from sitmo::prng_engine + libstdc++ <random> and evaluates weights with glibc
 * libm (src/model_ssm_ung.cpp:396-429, :536-546, :620-629).  Neither is
 * reproducible on a GPU.  The north star prescribes a counter-based Philox
 * stream; "identical seeds" therefore means: the oracle and the kernels consume
 * the stream defined HERE, and evaluate transcendentals with the functions
 * defined HERE, so that host (gcc -ffp-contract=off) and device (nvcc
 * -fmad=false) results are bit-identical.  Accuracy of each function against
 * glibc is checked in tests/test_pf_math.py (<= 4 ulp), so log-likelihoods stay
 * far inside the 1e-6 relative tolerance against what libm would give.
 *
 * This header is part of the public boundary: the random stream layout below
 * is the contract that makes results independent of launch geometry and of the
 * number of GPUs.
 */
#ifndef BSSM_B200_PF_MATH_H
#define BSSM_B200_PF_MATH_H

#if defined(__CUDACC_RTC__)
/* run-time compilation of user model functors (NVRTC, bssm_b200/csrc/user_models.cu): no host headers */
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#include <string.h>
#include <math.h>
#endif

#if defined(__CUDACC__)
#define PF_HD __host__ __device__ __forceinline__
#else
#define PF_HD static inline
#endif
/* The heavy kernels (exp, log, sincos, Philox, the Box-Muller pair) can be compiled as real
 * functions on the device (-DPF_NOINLINE_MATH=1): the unrolled per-particle copies then share one
 * body, which shrinks the instruction footprint of the filter kernel (experiment switch). */
#if defined(__CUDACC__) && defined(PF_NOINLINE_MATH) && PF_NOINLINE_MATH
#define PF_HDX static __host__ __device__ __noinline__
#else
#define PF_HDX PF_HD
#endif

/* ------------------------------------------------------------------ bits */

PF_HD uint64_t pf_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PF_HD double pf_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
PF_HD double pf_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
PF_HD double pf_inf(void) { return pf_u2d(0x7ff0000000000000ULL); }
PF_HD double pf_nan(void) { return pf_u2d(0x7ff8000000000000ULL); }
PF_HD int pf_isfinite(double x) {
  return ((pf_d2u(x) >> 52) & 0x7ff) != 0x7ff;
}

/* --------------------------------------------------------------- Philox */
/* Philox4x32-10 (Salmon et al., SC'11; Random123).  Known-answer vectors in
 * tests/test_pf_math.py.  Multipliers / Weyl constants are the published ones. */

typedef struct { uint32_t v[4]; } pf_u32x4;

PF_HD void pf_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  /* one 32x32->64 multiply (IMAD.WIDE.U32 on the device), not mul.lo + mul.hi */
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *lo = (uint32_t)p;
  *hi = (uint32_t)(p >> 32);
}

PF_HDX pf_u32x4 pf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    pf_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    pf_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  pf_u32x4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

/* (double)v for a 32-bit unsigned v.  On the device the conversion instructions (I2F / F2I /
 * FRND) run at 16 lanes per clock and SM, a quarter of the DFMA rate (measured, tools/bench_lat.cu);
 * the exponent trick -- v placed in the low mantissa bits of 2^52, minus 2^52 -- is one DADD and
 * gives the same value exactly.  Measured on the C2 batch it is a LOSS (32.3 -> 34.4 ms together
 * with the rounding-trick variants of floor / F2I below): the conversion pipe overlaps with the
 * FP64 pipe, which is the one that binds, so the trick stays behind -DPF_CONVERT_BY_DADD. */
PF_HD double pf_u32_to_f64(uint32_t v) {
#if defined(__CUDA_ARCH__) && defined(PF_CONVERT_BY_DADD)
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#else
  return (double)v;
#endif
}

/* (double)v for a signed 32-bit v, same trick with the sign bit flipped (bias 2^52 + 2^31) */
PF_HD double pf_i32_to_f64(int32_t v) {
#if defined(__CUDA_ARCH__) && defined(PF_CONVERT_BY_DADD)
  return __hiloint2double(0x43300000, v ^ (int)0x80000000) - 4503601774854144.0;
#else
  return (double)v;
#endif
}
/* round-to-nearest-even integer of x (|x| < 2^51) as a double and as a two's-complement integer,
 * by adding 1.5 * 2^52: one DADD each instead of FRND / F2I */
#define PF_RN_MAGIC 6755399441055744.0
PF_HD double pf_rn_magic(double x) { return x + PF_RN_MAGIC; }
PF_HD int64_t pf_rn_int64(double t) { return (int64_t)(pf_d2u(t) - 0x4338000000000000ULL); }
PF_HD int32_t pf_rn_int32(double t) { return (int32_t)(uint32_t)pf_d2u(t); }

/* 53-bit uniform in [0,1) from two 32-bit words: k = (hi>>5)*2^26 + (lo>>6).
 * Every step is exact in binary64, so host and device agree trivially. */
PF_HD double pf_u53(uint32_t hi, uint32_t lo) {
  double a = pf_u32_to_f64(hi >> 5);
  double b = pf_u32_to_f64(lo >> 6);
  return pf_fma(a, 67108864.0, b) * 1.1102230246251565e-16; /* 2^26, 2^-53; the fma is exact (k < 2^53) */
}
/* (u + j) for the stratum thresholds, u the 53-bit uniform of (hi, lo): one rounding, as pf_u53() + j */
PF_HD double pf_u53_plus(uint32_t hi, uint32_t lo, double j) {
  double a = pf_u32_to_f64(hi >> 5);
  double b = pf_u32_to_f64(lo >> 6);
  return pf_fma(pf_fma(a, 67108864.0, b), 1.1102230246251565e-16, j);
}
/* 53-bit uniform in (0,1]: (k+1) * 2^-53 (argument of the Box-Muller log) */
PF_HD double pf_u53_open0(uint32_t hi, uint32_t lo) {
  double a = pf_u32_to_f64(hi >> 5);
  double b = pf_u32_to_f64(lo >> 6);
  return pf_fma(a, 67108864.0, b + 1.0) * 1.1102230246251565e-16;
}

/* ------------------------------------------------------------ exp / log */

/* Polynomial coefficients live in tables: on the device they are __constant__
 * memory, so a DFMA takes its coefficient as a c[bank][offset] operand instead
 * of two UMOVs building a 64-bit immediate (that alone was 12% of the executed
 * instructions of the first filter kernel, profiles/r1_v0_filter_single.md). */
#define PF_EXP_COEFFS                                                             \
  1.6059043836821613e-10 /* 1/13! */, 2.08767569878681e-09 /* 1/12! */,           \
  2.505210838544172e-08 /* 1/11! */, 2.755731922398589e-07 /* 1/10! */,           \
  2.7557319223985893e-06 /* 1/9! */, 2.48015873015873e-05 /* 1/8! */,             \
  1.984126984126984e-04 /* 1/7! */, 1.388888888888889e-03 /* 1/6! */,             \
  8.333333333333333e-03 /* 1/5! */, 4.1666666666666664e-02 /* 1/4! */,            \
  1.6666666666666666e-01 /* 1/3! */, 0.5,                                         \
  1.4426950408889634 /* 12: 1/ln2 */, 6755399441055744.0 /* 13: 1.5*2^52 */,      \
  6.93147180369123816490e-01 /* 14: ln2 hi */, 1.90821492927058770002e-10 /* 15: ln2 lo */
#define PF_LOG_COEFFS                                                             \
  9.523809523809523e-02 /* 2/21 */, 1.0526315789473684e-01 /* 2/19 */,            \
  1.1764705882352941e-01 /* 2/17 */, 1.3333333333333333e-01 /* 2/15 */,           \
  1.5384615384615385e-01 /* 2/13 */, 1.8181818181818182e-01 /* 2/11 */,           \
  2.2222222222222222e-01 /* 2/9 */, 2.857142857142857e-01 /* 2/7 */,              \
  4.0e-01 /* 2/5 */, 6.666666666666666e-01 /* 2/3 */,                             \
  6.93147180369123816490e-01 /* 10: ln2 hi */, 1.90821492927058770002e-10 /* 11: ln2 lo */, \
  1.4142135623730951 /* 12: sqrt 2 */
#define PF_SIN_COEFFS                                                             \
  -7.647163731819816e-13 /* -1/15! */, 1.6059043836821613e-10 /* 1/13! */,        \
  -2.505210838544172e-08 /* -1/11! */, 2.7557319223985893e-06 /* 1/9! */,         \
  -1.984126984126984e-04 /* -1/7! */, 8.333333333333333e-03 /* 1/5! */,           \
  -1.6666666666666666e-01 /* -1/3! */, 1.5707963267948966 /* 7: pi/2 */
#define PF_COS_COEFFS                                                             \
  4.779477332387385e-14 /* 1/16! */, -1.1470745597729725e-11 /* -1/14! */,        \
  2.08767569878681e-09 /* 1/12! */, -2.755731922398589e-07 /* -1/10! */,          \
  2.48015873015873e-05 /* 1/8! */, -1.388888888888889e-03 /* -1/6! */,            \
  4.1666666666666664e-02 /* 1/4! */, -0.5
static const double pf_hc_exp[] = {PF_EXP_COEFFS};
static const double pf_hc_log[] = {PF_LOG_COEFFS};
static const double pf_hc_sin[] = {PF_SIN_COEFFS};
static const double pf_hc_cos[] = {PF_COS_COEFFS};
#if defined(__CUDACC__)
static __constant__ double pf_dc_exp[] = {PF_EXP_COEFFS};
static __constant__ double pf_dc_log[] = {PF_LOG_COEFFS};
static __constant__ double pf_dc_sin[] = {PF_SIN_COEFFS};
static __constant__ double pf_dc_cos[] = {PF_COS_COEFFS};
#endif
#if defined(__CUDA_ARCH__)
#define PF_CE pf_dc_exp
#define PF_CL pf_dc_log
#define PF_CS pf_dc_sin
#define PF_CC pf_dc_cos
#else
#define PF_CE pf_hc_exp
#define PF_CL pf_hc_log
#define PF_CS pf_hc_sin
#define PF_CC pf_hc_cos
#endif

/* y * 2^k as 2 * (y * 2^(k-1)), k in two's complement: add (k - 1) to the exponent field (only
 * the high word changes), then double */
PF_HD double pf_exp_scale(double y, uint32_t k) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(y) + (int)((k - 1u) << 20), __double2loint(y)) * 2.0;
#else
  uint64_t yb = pf_d2u(y);
  uint32_t hi = (uint32_t)(yb >> 32) + ((k - 1u) << 20);
  return pf_u2d(((uint64_t)hi << 32) | (yb & 0xffffffffULL)) * 2.0;
#endif
}

/* exp(x): Cody-Waite reduction by ln2 (hi/lo split), degree-13 Taylor kernel on
 * |r| <= ln2/2 (truncation 4e-18), power-of-two scaling.  Straight-line code:
 * y * 2^k is evaluated as 2 * (y * 2^(k-1)); the inner scaling is an integer add
 * on the exponent field (y in [0.70, 1.42] and k - 1 in [-1020, 1023] keep the
 * field inside [2, 2046]), the final doubling is exact and overflows to +inf where
 * exp does.  Results below 2^-1020 (x < -706.5) FLUSH TO ZERO: the reference's
 * std::exp returns subnormals there (< 1e-307; as a weight relative to the largest
 * weight of a column that is nothing).  NaN flows through. */
PF_HDX double pf_exp(double x) {
  double xc = x > 710.0 ? 710.0 : x;     /* NaN compares false: stays NaN */
  const double MAGIC = PF_CE[13];
  double t = pf_fma(xc, PF_CE[12], MAGIC);
  double kf = t - MAGIC;
  uint32_t k = (uint32_t)pf_d2u(t);      /* two's complement of round(x / ln2) */
  double r = pf_fma(-kf, PF_CE[14], xc);
  r = pf_fma(-kf, PF_CE[15], r);
  double q = PF_CE[0];
  q = pf_fma(q, r, PF_CE[1]);
  q = pf_fma(q, r, PF_CE[2]);
  q = pf_fma(q, r, PF_CE[3]);
  q = pf_fma(q, r, PF_CE[4]);
  q = pf_fma(q, r, PF_CE[5]);
  q = pf_fma(q, r, PF_CE[6]);
  q = pf_fma(q, r, PF_CE[7]);
  q = pf_fma(q, r, PF_CE[8]);
  q = pf_fma(q, r, PF_CE[9]);
  q = pf_fma(q, r, PF_CE[10]);
  q = pf_fma(q, r, PF_CE[11]);
  double e = pf_fma(q, r * r, r);
  double res = pf_exp_scale(1.0 + e, k);
  if (x < -706.5) res = 0.0;
  if (!(x == x)) res = x;
  return res;
}

/* The same for an argument that is known to be <= 0 and not NaN (the weights exp(logw - max)):
 * no upper clamp, no NaN guard; -inf gives 0 through the flush. */
PF_HDX double pf_exp_nonpos(double x) {
  const double MAGIC = PF_CE[13];
  double t = pf_fma(x, PF_CE[12], MAGIC);
  double kf = t - MAGIC;
  uint32_t k = (uint32_t)pf_d2u(t);
  double r = pf_fma(-kf, PF_CE[14], x);
  r = pf_fma(-kf, PF_CE[15], r);
  double q = PF_CE[0];
  q = pf_fma(q, r, PF_CE[1]);
  q = pf_fma(q, r, PF_CE[2]);
  q = pf_fma(q, r, PF_CE[3]);
  q = pf_fma(q, r, PF_CE[4]);
  q = pf_fma(q, r, PF_CE[5]);
  q = pf_fma(q, r, PF_CE[6]);
  q = pf_fma(q, r, PF_CE[7]);
  q = pf_fma(q, r, PF_CE[8]);
  q = pf_fma(q, r, PF_CE[9]);
  q = pf_fma(q, r, PF_CE[10]);
  q = pf_fma(q, r, PF_CE[11]);
  double e = pf_fma(q, r * r, r);
  double res = pf_exp_scale(1.0 + e, k);
  return x < -706.5 ? 0.0 : res;
}

/* log(x): x = 2^e * m, m in (sqrt(1/2), sqrt(2)], s = f/(2+f) with f = m-1,
 * log(m) = 2 atanh(s) = f - (f^2/2 - s*(f^2/2 + R(s^2))),
 * R(z) = sum_{j=1..10} 2 z^j/(2j+1) (Taylor; truncation 6e-19 relative).
 * pf_log_normal is the kernel for positive NORMAL finite x (no special cases):
 * Box-Muller's u1 in [2^-53, 1] and the log(sum/N) of the filter need no more. */
PF_HDX double pf_log_core(uint64_t ux, int e) {
  e += (int)(ux >> 52) - 1023;
  double m = pf_u2d((ux & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > PF_CL[12]) { m = m * 0.5; e += 1; }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double R = PF_CL[0];
  R = pf_fma(R, z, PF_CL[1]);
  R = pf_fma(R, z, PF_CL[2]);
  R = pf_fma(R, z, PF_CL[3]);
  R = pf_fma(R, z, PF_CL[4]);
  R = pf_fma(R, z, PF_CL[5]);
  R = pf_fma(R, z, PF_CL[6]);
  R = pf_fma(R, z, PF_CL[7]);
  R = pf_fma(R, z, PF_CL[8]);
  R = pf_fma(R, z, PF_CL[9]);
  R = R * z;
  double hfsq = 0.5 * f * f;
  double dk = pf_i32_to_f64(e);
  /* ln2 split: hi has 21 trailing zero bits, so dk*hi is exact */
  return dk * PF_CL[10] - ((hfsq - (s * (hfsq + R) + dk * PF_CL[11])) - f);
}
PF_HD double pf_log_normal(double x) { return pf_log_core(pf_d2u(x), 0); }
PF_HD double pf_log(double x) {
  if (!(x == x)) return x;
  if (x < 0.0) return pf_nan();
  if (x == 0.0) return -pf_inf();
  uint64_t ux = pf_d2u(x);
  if ((ux >> 52) == 0x7ff) return x; /* +inf */
  int e = 0;
  if ((ux >> 52) == 0) { /* subnormal: scale by 2^54 (exact) */
    x = x * 18014398509481984.0;
    ux = pf_d2u(x);
    e = -54;
  }
  return pf_log_core(ux, e);
}

/* log(1+x) for x >= 0 (the only use: binomial log1p(exp(signal))) */
PF_HD double pf_log1p(double x) {
  double w = 1.0 + x;
  if (w == 1.0) return x;
  if (!pf_isfinite(w)) return pf_log(w);
  return pf_log(w) * (x / (w - 1.0));
}

/* sin(2*pi*u), cos(2*pi*u) for u in [0,1): exact quadrant reduction, Taylor
 * kernels on |a| <= pi/4. */
PF_HDX void pf_sincos2pi(double u, double* sn, double* cs) {
  double x4 = u * 4.0;               /* exact */
  double q = floor(x4 + 0.5);        /* nearest quadrant boundary, 0..4 */
  double f = x4 - q;                 /* exact, in [-0.5, 0.5] */
  double a = f * PF_CS[7]; /* pi/2 */
  double a2 = a * a;
  double ps = PF_CS[0];
  ps = pf_fma(ps, a2, PF_CS[1]);
  ps = pf_fma(ps, a2, PF_CS[2]);
  ps = pf_fma(ps, a2, PF_CS[3]);
  ps = pf_fma(ps, a2, PF_CS[4]);
  ps = pf_fma(ps, a2, PF_CS[5]);
  ps = pf_fma(ps, a2, PF_CS[6]);
  double s = pf_fma(ps * a2, a, a);
  double pc = PF_CC[0];
  pc = pf_fma(pc, a2, PF_CC[1]);
  pc = pf_fma(pc, a2, PF_CC[2]);
  pc = pf_fma(pc, a2, PF_CC[3]);
  pc = pf_fma(pc, a2, PF_CC[4]);
  pc = pf_fma(pc, a2, PF_CC[5]);
  pc = pf_fma(pc, a2, PF_CC[6]);
  pc = pf_fma(pc, a2, PF_CC[7]);
  double c = pf_fma(pc, a2, 1.0);
  int qi = ((int)q) & 3;
  double so = (qi & 1) ? c : s;
  double co = (qi & 1) ? s : c;
  if (qi == 2 || qi == 3) so = -so;
  if (qi == 1 || qi == 2) co = -co;
  *sn = so; *cs = co;
}

/* sin(x), cos(x) for moderate |x| (state values): one multiplication by 1/(2 pi), the exact
 * fractional part, then the kernel above.  The reduction error is |x| * 2^-53 in the angle. */
PF_HD void pf_sincos(double x, double* sn, double* cs) {
  double t = x * 0.15915494309189535; /* 1 / (2 pi) */
  double u = t - floor(t);            /* [0, 1]; 1 only by rounding, which the kernel maps onto 0 */
  pf_sincos2pi(u, sn, cs);
}
PF_HD double pf_sin(double x) { double s, c; pf_sincos(x, &s, &c); return s; }
PF_HD double pf_cos(double x) { double s, c; pf_sincos(x, &s, &c); return c; }

/* ----------------------------------------------------- the random stream */
/* Stream contract (key = (seed, purpose); counter = (index, sub, t, filter)):
 *
 *  PF_RNG_NORMAL   index = (i*K + d) >> 1 where i = particle, d = noise
 *                  component, K = normals per particle at that time (m for the
 *                  initial draw and for psi-APF steps, k for BSF steps, 1 for
 *                  SDE sub-steps).  The flat index i*K+d is the reference's
 *                  own consumption order (particle-major, component-minor:
 *                  src/model_ssm_ung.cpp:536-542, :663-667).  One Philox call
 *                  gives one Box-Muller pair; flat index even -> cos branch,
 *                  odd -> sin branch.  t = time column the draw produces
 *                  (0 = initial state).  sub = Milstein sub-step (else 0).
 *  PF_RNG_RESAMPLE index = j >> 1, j = stratum; one call gives two 53-bit
 *                  uniforms, (v0,v1) for even j and (v2,v3) for odd j.
 *                  t = column of `indices` being produced.
 *  PF_RNG_PICK     one uniform per filter for drawing a single trajectory
 *                  (post_correct output_type = 1, src/approx_mcmc.cpp:391).
 *  filter = replicate / MCMC-draw index, so every filter of a batch has its
 *  own stream whatever GPU or CTA runs it.
 */
enum { PF_RNG_NORMAL = 0, PF_RNG_RESAMPLE = 1, PF_RNG_PICK = 2, PF_RNG_SIM = 3 };

/* Box-Muller pair for flat normal index pair `pair` of the stream with key (seed, purpose) */
PF_HD void pf_normal_pair_of(uint32_t purpose, uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub,
                             uint32_t pair, double* z0, double* z1) {
  pf_u32x4 x = pf_philox4x32_10(pair, sub, t, filter, seed, purpose);
  double u1 = pf_u53_open0(x.v[0], x.v[1]);
  double u2 = pf_u53(x.v[2], x.v[3]);
  double rad = sqrt(-2.0 * pf_log_normal(u1));
  double s, c;
  pf_sincos2pi(u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

PF_HD void pf_normal_pair(uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub,
                          uint32_t pair, double* z0, double* z1) {
  pf_normal_pair_of((uint32_t)PF_RNG_NORMAL, seed, filter, t, sub, pair, z0, z1);
}

/* single normal with flat index f (convenience; evaluates the whole pair) */
PF_HD double pf_normal(uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub, uint32_t f) {
  double z0, z1;
  pf_normal_pair(seed, filter, t, sub, f >> 1, &z0, &z1);
  return (f & 1u) ? z1 : z0;
}

/* PF_RNG_SIM: the normals of the simulation smoother (ssm_ulg::simulate_states, src/model_ssm_ulg.cpp:141-247),
 * in the reference's consumption order per simulated trajectory i: column t = 0 takes m normals (the initial
 * state, flat index i*m + d), column t >= 1 takes K = k + 1 (flat index i*K + d: d = 0 the observation noise of
 * y_{t-1}, reserved also where y is missing, d = 1..k the state noise). */
PF_HD double pf_sim_normal(uint32_t seed, uint32_t filter, uint32_t t, uint32_t f) {
  double z0, z1;
  pf_normal_pair_of((uint32_t)PF_RNG_SIM, seed, filter, t, 0u, f >> 1, &z0, &z1);
  return (f & 1u) ? z1 : z0;
}

/* resampling uniform of stratum j at step t */
PF_HD double pf_resample_uniform(uint32_t seed, uint32_t filter, uint32_t t, uint32_t j) {
  pf_u32x4 x = pf_philox4x32_10(j >> 1, 0u, t, filter, seed, (uint32_t)PF_RNG_RESAMPLE);
  return (j & 1u) ? pf_u53(x.v[2], x.v[3]) : pf_u53(x.v[0], x.v[1]);
}

PF_HD double pf_pick_uniform(uint32_t seed, uint32_t filter) {
  pf_u32x4 x = pf_philox4x32_10(0u, 0u, 0u, filter, seed, (uint32_t)PF_RNG_PICK);
  return pf_u53(x.v[0], x.v[1]);
}

/* -------------------------------------------- reproducible sum / cumsum */
/* Order-independent arithmetic for the two places where the reference's result
 * depends on floating-point summation order: the sum of the
 * unnormalised weights w_i = exp(logw_i - max) in [0,1] (arma::accu,
 * src/model_ssm_ung.cpp:555) and the cumulative sum of the normalised weights
 * (arma::cumsum, src/stratified_sample.cpp:12).
 *
 * Each w is split exactly into two integers, hi = floor(w*2^40) and
 * lo = floor((w*2^40 - hi)*2^43) (truncation 2^-83 absolute).  Integer sums are
 * associative, so ONE inclusive scan of the (hi, lo) pairs gives, for any
 * reduction tree, block size or number of CTAs,
 *     C_k = value(sum_{i<=k} hi_i, sum_{i<=k} lo_i),   S = C_{N-1},
 *     p_k = C_k * (1/S),  p_{N-1} := 1   (the reference forces p[last] = 1 too)
 * with value(H, L) = fl(H*2^-40 + L*2^-83) rounded once.  p is monotone by
 * construction.  Against the reference's p_k = sum_{i<=k} fl(w_i/S) (sequential
 * binary64) the difference is O(2^-53) relative -- the same size as the
 * reference's own rounding error -- so ancestors differ only when a stratum
 * threshold falls within ~1e-16 of a cumulative weight.  The oracle implements
 * both this mode and the verbatim sequential mode and tests compare them
 * (oracle/pf_oracle.cpp, tests/test_oracle.py). */
typedef struct { int64_t hi, lo; } pf_fixsum;

/* floor() and the conversions run on the conversion pipe (16 lanes per clock and SM), which overlaps
 * with the FP64 pipe; a variant that rounds by adding 1.5 * 2^52 (no conversion instruction, three
 * more DADD) was measured slower on the C2 batch because the FP64 pipe is the one that binds. */
PF_HD pf_fixsum pf_fixsum_term(double w) { /* w in [0,1] (NaN is caught before) */
  double sc = w * 1099511627776.0;              /* 2^40, exact */
  double h = floor(sc);
  double l = floor((sc - h) * 8796093022208.0); /* 2^43; sc-h exact */
  pf_fixsum r;
  r.hi = (int64_t)h;
  r.lo = (int64_t)l;
  return r;
}
PF_HD double pf_fixsum_value(int64_t hi, int64_t lo) {
  /* hi*2^-40 + lo*2^-83; conversions round-to-nearest on both sides */
  return pf_fma((double)lo, 1.0339757656912846e-25, (double)hi * 9.094947017729282e-13);
}

#endif /* BSSM_B200_PF_MATH_H */
