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

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define PF_HD __host__ __device__ __forceinline__
#else
#define PF_HD static inline
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
#if defined(__CUDA_ARCH__)
  *lo = a * b;
  *hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *lo = (uint32_t)p;
  *hi = (uint32_t)(p >> 32);
#endif
}

PF_HD pf_u32x4 pf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
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

/* 53-bit uniform in [0,1) from two 32-bit words: k = (hi>>5)*2^26 + (lo>>6).
 * Every step is exact in binary64, so host and device agree trivially. */
PF_HD double pf_u53(uint32_t hi, uint32_t lo) {
  double a = (double)(hi >> 5);
  double b = (double)(lo >> 6);
  return (a * 67108864.0 + b) * 1.1102230246251565e-16; /* 2^26, 2^-53 */
}
/* 53-bit uniform in (0,1]: (k+1) * 2^-53 (argument of the Box-Muller log) */
PF_HD double pf_u53_open0(uint32_t hi, uint32_t lo) {
  double a = (double)(hi >> 5);
  double b = (double)(lo >> 6);
  return (a * 67108864.0 + (b + 1.0)) * 1.1102230246251565e-16;
}

/* ------------------------------------------------------------ exp / log */

/* x * 2^k for |k| <= 2098, one rounding (only when the result is subnormal) */
PF_HD double pf_scalbn2(double x, int k) {
  int k1 = k >> 1;           /* arithmetic shift: floor(k/2) */
  int k2 = k - k1;
  double s1 = pf_u2d((uint64_t)(k1 + 1023) << 52);
  double s2 = pf_u2d((uint64_t)(k2 + 1023) << 52);
  return (x * s1) * s2;
}

/* exp(x): Cody-Waite reduction by ln2 (hi/lo split), degree-13 Taylor kernel on
 * |r| <= ln2/2 (truncation 4e-18), exact power-of-two scaling. */
PF_HD double pf_exp(double x) {
  if (!(x == x)) return x;
  if (x > 709.782712893384) return pf_inf();
  if (x < -745.1332191019412) return 0.0;
  const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52 */
  double t = pf_fma(x, 1.4426950408889634, MAGIC);
  double kf = t - MAGIC;
  int k = (int)(int32_t)(uint32_t)pf_d2u(t);
  double r = pf_fma(-kf, 6.93147180369123816490e-01, x);
  r = pf_fma(-kf, 1.90821492927058770002e-10, r);
  double q = 1.6059043836821613e-10;            /* 1/13! */
  q = pf_fma(q, r, 2.08767569878681e-09);       /* 1/12! */
  q = pf_fma(q, r, 2.505210838544172e-08);      /* 1/11! */
  q = pf_fma(q, r, 2.755731922398589e-07);      /* 1/10! */
  q = pf_fma(q, r, 2.7557319223985893e-06);     /* 1/9!  */
  q = pf_fma(q, r, 2.48015873015873e-05);       /* 1/8!  */
  q = pf_fma(q, r, 1.984126984126984e-04);      /* 1/7!  */
  q = pf_fma(q, r, 1.388888888888889e-03);      /* 1/6!  */
  q = pf_fma(q, r, 8.333333333333333e-03);      /* 1/5!  */
  q = pf_fma(q, r, 4.1666666666666664e-02);     /* 1/4!  */
  q = pf_fma(q, r, 1.6666666666666666e-01);     /* 1/3!  */
  q = pf_fma(q, r, 0.5);
  double e = pf_fma(q, r * r, r);
  return pf_scalbn2(1.0 + e, k);
}

/* log(x): x = 2^e * m, m in (sqrt(1/2), sqrt(2)], s = f/(2+f) with f = m-1,
 * log(m) = 2 atanh(s) = f - (f^2/2 - s*(f^2/2 + R(s^2))),
 * R(z) = sum_{j=1..10} 2 z^j/(2j+1) (Taylor; truncation 6e-19 relative). */
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
  e += (int)(ux >> 52) - 1023;
  double m = pf_u2d((ux & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double R = 9.523809523809523e-02;            /* 2/21 */
  R = pf_fma(R, z, 1.0526315789473684e-01);    /* 2/19 */
  R = pf_fma(R, z, 1.1764705882352941e-01);    /* 2/17 */
  R = pf_fma(R, z, 1.3333333333333333e-01);    /* 2/15 */
  R = pf_fma(R, z, 1.5384615384615385e-01);    /* 2/13 */
  R = pf_fma(R, z, 1.8181818181818182e-01);    /* 2/11 */
  R = pf_fma(R, z, 2.2222222222222222e-01);    /* 2/9  */
  R = pf_fma(R, z, 2.857142857142857e-01);     /* 2/7  */
  R = pf_fma(R, z, 4.0e-01);                   /* 2/5  */
  R = pf_fma(R, z, 6.666666666666666e-01);     /* 2/3  */
  R = R * z;
  double hfsq = 0.5 * f * f;
  double dk = (double)e;
  /* ln2 split: hi has 21 trailing zero bits, so dk*hi is exact */
  return dk * 6.93147180369123816490e-01 -
         ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
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
PF_HD void pf_sincos2pi(double u, double* sn, double* cs) {
  double x4 = u * 4.0;               /* exact */
  double q = floor(x4 + 0.5);        /* nearest quadrant boundary, 0..4 */
  double f = x4 - q;                 /* exact, in [-0.5, 0.5] */
  double a = f * 1.5707963267948966; /* pi/2 */
  double a2 = a * a;
  double ps = -7.647163731819816e-13;          /* -1/15! */
  ps = pf_fma(ps, a2, 1.6059043836821613e-10); /*  1/13! */
  ps = pf_fma(ps, a2, -2.505210838544172e-08); /* -1/11! */
  ps = pf_fma(ps, a2, 2.7557319223985893e-06); /*  1/9!  */
  ps = pf_fma(ps, a2, -1.984126984126984e-04); /* -1/7!  */
  ps = pf_fma(ps, a2, 8.333333333333333e-03);  /*  1/5!  */
  ps = pf_fma(ps, a2, -1.6666666666666666e-01);/* -1/3!  */
  double s = pf_fma(ps * a2, a, a);
  double pc = 4.779477332387385e-14;           /*  1/16! */
  pc = pf_fma(pc, a2, -1.1470745597729725e-11);/* -1/14! */
  pc = pf_fma(pc, a2, 2.08767569878681e-09);   /*  1/12! */
  pc = pf_fma(pc, a2, -2.755731922398589e-07); /* -1/10! */
  pc = pf_fma(pc, a2, 2.48015873015873e-05);   /*  1/8!  */
  pc = pf_fma(pc, a2, -1.388888888888889e-03); /* -1/6!  */
  pc = pf_fma(pc, a2, 4.1666666666666664e-02); /*  1/4!  */
  pc = pf_fma(pc, a2, -0.5);
  double c = pf_fma(pc, a2, 1.0);
  int qi = ((int)q) & 3;
  double so = (qi & 1) ? c : s;
  double co = (qi & 1) ? s : c;
  if (qi == 2 || qi == 3) so = -so;
  if (qi == 1 || qi == 2) co = -co;
  *sn = so; *cs = co;
}

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
enum { PF_RNG_NORMAL = 0, PF_RNG_RESAMPLE = 1, PF_RNG_PICK = 2 };

/* Box-Muller pair for flat normal index pair `pair` */
PF_HD void pf_normal_pair(uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub,
                          uint32_t pair, double* z0, double* z1) {
  pf_u32x4 x = pf_philox4x32_10(pair, sub, t, filter, seed, (uint32_t)PF_RNG_NORMAL);
  double u1 = pf_u53_open0(x.v[0], x.v[1]);
  double u2 = pf_u53(x.v[2], x.v[3]);
  double rad = sqrt(-2.0 * pf_log(u1));
  double s, c;
  pf_sincos2pi(u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

/* single normal with flat index f (convenience; evaluates the whole pair) */
PF_HD double pf_normal(uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub, uint32_t f) {
  double z0, z1;
  pf_normal_pair(seed, filter, t, sub, f >> 1, &z0, &z1);
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
 * depends on floating-point summation order:
 *
 *  (1) sum of the unnormalised weights w_i = exp(logw_i - max) in [0,1]
 *      (arma::accu, src/model_ssm_ung.cpp:555): each w is split exactly into
 *      hi = floor(w*2^40) and lo = floor((w*2^40 - hi)*2^43); the two integer
 *      sums are associative, and the total is rounded to binary64 once at the
 *      end.  Absolute truncation 2^-83 per term: the result is the correctly
 *      rounded true sum except in astronomically rare double-rounding cases,
 *      and it is the same for ANY reduction tree.
 *  (2) cumulative sum of the normalised weights W_i = w_i / sum
 *      (arma::cumsum, src/stratified_sample.cpp:12): Q_i = floor(W_i*2^62),
 *      integer inclusive scan, p_k = (double)(sum_{i<=k} Q_i) * 2^-62.
 *      Monotone by construction; truncation 2^-62 per term, i.e. closer to
 *      the real-number cumsum than a sequential binary64 cumsum is.
 *
 * The oracle implements both this mode and the reference's verbatim sequential
 * binary64 mode and reports where they differ (oracle/pf_oracle.cpp).
 */
typedef struct { int64_t hi, lo; } pf_fixsum;

PF_HD pf_fixsum pf_fixsum_term(double w) { /* w in [0,1] (NaN -> flagged by caller) */
  double sc = w * 1099511627776.0;             /* 2^40, exact */
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
PF_HD uint64_t pf_cumsum_term(double W) { /* W in [0,1] */
  return (uint64_t)(W * 4611686018427387904.0); /* 2^62, truncation */
}
PF_HD double pf_cumsum_value(uint64_t q) {
  return (double)q * 2.168404344971009e-19; /* 2^-62 */
}

#endif /* BSSM_B200_PF_MATH_H */
