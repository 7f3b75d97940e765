/* pimdk_detmath.h — deterministic FP64 elementary functions (the "math policy").
 *
 * Why this exists: the reference's CCpol gradient is a central finite difference of the energy
 * with eps = 1e-4 bohr (mcmod_waterdimer_ccpol.f90:40-58).  That amplifies every last-bit
 * difference of V by ~5000/|grad|: two correct builds of the *same* Fortran that differ only in
 * FMA contraction or in whose libm supplies exp() disagree at ~1e-9 relative in the gradient
 * (measured).  The 1e-10 contract is therefore only meaningful if the CPU
 * oracle and the CUDA kernels evaluate V with bit-identical arithmetic.  IEEE-754 fixes
 * + - * / sqrt and fma; it does not fix exp, log, pow, sin, cos, acos, tanh — so both sides take
 * those from THIS header, written only in terms of IEEE operations, explicit fma() and integer
 * bit manipulation.  Every function is within ~1-2 ulp of the correctly rounded result on the
 * argument ranges the hot path uses (tests/test_detmath.py measures it against libm/libdevice),
 * i.e. it is as good a stand-in for the reference's unknown Intel libm as glibc is.
 *
 * Rules that keep host and device identical:
 *   - no reliance on compiler contraction: every multiply-add that may fuse is an explicit
 *     PIMDK_FMA; every other product/sum goes through PIMDK_MUL/PIMDK_ADD, which map to
 *     __dmul_rn/__dadd_rn on the device (never contracted) and to plain operators on the host
 *     (the oracle is built with -ffp-contract=off);
 *   - round-to-nearest-even throughout; no denormal special-casing is needed on the hot path
 *     (arguments of exp lie in about [-200, 50]).
 */
#ifndef PIMDK_DETMATH_H
#define PIMDK_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PIMDK_HD __host__ __device__ __forceinline__
#else
#define PIMDK_HD static inline
#include <math.h>
#endif

#if defined(__CUDA_ARCH__)
#define PIMDK_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PIMDK_MUL(a, b) __dmul_rn((a), (b))
#define PIMDK_ADD(a, b) __dadd_rn((a), (b))
#define PIMDK_SUB(a, b) __dsub_rn((a), (b))
#define PIMDK_DIV(a, b) __ddiv_rn((a), (b))
#define PIMDK_SQRT(a) __dsqrt_rn((a))
#else
#define PIMDK_FMA(a, b, c) fma((a), (b), (c))
#define PIMDK_MUL(a, b) ((a) * (b))
#define PIMDK_ADD(a, b) ((a) + (b))
#define PIMDK_SUB(a, b) ((a) - (b))
#define PIMDK_DIV(a, b) ((a) / (b))
#define PIMDK_SQRT(a) sqrt((a))
#endif


/* Polynomial coefficient tables.  On the device they live in constant memory: the compiler then
 * fetches them with wide uniform loads (LDCU.128, two coefficients per instruction, hoistable out of
 * loops) instead of materialising every 64-bit literal with two UMOVs per use — measured 10% of all
 * issued instructions in the CCpol kernel before this change (profiles/r1_ccpol_grad_v0.md). */
#define PIMDK_EXP_COEFS {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, \
  2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, \
  0.041666666666666664, 0.16666666666666666, 0.5, 1.0, 1.0}
#define PIMDK_LOG_COEFS {0.08695652173913043, 0.09523809523809523, 0.10526315789473684, 0.11764705882352941, \
  0.13333333333333333, 0.15384615384615385, 0.18181818181818182, 0.2222222222222222, 0.2857142857142857, 0.4, \
  0.6666666666666666}
#define PIMDK_SIN_COEFS {-8.22063524662433e-18, 2.8114572543455206e-15, -7.647163731819816e-13, 1.6059043836821613e-10, \
  -2.505210838544172e-08, 2.7557319223985893e-06, -0.0001984126984126984, 0.008333333333333333, -0.16666666666666666}
#define PIMDK_COS_COEFS {4.110317623312165e-19, -1.5619206968586225e-16, 4.779477332387385e-14, -1.1470745597729725e-11, \
  2.08767569878681e-09, -2.755731922398589e-07, 2.48015873015873e-05, -0.001388888888888889, 0.041666666666666664, -0.5}
#define PIMDK_ASIN_COEFS {0.0018622264064031275, 0.0019650336162772837, 0.0020776610325181676, 0.0022014739737101384, \
  0.002338091892111975, 0.0024894486782468836, 0.00265787063820729, 0.002846178401108942, 0.0030578216492580306, \
  0.003297059503473485, 0.0035692053938259347, 0.003880964558837669, 0.004240907093679363, 0.004660143486915096, \
  0.005153309682319905, 0.005740037670841924, 0.006447210311889649, 0.0073125258735988454, 0.008390335809616815, \
  0.009761609529194078, 0.011551800896139705, 0.01396484375, 0.017352764423076924, 0.022372159090909092, \
  0.030381944444444444, 0.044642857142857144, 0.075, 0.16666666666666666}
#if defined(__CUDACC__)
static __constant__ double pimdk_dc_exp[14] = PIMDK_EXP_COEFS;
static __constant__ double pimdk_dc_log[11] = PIMDK_LOG_COEFS;
static __constant__ double pimdk_dc_sin[9] = PIMDK_SIN_COEFS;
static __constant__ double pimdk_dc_cos[10] = PIMDK_COS_COEFS;
static __constant__ double pimdk_dc_asin[28] = PIMDK_ASIN_COEFS;
#endif
static const double pimdk_hc_exp[14] = PIMDK_EXP_COEFS;
static const double pimdk_hc_log[11] = PIMDK_LOG_COEFS;
static const double pimdk_hc_sin[9] = PIMDK_SIN_COEFS;
static const double pimdk_hc_cos[10] = PIMDK_COS_COEFS;
static const double pimdk_hc_asin[28] = PIMDK_ASIN_COEFS;
#if defined(__CUDA_ARCH__)
#define PIMDK_TAB(n) pimdk_dc_##n
#define PIMDK_UNROLL _Pragma("unroll")
#else
#define PIMDK_TAB(n) pimdk_hc_##n
#define PIMDK_UNROLL
#endif

PIMDK_HD uint64_t pimdk_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
PIMDK_HD double pimdk_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

/* exp(x): x = k ln2 + r, |r| <= ln2/2 (Cody-Waite, two-part ln2), degree-13 Taylor in Horner/fma
 * form, scaled by 2^k through the exponent field.  Flushes to 0 below -708, +inf above 709.
 * Device fast path (|x| < 700): one scaling by 2^k — the same bits as the general two-step scaling
 * whenever neither overflows nor underflows, which |x| < 700 guarantees. */
PIMDK_HD double pimdk_exp_poly(double r) {
  double p = PIMDK_TAB(exp)[0];
  PIMDK_UNROLL
  for (int i = 1; i < 14; ++i) p = PIMDK_FMA(p, r, PIMDK_TAB(exp)[i]);
  return p;
}
/* general path: any x (the device keeps it out of line so that each inlined copy of pimdk_exp is only
 * the ~22-instruction fast path; the hot kernels' code must stay inside the 32 KB instruction cache) */
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ double pimdk_exp_general(double x) {
#else
static inline double pimdk_exp_general(double x) {
#endif
  const double shifter = 6755399441055744.0; /* 1.5 * 2^52 */
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  if (x > 709.0) return pimdk_u2d(0x7ff0000000000000ull);
  double t = PIMDK_FMA(x, 1.4426950408889634, shifter);
  int64_t k = (int64_t)(int32_t)(pimdk_d2u(t) & 0xffffffffull);
  double kd = PIMDK_SUB(t, shifter);
  double r = PIMDK_FMA(kd, -6.93147180369123816490e-01, x); /* ln2 high part (fdlibm split) */
  r = PIMDK_FMA(kd, -1.90821492927058770002e-10, r);        /* ln2 low part */
  double p = pimdk_exp_poly(r);
  /* 2^k in two halves so that k in [-1022-52, 1023] never overflows the exponent field */
  int64_t k1 = k / 2, k2 = k - k1;
  double s1 = pimdk_u2d((uint64_t)(k1 + 1023) << 52);
  double s2 = pimdk_u2d((uint64_t)(k2 + 1023) << 52);
  return PIMDK_MUL(PIMDK_MUL(p, s1), s2);
}
/* Device: branch-free.  On (-708, 709] the general path's two-step scaling (p*2^k1)*2^k2 is exact in both
 * steps (the result is a normal number: k + 1023 lies in [2, 2046]), so it equals the single scaling p*2^k
 * bit for bit; outside that interval the general path returns 0, x (NaN) or +inf without arithmetic, which
 * two selects (and NaN propagation through the arithmetic) reproduce.  No branch means no convergence barrier around every exp: independent exps of
 * neighbouring site pairs interleave in the FP64 pipe instead of running one after the other. */
PIMDK_HD double pimdk_exp(double x) {
#if defined(__CUDA_ARCH__)
  const double shifter = 6755399441055744.0;
  double t = PIMDK_FMA(x, 1.4426950408889634, shifter);
  int k = __double2loint(t);
  double kd = PIMDK_SUB(t, shifter);
  double r = PIMDK_FMA(kd, -6.93147180369123816490e-01, x);
  r = PIMDK_FMA(kd, -1.90821492927058770002e-10, r);
  double res = PIMDK_MUL(pimdk_exp_poly(r), __hiloint2double((k + 1023) << 20, 0));
  /* +inf above 709, 0 at or below -708; a NaN argument fails both tests and leaves the NaN the arithmetic produced
   * (the host form returns x itself: the same value up to the NaN payload) */
  res = (x > 709.0) ? __longlong_as_double(0x7ff0000000000000ll) : res;
  res = (x <= -708.0) ? 0.0 : res;
  return res;
#else
  return pimdk_exp_general(x);
#endif
}

/* exp(x) for callers that guarantee x <= 0 (or NaN): e^{-beta R} with beta, R >= 0.  Same bits as pimdk_exp there;
 * the device form drops the overflow test (one FP64-pipe compare and a 64-bit select per call). */
PIMDK_HD double pimdk_exp_nonpos(double x) {
#if defined(__CUDA_ARCH__)
  const double shifter = 6755399441055744.0;
  double t = PIMDK_FMA(x, 1.4426950408889634, shifter);
  int k = __double2loint(t);
  double kd = PIMDK_SUB(t, shifter);
  double r = PIMDK_FMA(kd, -6.93147180369123816490e-01, x);
  r = PIMDK_FMA(kd, -1.90821492927058770002e-10, r);
  double res = PIMDK_MUL(pimdk_exp_poly(r), __hiloint2double((k + 1023) << 20, 0));
  return (x <= -708.0) ? 0.0 : res;
#else
  return pimdk_exp_general(x);
#endif
}

/* log(x), x > 0 finite normal: x = 2^e m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s),
 * s = (m-1)/(m+1), odd series to s^23; e ln2 added in two parts. */
PIMDK_HD double pimdk_log(double x) {
  uint64_t u = pimdk_d2u(x);
  int64_t e = (int64_t)((u >> 52) & 0x7ff) - 1023;
  uint64_t mant = u & 0x000fffffffffffffull;
  if (mant > 0x6a09e667f3bcdull) { /* m > sqrt(2): halve */
    e += 1;
    u = mant | ((uint64_t)1022 << 52);
  } else {
    u = mant | ((uint64_t)1023 << 52);
  }
  double m = pimdk_u2d(u);
  double f = PIMDK_SUB(m, 1.0);
  double s = PIMDK_DIV(f, PIMDK_ADD(m, 1.0));
  double z = PIMDK_MUL(s, s);
  double q = PIMDK_TAB(log)[0];
  PIMDK_UNROLL
  for (int i = 1; i < 11; ++i) q = PIMDK_FMA(q, z, PIMDK_TAB(log)[i]);
  /* log m = 2s + s*z*q ; 2s = f - s*f exactly-ish: use f - s*f to avoid the rounding of 2s */
  double sf = PIMDK_MUL(s, f);
  double lm = PIMDK_ADD(PIMDK_SUB(f, sf), PIMDK_MUL(PIMDK_MUL(s, z), q)); /* f - s f = 2s */
  double ed = (double)e;
  double hi = PIMDK_MUL(ed, 6.93147180369123816490e-01);
  double lo = PIMDK_FMA(ed, 1.90821492927058770002e-10, lm);
  return PIMDK_ADD(hi, lo);
}

/* cbrt(x), x > 0: exp(log(x)/3) refined by one Newton step with an fma-exact residual (<= 1 ulp) */
PIMDK_HD double pimdk_cbrt(double x) {
  double y = pimdk_exp(PIMDK_MUL(pimdk_log(x), 0.3333333333333333));
  double yy = PIMDK_MUL(y, y);
  double e1 = PIMDK_FMA(y, y, -yy);                       /* y*y = yy + e1 exactly */
  double r = PIMDK_FMA(e1, y, PIMDK_FMA(yy, y, -x));      /* y^3 - x */
  return PIMDK_SUB(y, PIMDK_DIV(r, PIMDK_MUL(3.0, yy)));
}

/* pow(x, y) for x > 0.  The exponents the reference actually uses are evaluated through
 * correctly rounded primitives (<= ~2.5 ulp): r**(-1.5d0) (proc_ccpol8s-dimer_xyz_ncd.f:288,404),
 * r**(-3.d0) (proc_sapt5sf_new_ncd.f:1511), r**(0.66666666666666666d0) (:1546).  Anything else
 * goes through exp(y log x) (a few ulp for |y log x| < 10). */
PIMDK_HD double pimdk_pow(double x, double y) {
  if (y == -1.5) return PIMDK_DIV(1.0, PIMDK_MUL(x, PIMDK_SQRT(x)));
  if (y == -3.0) return PIMDK_DIV(1.0, PIMDK_MUL(PIMDK_MUL(x, x), x));
  if (y == 0.66666666666666666) {
    double c = pimdk_cbrt(x);
    return PIMDK_MUL(c, c);
  }
  return pimdk_exp(PIMDK_MUL(y, pimdk_log(x)));
}

/* sin/cos kernels on |r| <= pi/4 */
PIMDK_HD double pimdk_sin_k(double r) {
  double z = PIMDK_MUL(r, r);
  double p = PIMDK_TAB(sin)[0];
  PIMDK_UNROLL
  for (int i = 1; i < 9; ++i) p = PIMDK_FMA(p, z, PIMDK_TAB(sin)[i]);
  return PIMDK_FMA(PIMDK_MUL(r, z), p, r);
}
PIMDK_HD double pimdk_cos_k(double r) {
  double z = PIMDK_MUL(r, r);
  double p = PIMDK_TAB(cos)[0];
  PIMDK_UNROLL
  for (int i = 1; i < 10; ++i) p = PIMDK_FMA(p, z, PIMDK_TAB(cos)[i]);
  return PIMDK_FMA(z, p, 1.0);
}
/* sin and cos together for |x| < ~1e5 (three-part pi/2 Cody-Waite reduction) */
PIMDK_HD void pimdk_sincos(double x, double* s, double* c) {
  const double shifter = 6755399441055744.0;
  double t = PIMDK_FMA(x, 0.6366197723675814, shifter); /* 2/pi */
  int32_t q = (int32_t)(pimdk_d2u(t) & 0xffffffffull);
  double kd = PIMDK_SUB(t, shifter);
  double r = PIMDK_FMA(kd, -1.57079632673412561417e+00, x); /* fdlibm pio2_1  */
  r = PIMDK_FMA(kd, -6.07710050630396597660e-11, r);        /* fdlibm pio2_2 */
  r = PIMDK_FMA(kd, -2.02226624879595063154e-21, r);        /* fdlibm pio2_2t */
  double sk = pimdk_sin_k(r), ck = pimdk_cos_k(r);
  switch (q & 3) {
    case 0: *s = sk; *c = ck; break;
    case 1: *s = ck; *c = -sk; break;
    case 2: *s = -sk; *c = -ck; break;
    default: *s = -ck; *c = sk; break;
  }
}
PIMDK_HD double pimdk_sin(double x) { double s, c; pimdk_sincos(x, &s, &c); return s; }
PIMDK_HD double pimdk_cos(double x) { double s, c; pimdk_sincos(x, &s, &c); return c; }

/* asin on |x| <= 0.5: odd Taylor series to x^57 */
PIMDK_HD double pimdk_asin_k(double x) {
  double z = PIMDK_MUL(x, x);
  /* coefficients c_k = (2k)! / (4^k (k!)^2 (2k+1)), k = 28 .. 1 */
  double p = PIMDK_TAB(asin)[0];
  PIMDK_UNROLL
  for (int i = 1; i < 28; ++i) p = PIMDK_FMA(p, z, PIMDK_TAB(asin)[i]);
  return PIMDK_FMA(PIMDK_MUL(x, z), p, x);
}
/* acos(x), |x| <= 1 */
PIMDK_HD double pimdk_acos(double x) {
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  if (x > 0.5) {
    double z = PIMDK_MUL(PIMDK_SUB(1.0, x), 0.5);
    double s = PIMDK_SQRT(z);
    return PIMDK_MUL(2.0, pimdk_asin_k(s));
  }
  if (x < -0.5) {
    double z = PIMDK_MUL(PIMDK_ADD(1.0, x), 0.5);
    double s = PIMDK_SQRT(z);
    double a = pimdk_asin_k(s);
    /* pi - 2a */
    return PIMDK_ADD(PIMDK_SUB(3.14159265358979311600e+00, PIMDK_MUL(2.0, a)), 1.22464679914735317720e-16);
  }
  return PIMDK_SUB(pio2_hi, PIMDK_SUB(pimdk_asin_k(x), pio2_lo));
}

/* expm1 on |r| <= 0.35: r (1 + r/2 + ... + r^13/14!) */
PIMDK_HD double pimdk_expm1_k(double r) {
  double p = 1.1470745597729725e-11;           /* 1/14! */
  p = PIMDK_FMA(p, r, 1.6059043836821613e-10);  /* 1/13! */
  p = PIMDK_FMA(p, r, 2.08767569878681e-09);
  p = PIMDK_FMA(p, r, 2.505210838544172e-08);
  p = PIMDK_FMA(p, r, 2.755731922398589e-07);
  p = PIMDK_FMA(p, r, 2.7557319223985893e-06);
  p = PIMDK_FMA(p, r, 2.48015873015873e-05);
  p = PIMDK_FMA(p, r, 0.0001984126984126984);
  p = PIMDK_FMA(p, r, 0.001388888888888889);
  p = PIMDK_FMA(p, r, 0.008333333333333333);
  p = PIMDK_FMA(p, r, 0.041666666666666664);
  p = PIMDK_FMA(p, r, 0.16666666666666666);
  p = PIMDK_FMA(p, r, 0.5);
  p = PIMDK_FMA(p, r, 1.0);
  return PIMDK_MUL(p, r);
}
/* tanh(x) = em/(em+2), em = e^{2x}-1 (Taylor expm1 for small |2x|, no cancellation) */
PIMDK_HD double pimdk_tanh(double x) {
  if (x > 20.0) return 1.0;
  if (x < -20.0) return -1.0;
  double t = PIMDK_MUL(2.0, x);
  double em = (t > -0.35 && t < 0.35) ? pimdk_expm1_k(t) : PIMDK_SUB(pimdk_exp(t), 1.0);
  return PIMDK_DIV(em, PIMDK_ADD(em, 2.0));
}

#endif /* PIMDK_DETMATH_H */
