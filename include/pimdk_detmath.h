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
/* exp: bit patterns of 2^(j/128), j = 0..127, correctly rounded (mpmath), each with j << 45 subtracted so that adding
 * ki << 45 (ki = 128 k + j) lands k in the exponent field; and the Taylor coefficients 1/120, 1/24, 1/6, 1/2 */
#define PIMDK_EXP2_TAB { \
  0x3ff0000000000000ull, 0x3feff63da9fb3335ull, 0x3fefec9a3e778061ull, 0x3fefe315e86e7f85ull, \
  0x3fefd9b0d3158574ull, 0x3fefd06b29ddf6deull, 0x3fefc74518759bc8ull, 0x3fefbe3ecac6f383ull, \
  0x3fefb5586cf9890full, 0x3fefac922b7247f7ull, 0x3fefa3ec32d3d1a2ull, 0x3fef9b66affed31bull, \
  0x3fef9301d0125b51ull, 0x3fef8abdc06c31ccull, 0x3fef829aaea92de0ull, 0x3fef7a98c8a58e51ull, \
  0x3fef72b83c7d517bull, 0x3fef6af9388c8deaull, 0x3fef635beb6fcb75ull, 0x3fef5be084045cd4ull, \
  0x3fef54873168b9aaull, 0x3fef4d5022fcd91dull, 0x3fef463b88628cd6ull, 0x3fef3f49917ddc96ull, \
  0x3fef387a6e756238ull, 0x3fef31ce4fb2a63full, 0x3fef2b4565e27cddull, 0x3fef24dfe1f56381ull, \
  0x3fef1e9df51fdee1ull, 0x3fef187fd0dad990ull, 0x3fef1285a6e4030bull, 0x3fef0cafa93e2f56ull, \
  0x3fef06fe0a31b715ull, 0x3fef0170fc4cd831ull, 0x3feefc08b26416ffull, 0x3feef6c55f929ff1ull, \
  0x3feef1a7373aa9cbull, 0x3feeecae6d05d866ull, 0x3feee7db34e59ff7ull, 0x3feee32dc313a8e5ull, \
  0x3feedea64c123422ull, 0x3feeda4504ac801cull, 0x3feed60a21f72e2aull, 0x3feed1f5d950a897ull, \
  0x3feece086061892dull, 0x3feeca41ed1d0057ull, 0x3feec6a2b5c13cd0ull, 0x3feec32af0d7d3deull, \
  0x3feebfdad5362a27ull, 0x3feebcb299fddd0dull, 0x3feeb9b2769d2ca7ull, 0x3feeb6daa2cf6642ull, \
  0x3feeb42b569d4f82ull, 0x3feeb1a4ca5d920full, 0x3feeaf4736b527daull, 0x3feead12d497c7fdull, \
  0x3feeab07dd485429ull, 0x3feea9268a5946b7ull, 0x3feea76f15ad2148ull, 0x3feea5e1b976dc09ull, \
  0x3feea47eb03a5585ull, 0x3feea34634ccc320ull, 0x3feea23882552225ull, 0x3feea155d44ca973ull, \
  0x3feea09e667f3bcdull, 0x3feea012750bdabfull, 0x3fee9fb23c651a2full, 0x3fee9f7df9519484ull, \
  0x3fee9f75e8ec5f74ull, 0x3fee9f9a48a58174ull, 0x3fee9feb564267c9ull, 0x3feea0694fde5d3full, \
  0x3feea11473eb0187ull, 0x3feea1ed0130c132ull, 0x3feea2f336cf4e62ull, 0x3feea427543e1a12ull, \
  0x3feea589994cce13ull, 0x3feea71a4623c7adull, 0x3feea8d99b4492edull, 0x3feeaac7d98a6699ull, \
  0x3feeace5422aa0dbull, 0x3feeaf3216b5448cull, 0x3feeb1ae99157736ull, 0x3feeb45b0b91ffc6ull, \
  0x3feeb737b0cdc5e5ull, 0x3feeba44cbc8520full, 0x3feebd829fde4e50ull, 0x3feec0f170ca07baull, \
  0x3feec49182a3f090ull, 0x3feec86319e32323ull, 0x3feecc667b5de565ull, 0x3feed09bec4a2d33ull, \
  0x3feed503b23e255dull, 0x3feed99e1330b358ull, 0x3feede6b5579fdbfull, 0x3feee36bbfd3f37aull, \
  0x3feee89f995ad3adull, 0x3feeee07298db666ull, 0x3feef3a2b84f15fbull, 0x3feef9728de5593aull, \
  0x3feeff76f2fb5e47ull, 0x3fef05b030a1064aull, 0x3fef0c1e904bc1d2ull, 0x3fef12c25bd71e09ull, \
  0x3fef199bdd85529cull, 0x3fef20ab5fffd07aull, 0x3fef27f12e57d14bull, 0x3fef2f6d9406e7b5ull, \
  0x3fef3720dcef9069ull, 0x3fef3f0b555dc3faull, 0x3fef472d4a07897cull, 0x3fef4f87080d89f2ull, \
  0x3fef5818dcfba487ull, 0x3fef60e316c98398ull, 0x3fef69e603db3285ull, 0x3fef7321f301b460ull, \
  0x3fef7c97337b9b5full, 0x3fef864614f5a129ull, 0x3fef902ee78b3ff6ull, 0x3fef9a51fbc74c83ull, \
  0x3fefa4afa2a490daull, 0x3fefaf482d8e67f1ull, 0x3fefba1bee615a27ull, 0x3fefc52b376bba97ull, \
  0x3fefd0765b6e4540ull, 0x3fefdbfdad9cbe14ull, 0x3fefe7c1819e90d8ull, 0x3feff3c22b8f71f1ull}
#define PIMDK_EXP_COEFS {0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5}
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
static __constant__ double pimdk_dc_exp[4] = PIMDK_EXP_COEFS;
/* the 2^(j/128) table is indexed per lane: global memory through the read-only path (L1-resident, 1 KB) */
static __device__ const unsigned long long pimdk_dg_exp2[128] = PIMDK_EXP2_TAB;
#if defined(PIMDK_EXP2_SHARED)
/* a translation unit that defines PIMDK_EXP2_SHARED also gets pimdk_exp_nonpos_sh(), which reads the table from shared
 * memory (a 32-bit address, no 64-bit address arithmetic per call); a kernel that uses it calls pimdk_exp2_stage() first */
__shared__ unsigned long long pimdk_sh_exp2[128];
static __device__ __forceinline__ void pimdk_exp2_stage() {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) pimdk_sh_exp2[i] = pimdk_dg_exp2[i];
  __syncthreads();
}
#endif
static __constant__ double pimdk_dc_log[11] = PIMDK_LOG_COEFS;
static __constant__ double pimdk_dc_sin[9] = PIMDK_SIN_COEFS;
static __constant__ double pimdk_dc_cos[10] = PIMDK_COS_COEFS;
static __constant__ double pimdk_dc_asin[28] = PIMDK_ASIN_COEFS;
#endif
static const double pimdk_hc_exp[4] = PIMDK_EXP_COEFS;
static const uint64_t pimdk_hc_exp2[128] = PIMDK_EXP2_TAB;
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

/* exp(x) = 2^k 2^(j/128) e^r:  ki = rint(x 128/ln2) = 128 k + j,  r = x - ki ln2/128 (Cody-Waite, two-part ln2;
 * the first step is exact), |r| <= ln2/256, so e^r - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120) to 6e-19.
 * The table value is scaled by 2^k through its exponent field BEFORE the last fma, T' (1 + q) = fma(T', q, T'):
 * ten FP64 operations in all (the degree-13 Taylor form this replaces took nineteen), < 1 ulp (table 0.5 + final
 * rounding 0.5).  Returns 0 at or below -708 and +inf above 709; on (-708, 709] every intermediate is a normal
 * number, so the scaling is exact.  NaN propagates through the arithmetic (fma(T', NaN, T')). */
#define PIMDK_EXP_CORE(x, res, T2K)                                                       \
  do {                                                                                \
    const double pimdk_t = PIMDK_FMA((x), 184.66496523378731, 6755399441055744.0);    \
    const int32_t pimdk_ki = (int32_t)(uint32_t)(pimdk_d2u(pimdk_t) & 0xffffffffull); \
    const double pimdk_kd = PIMDK_SUB(pimdk_t, 6755399441055744.0);                   \
    double pimdk_r = PIMDK_FMA(pimdk_kd, -5.41521234663377981633e-03, (x)); /* ln2 high part / 128 (fdlibm split) */ \
    pimdk_r = PIMDK_FMA(pimdk_kd, -1.49079291349264664064e-12, pimdk_r);    /* ln2 low part / 128 */ \
    const double pimdk_T = T2K(pimdk_ki);                                             \
    const double pimdk_r2 = PIMDK_MUL(pimdk_r, pimdk_r);                              \
    double pimdk_p = PIMDK_FMA(PIMDK_TAB(exp)[0], pimdk_r, PIMDK_TAB(exp)[1]);        \
    pimdk_p = PIMDK_FMA(pimdk_p, pimdk_r, PIMDK_TAB(exp)[2]);                         \
    pimdk_p = PIMDK_FMA(pimdk_p, pimdk_r, PIMDK_TAB(exp)[3]);                         \
    const double pimdk_q = PIMDK_FMA(pimdk_r2, pimdk_p, pimdk_r);                     \
    (res) = PIMDK_FMA(pimdk_T, pimdk_q, pimdk_T);                                     \
  } while (0)
/* 2^(ki >> 7) * 2^((ki & 127)/128): the (pre-biased) table entry with ki << 13 added to its high word — one integer
 * multiply-add.  For arguments outside (-708, 709] the result is garbage that the callers' range tests discard. */
PIMDK_HD double pimdk_exp2_scaled(int32_t ki) {
#if defined(__CUDA_ARCH__)
  const unsigned long long b = __ldg(&pimdk_dg_exp2[ki & 127]);
  return __hiloint2double((int)((unsigned)(b >> 32) + ((unsigned)ki << 13)), (int)(unsigned)b);
#else
  return pimdk_u2d(pimdk_hc_exp2[ki & 127] + ((uint64_t)(uint32_t)ki << 45));
#endif
}
PIMDK_HD double pimdk_exp(double x) {
  double res;
#if defined(__CUDA_ARCH__)
  /* branch-free: no convergence barrier around every exp, so the independent exps of neighbouring site pairs
   * interleave in the FP64 pipe instead of running one after the other */
  PIMDK_EXP_CORE(x, res, pimdk_exp2_scaled);
  res = (x > 709.0) ? __longlong_as_double(0x7ff0000000000000ll) : res;
  res = (x <= -708.0) ? 0.0 : res;
  return res;
#else
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  if (x > 709.0) return pimdk_u2d(0x7ff0000000000000ull);
  PIMDK_EXP_CORE(x, res, pimdk_exp2_scaled);
  return res;
#endif
}

/* exp(x) for callers that guarantee x <= 0 (or NaN): e^{-beta R} with beta, R >= 0.  Same bits as pimdk_exp there;
 * the device form drops the overflow test (one FP64-pipe compare and a 64-bit select per call). */
PIMDK_HD double pimdk_exp_nonpos(double x) {
#if defined(__CUDA_ARCH__)
  double res;
  PIMDK_EXP_CORE(x, res, pimdk_exp2_scaled);
  return (x <= -708.0) ? 0.0 : res;
#else
  return pimdk_exp(x);
#endif
}

#if defined(__CUDACC__) && defined(PIMDK_EXP2_SHARED)
/* pimdk_exp_nonpos with the table read from shared memory (kernels that called pimdk_exp2_stage()): same bits */
static __device__ __forceinline__ double pimdk_exp2_scaled_sh(int32_t ki) {
  const unsigned long long b = pimdk_sh_exp2[ki & 127];
  return __hiloint2double((int)((unsigned)(b >> 32) + ((unsigned)ki << 13)), (int)(unsigned)b);
}
static __device__ __forceinline__ double pimdk_exp_nonpos_sh(double x) {
  double res;
  PIMDK_EXP_CORE(x, res, pimdk_exp2_scaled_sh);
  return (x <= -708.0) ? 0.0 : res;
}
#endif

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

/* atan(x): t = |x| or 1/|x| in [0, 1]; atan t = asin(y), y = t / sqrt(1 + t^2) <= 0.7072, with
 * asin(y) = pi/2 - 2 asin(sqrt((1 - y)/2)) above 0.5; atan|x| = pi/2 - atan(1/|x|) for |x| > 1 (<= 3 ulp).
 * Used by the Eckart embedding (eck_rad_tst, main_CCpol-8sf.f:597-716) only: one call per monomer and energy. */
PIMDK_HD double pimdk_atan(double x) {
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  const double ax = x < 0.0 ? -x : x;
  const int inv = ax > 1.0;
  const double t = inv ? PIMDK_DIV(1.0, ax) : ax;
  const double y = PIMDK_DIV(t, PIMDK_SQRT(PIMDK_FMA(t, t, 1.0)));
  double a;
  if (y > 0.5) {
    const double s = PIMDK_SQRT(PIMDK_MUL(PIMDK_SUB(1.0, y), 0.5));
    a = PIMDK_ADD(PIMDK_SUB(pio2_hi, PIMDK_MUL(2.0, pimdk_asin_k(s))), pio2_lo);
  } else {
    a = pimdk_asin_k(y);
  }
  if (inv) a = PIMDK_ADD(PIMDK_SUB(pio2_hi, a), pio2_lo);
  return x < 0.0 ? -a : a;
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
