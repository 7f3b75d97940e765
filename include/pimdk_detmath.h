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
/* exp: 2^(j/128), j = 0..127, correctly rounded (mpmath), and the Taylor coefficients 1/120, 1/24, 1/6, 1/2 */
#define PIMDK_EXP2_TAB { \
  1, 1.0054299011128027, 1.0108892860517005, 1.0163783149109531, \
  1.0218971486541166, 1.0274459491187637, 1.0330248790212284, 1.0386341019613787, \
  1.0442737824274138, 1.0499440858006872, 1.0556451783605572, 1.0613772272892621, \
  1.0671404006768237, 1.0729348675259756, 1.0787607977571199, 1.0846183622133092, \
  1.0905077326652577, 1.0964290818163769, 1.1023825833078409, 1.1083684117236787, \
  1.1143867425958924, 1.1204377524096067, 1.1265216186082418, 1.1326385195987192, \
  1.1387886347566916, 1.1449721444318042, 1.1511892299529827, 1.1574400736337511, \
  1.1637248587775775, 1.1700437696832502, 1.1763969916502812, 1.182784710984341, \
  1.189207115002721, 1.1956643920398273, 1.2021567314527031, 1.2086843236265816, \
  1.215247359980469, 1.2218460329727576, 1.22848053610687, 1.2351510639369334, \
  1.241857812073484, 1.2486009771892048, 1.2553807570246911, 1.2621973503942507, \
  1.2690509571917332, 1.275941778396392, 1.2828700160787783, 1.2898358734066657, \
  1.2968395546510096, 1.3038812651919358, 1.3109612115247644, 1.318079601266064, \
  1.3252366431597413, 1.3324325470831615, 1.3396675240533029, 1.3469417862329458, \
  1.3542555469368927, 1.3616090206382248, 1.3690024229745905, 1.3764359707545302, \
  1.383909881963832, 1.3914243757719262, 1.3989796725383112, 1.4065759938190154, \
  1.4142135623730951, 1.4218926021691656, 1.42961333839197, 1.4373759974489824, \
  1.4451808069770467, 1.4530279958490526, 1.460917794180647, 1.4688504333369818, \
  1.4768261459394993, 1.4848451658727524, 1.4929077282912648, 1.5010140696264256, \
  1.5091644275934228, 1.5173590411982147, 1.5255981507445384, 1.5338819978409559, \
  1.5422108254079407, 1.550584877685, 1.5590044002378369, 1.567469639965553, \
  1.5759808451078865, 1.5845382652524937, 1.593142151342267, 1.6017927556826934, \
  1.6104903319492543, 1.6192351351948637, 1.6280274218573478, 1.6368674497669644, \
  1.6457554781539649, 1.6546917676561943, 1.6636765803267364, 1.6727101796415966, \
  1.681792830507429, 1.6909247992693053, 1.7001063537185235, 1.7093377631004629, \
  1.7186192981224779, 1.7279512309618377, 1.7373338352737062, 1.746767386199169, \
  1.7562521603732995, 1.7657884359332727, 1.7753764925265212, 1.785016611318935, \
  1.7947090750031072, 1.8044541678066239, 1.8142521755003989, 1.8241033854070534, \
  1.8340080864093424, 1.843966568958626, 1.8539791250833855, 1.864046048397789, \
  1.8741676341103, 1.8843441790323345, 1.8945759815869656, 1.9048633418176741, \
  1.9152065613971474, 1.925605943636125, 1.9360617934922943, 1.9465744175792332, \
  1.9571441241754002, 1.9677712232331759, 1.9784560263879509, 1.9891988469672663}
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
static __device__ const double pimdk_dg_exp2[128] = PIMDK_EXP2_TAB;
static __constant__ double pimdk_dc_log[11] = PIMDK_LOG_COEFS;
static __constant__ double pimdk_dc_sin[9] = PIMDK_SIN_COEFS;
static __constant__ double pimdk_dc_cos[10] = PIMDK_COS_COEFS;
static __constant__ double pimdk_dc_asin[28] = PIMDK_ASIN_COEFS;
#endif
static const double pimdk_hc_exp[4] = PIMDK_EXP_COEFS;
static const double pimdk_hc_exp2[128] = PIMDK_EXP2_TAB;
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
#define PIMDK_EXP_CORE(x, res)                                                        \
  do {                                                                                \
    const double pimdk_t = PIMDK_FMA((x), 184.66496523378731, 6755399441055744.0);    \
    const int32_t pimdk_ki = (int32_t)(uint32_t)(pimdk_d2u(pimdk_t) & 0xffffffffull); \
    const double pimdk_kd = PIMDK_SUB(pimdk_t, 6755399441055744.0);                   \
    double pimdk_r = PIMDK_FMA(pimdk_kd, -5.41521234663377981633e-03, (x)); /* ln2 high part / 128 (fdlibm split) */ \
    pimdk_r = PIMDK_FMA(pimdk_kd, -1.49079291349264664064e-12, pimdk_r);    /* ln2 low part / 128 */ \
    const double pimdk_T = pimdk_exp2_scaled(pimdk_ki);                               \
    const double pimdk_r2 = PIMDK_MUL(pimdk_r, pimdk_r);                              \
    double pimdk_p = PIMDK_FMA(PIMDK_TAB(exp)[0], pimdk_r, PIMDK_TAB(exp)[1]);        \
    pimdk_p = PIMDK_FMA(pimdk_p, pimdk_r, PIMDK_TAB(exp)[2]);                         \
    pimdk_p = PIMDK_FMA(pimdk_p, pimdk_r, PIMDK_TAB(exp)[3]);                         \
    const double pimdk_q = PIMDK_FMA(pimdk_r2, pimdk_p, pimdk_r);                     \
    (res) = PIMDK_FMA(pimdk_T, pimdk_q, pimdk_T);                                     \
  } while (0)
/* 2^(ki >> 7) * 2^((ki & 127)/128): the table entry with ki >> 7 added to its exponent field (integer pipe).  For
 * arguments outside (-708, 709] the result is garbage that the callers' range tests discard. */
PIMDK_HD double pimdk_exp2_scaled(int32_t ki) {
#if defined(__CUDA_ARCH__)
  const double T = __ldg(&pimdk_dg_exp2[ki & 127]);
  return __hiloint2double(__double2hiint(T) + (int)((uint32_t)(ki >> 7) << 20), __double2loint(T));
#else
  return pimdk_u2d(pimdk_d2u(pimdk_hc_exp2[ki & 127]) + ((uint64_t)((uint32_t)(ki >> 7) << 20) << 32));
#endif
}
PIMDK_HD double pimdk_exp(double x) {
  double res;
#if defined(__CUDA_ARCH__)
  /* branch-free: no convergence barrier around every exp, so the independent exps of neighbouring site pairs
   * interleave in the FP64 pipe instead of running one after the other */
  PIMDK_EXP_CORE(x, res);
  res = (x > 709.0) ? __longlong_as_double(0x7ff0000000000000ll) : res;
  res = (x <= -708.0) ? 0.0 : res;
  return res;
#else
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  if (x > 709.0) return pimdk_u2d(0x7ff0000000000000ull);
  PIMDK_EXP_CORE(x, res);
  return res;
#endif
}

/* exp(x) for callers that guarantee x <= 0 (or NaN): e^{-beta R} with beta, R >= 0.  Same bits as pimdk_exp there;
 * the device form drops the overflow test (one FP64-pipe compare and a 64-bit select per call). */
PIMDK_HD double pimdk_exp_nonpos(double x) {
#if defined(__CUDA_ARCH__)
  double res;
  PIMDK_EXP_CORE(x, res);
  return (x <= -708.0) ? 0.0 : res;
#else
  return pimdk_exp(x);
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
