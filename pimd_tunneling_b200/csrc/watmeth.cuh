// Water-methane rigid-body surface of module watermethane_mod (watermethane.f90:1-520; plugin mcmod_watmeth.f90:1-65):
// 7 water sites (H H Q D D T T; the eighth site, O, carries no interaction) against 9 methane sites (H H H H C M M M M),
// each pair a Tang-Toennies term  e^{-beta r}(A0 + A1 r + AM/r) + q_a q_b/r + sum_n C_n/r^n P(n+1, delta_n r)  with the
// regularised incomplete gamma function P of Numerical Recipes (gammp / gser / gcf / gammln, :411-517, EPS = 3e-7).
// x is (3, 17): sites 1..8 water, 9..17 methane; energy in Hartree, distances in bohr.  V does not subtract V0
// (mcmod_watmeth.f90:15-27).  Arithmetic in the reference's operation order (built with -fmad=false); choices where Fortran
// leaves bits to the compiler: x**k and r**6.0d0 by binary powering, exp/log from the shared math policy.
#pragma once
#include "../../include/pimdk_detmath.h"

namespace pimdk {

constexpr int kWmWater = 7, kWmMethane = 9, kWmSites = 17;

struct WatMethTab {   // index [a*9 + b], a = water site, b = methane site
  double beta0[63], A0[63], A1[63], AM[63], C6[63], C8[63], C10[63], delta6[63], delta8[63], qq[63];
  double d6p7[63], d8p9[63], d8p11[63];   // delta6**7, delta8**9, delta8**11
  double g7, g9, g11;                     // exp(gammln(7)), exp(gammln(9)), exp(gammln(11))
};

PIMDK_HD double wm_ipow(double x, int n) {
  double result = 1.0;
  bool first = true;
  while (n) {
    if (n & 1) {
      if (first) { result = x; first = false; }
      else result = result * x;
    }
    n >>= 1;
    if (n) x = x * x;
  }
  return result;
}

// gammln (:496-517)
PIMDK_HD double wm_gammln(double xx) {
  const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, .1208650973866179e-2,
                         -.5395239384953e-5};
  const double stp = 2.5066282746310005;
  double x = xx, y = x;
  double tmp = x + 5.5;
  tmp = (x + 0.5) * pimdk_log(tmp) - tmp;
  double ser = 1.000000000190015;
  for (int j = 0; j < 6; ++j) {
    y = y + 1.0;
    ser = ser + cof[j] / y;
  }
  return tmp + pimdk_log(stp * ser / x);
}

// gammp (:411-430) with gser (:432-459) and gcf (:461-494); gln = gammln(a) is passed in (it depends on a alone)
PIMDK_HD double wm_gammp(double a, double x, double gln) {
  const double EPS = 3.e-7, fpmin = 1.e-30;
  if (x < a + 1.0) {            // series
    if (x <= 0.0) return 0.0;
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 1; n <= 100; ++n) {
      ap = ap + 1.0;
      del = del * x / ap;
      sum = sum + del;
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    return sum * pimdk_exp(-x + a * pimdk_log(x) - gln);
  }
  double b = x + 1.0 - a;       // continued fraction, modified Lentz
  double c = 1.0 / fpmin, d = 1.0 / b, h = d;
  for (int i = 1; i <= 100; ++i) {
    const double an = -(double)i * ((double)i - a);
    b = b + 2.0;
    d = an * d + b;
    if (fabs(d) < fpmin) d = fpmin;
    c = b + an / c;
    if (fabs(c) < fpmin) c = fpmin;
    d = 1.0 / d;
    const double del = d * c;
    h = h * del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return 1.0 - pimdk_exp(-x + a * pimdk_log(x) - gln) * h;
}

struct WmGln { double g7, g9, g11; };   // gammln(7), gammln(9), gammln(11)

// tangtoennies (:369-384)
PIMDK_HD double wm_pair_energy(const WatMethTab& T, const WmGln& G, int ab, double r) {
  double eint = pimdk_exp(-T.beta0[ab] * r) * (T.A0[ab] + T.A1[ab] * r + T.AM[ab] / r);
  eint = eint + T.qq[ab] / r;
  eint = eint + (T.C6[ab] / wm_ipow(r, 6)) * wm_gammp(7.0, T.delta6[ab] * r, G.g7);
  eint = eint + (T.C8[ab] / wm_ipow(r, 8)) * wm_gammp(9.0, T.delta8[ab] * r, G.g9);
  eint = eint + (T.C10[ab] / wm_ipow(r, 10)) * wm_gammp(11.0, T.delta8[ab] * r, G.g11);
  return eint;
}
// gradtangtoennies (:386-405)
PIMDK_HD double wm_pair_gradient(const WatMethTab& T, const WmGln& G, int ab, double r) {
  const double ex = pimdk_exp(-T.beta0[ab] * r);
  double grad = -T.beta0[ab] * ex * (T.A0[ab] + T.A1[ab] * r + T.AM[ab] / r);
  grad = grad + ex * (T.A1[ab] - T.AM[ab] / (r * r));
  grad = grad - T.qq[ab] / (r * r);
  grad = grad - 6.0 * (T.C6[ab] / wm_ipow(r, 7)) * wm_gammp(7.0, T.delta6[ab] * r, G.g7);
  grad = grad - 8.0 * (T.C8[ab] / wm_ipow(r, 9)) * wm_gammp(9.0, T.delta8[ab] * r, G.g9);
  grad = grad - 10.0 * (T.C10[ab] / wm_ipow(r, 11)) * wm_gammp(11.0, T.delta8[ab] * r, G.g11);
  grad = grad + T.C6[ab] * T.d6p7[ab] * pimdk_exp(-T.delta6[ab] * r) / T.g7;
  grad = grad + T.C8[ab] * T.d8p9[ab] * pimdk_exp(-T.delta8[ab] * r) / T.g9;
  grad = grad + T.C10[ab] * T.d8p11[ab] * pimdk_exp(-T.delta8[ab] * r) / T.g11;
  return grad;
}

// the unit conversions wmrb performs at every call (:279-287, 318-326), done once
inline void build_watmeth_tab(WatMethTab* T) {
  // columns: methane H (x4), C, M (x4); rows: water H (x2), Q, D (x2), T (x2)   (watermethane.f90:49-156)
  static const double betaang[4][3] = {{2.84808454, 2.7971225, 2.75581866}, {2.86928398, 2.3463075, 2.31474866},
                                       {5.71995231, 3.16999754, 2.35594058}, {6.24776382, 2.31915671, 2.28762859}};
  static const double Aang0[4][3] = {{-752.765963, -40504.8858, 5933.09667}, {4592.62807, 43408.9282, -5121.6292},
                                     {5367.76805, 55943.4633, -2584.27027}, {1258.12101, -19777.5292, 2979.79274}};
  static const double AangM[4][3] = {{908.685355, 26577.2947, -2622.41721}, {1252.30889, -23339.3574, 8389.34399},
                                     {-1430.20075, -84638.8183, 1242.92288}, {-162.796205, 8863.16664, -668.609725}};
  static const double Aang1[4][3] = {{417.797177, 19352.3942, -3719.99887}, {-1789.69987, -29232.8793, 5058.10255},
                                     {-14542.3959, -5391.49926, 651.930524}, {-9354.24387, 10463.6792, -2062.86173}};
  static const double Cang6[4][3] = {{-31.1396325, -176.385261, 26.1819133}, {-1291.50705, -23944.8325, 7240.4699},
                                     {172.100547, 4688.06162, -1425.39796}, {0.0, 0.0, 0.0}};
  static const double Cang8[4][3] = {{40.6973228, -470.183908, 212.925753}, {7345.62345, 132928.009, -46577.4854},
                                     {-1195.7863, -32880.3016, 11355.4445}, {0.0, 0.0, 0.0}};
  static const double Cang10[4][3] = {{-13.9555905, 334.00843, -620.561765}, {-12518.903, -119240.978, 62124.298},
                                      {1655.75062, -8388.62553, -13763.1195}, {0.0, 0.0, 0.0}};
  static const double deltaang6[4][3] = {{7.335799, 2.825277, 1.943410}, {4.341591, 4.288189, 4.259787},
                                         {5.759895, 6.129260, 3.737571}, {0.0, 0.0, 0.0}};
  static const double deltaang8[4][3] = {{1.2192e-2, 31.106042, 5.747e-3}, {3.643903, 4.138380, 4.368576},
                                         {4.415080, 3.962102, 3.741763}, {0.0, 0.0, 0.0}};
  static const double watercharge[7] = {0.494714, 0.494714, -1.830627, 0.420599, 0.420599, 0.0, 0.0};
  static const double methanecharge[9] = {0.279901, 0.279901, 0.279901, 0.279901, 3.590472, -1.177519, -1.177519, -1.177519, -1.177519};
  static const int wclass[7] = {0, 0, 1, 2, 2, 3, 3};
  static const int mclass[9] = {0, 0, 0, 0, 1, 2, 2, 2, 2};
  const double ang = 0.529177, kc = 1.59362e-3;
  for (int a = 0; a < 7; ++a)
    for (int b = 0; b < 9; ++b) {
      const int ab = a * 9 + b, r = wclass[a], c = mclass[b];
      T->beta0[ab] = betaang[r][c] * ang;
      T->delta6[ab] = deltaang6[r][c] * ang;
      T->delta8[ab] = deltaang8[r][c] * ang;
      T->A0[ab] = Aang0[r][c] * kc;
      T->AM[ab] = AangM[r][c] * kc / ang;
      T->A1[ab] = Aang1[r][c] * kc * ang;
      T->C6[ab] = Cang6[r][c] * kc / wm_ipow(ang, 6);
      T->C8[ab] = Cang8[r][c] * kc / wm_ipow(ang, 8);
      T->C10[ab] = Cang10[r][c] * kc / wm_ipow(ang, 10);
      T->qq[ab] = watercharge[a] * methanecharge[b];
      T->d6p7[ab] = wm_ipow(T->delta6[ab], 7);
      T->d8p9[ab] = wm_ipow(T->delta8[ab], 9);
      T->d8p11[ab] = wm_ipow(T->delta8[ab], 11);
    }
  T->g7 = pimdk_exp(wm_gammln(7.0));
  T->g9 = pimdk_exp(wm_gammln(9.0));
  T->g11 = pimdk_exp(wm_gammln(11.0));
}

}  // namespace pimdk
