// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// CPU restatement of module watermethane_mod's rigid-body water-methane surface (watermethane.f90): wmrb (:307-354),
// wmrb_grad (:267-305), calcr (:356-367), tangtoennies (:369-384), gradtangtoennies (:386-405) and the Numerical Recipes
// incomplete gamma function gammp / gser / gcf / gammln (:411-517), as driven by mcmod_watmeth.f90:15-62 (V, Vprime,
// Vdoubleprime).  x(51) = 17 sites x 3 in bohr: H H Q D D T T O | H H H H C M M M M; energies in Hartree.
// Compiled as if -r8 (3.e-7, 1.e-30, 1. are double).  x**k and r**6.0d0: binary powering (the repository's stated choice
// where Fortran leaves the bits to the compiler).
#pragma once
#include <cmath>
#include <stdexcept>

#include "../include/pimdk_detmath.h"

namespace oracle {

struct WaterMethane {
  static constexpr int waterdof = 7, methanedof = 9;
  double beta0[7][9], A0[7][9], A1[7][9], AM[7][9], C6[7][9], C8[7][9], C10[7][9], delta6[7][9], delta8[7][9];
  double watercharge[7], methanecharge[9];

  static double ipow(double x, int n) {
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

  WaterMethane() {
    // data statements :49-156: each row is (methane H x4, C, M x4); water rows H H Q D D T T
    const double betaang[4][3] = {{2.84808454, 2.7971225, 2.75581866}, {2.86928398, 2.3463075, 2.31474866},
                                  {5.71995231, 3.16999754, 2.35594058}, {6.24776382, 2.31915671, 2.28762859}};
    const double Aang0[4][3] = {{-752.765963, -40504.8858, 5933.09667}, {4592.62807, 43408.9282, -5121.6292},
                                {5367.76805, 55943.4633, -2584.27027}, {1258.12101, -19777.5292, 2979.79274}};
    const double AangM[4][3] = {{908.685355, 26577.2947, -2622.41721}, {1252.30889, -23339.3574, 8389.34399},
                                {-1430.20075, -84638.8183, 1242.92288}, {-162.796205, 8863.16664, -668.609725}};
    const double Aang1[4][3] = {{417.797177, 19352.3942, -3719.99887}, {-1789.69987, -29232.8793, 5058.10255},
                                {-14542.3959, -5391.49926, 651.930524}, {-9354.24387, 10463.6792, -2062.86173}};
    const double Cang6[4][3] = {{-31.1396325, -176.385261, 26.1819133}, {-1291.50705, -23944.8325, 7240.4699},
                                {172.100547, 4688.06162, -1425.39796}, {0.0, 0.0, 0.0}};
    const double Cang8[4][3] = {{40.6973228, -470.183908, 212.925753}, {7345.62345, 132928.009, -46577.4854},
                                {-1195.7863, -32880.3016, 11355.4445}, {0.0, 0.0, 0.0}};
    const double Cang10[4][3] = {{-13.9555905, 334.00843, -620.561765}, {-12518.903, -119240.978, 62124.298},
                                 {1655.75062, -8388.62553, -13763.1195}, {0.0, 0.0, 0.0}};
    const double deltaang6[4][3] = {{7.335799, 2.825277, 1.943410}, {4.341591, 4.288189, 4.259787},
                                    {5.759895, 6.129260, 3.737571}, {0.0, 0.0, 0.0}};
    const double deltaang8[4][3] = {{1.2192e-2, 31.106042, 5.747e-3}, {3.643903, 4.138380, 4.368576},
                                    {4.415080, 3.962102, 3.741763}, {0.0, 0.0, 0.0}};
    const double wq[7] = {0.494714, 0.494714, -1.830627, 0.420599, 0.420599, 0.0, 0.0};
    const double mq[9] = {0.279901, 0.279901, 0.279901, 0.279901, 3.590472, -1.177519, -1.177519, -1.177519, -1.177519};
    const int wc[7] = {0, 0, 1, 2, 2, 3, 3}, mc[9] = {0, 0, 0, 0, 1, 2, 2, 2, 2};
    for (int i = 0; i < 7; ++i) watercharge[i] = wq[i];
    for (int j = 0; j < 9; ++j) methanecharge[j] = mq[j];
    // "Convert these from crappy units to atomic units" (:279-287)
    for (int i = 0; i < 7; ++i)
      for (int j = 0; j < 9; ++j) {
        const int r = wc[i], c = mc[j];
        beta0[i][j] = betaang[r][c] * 0.529177;
        delta6[i][j] = deltaang6[r][c] * 0.529177;
        delta8[i][j] = deltaang8[r][c] * 0.529177;
        A0[i][j] = Aang0[r][c] * 1.59362e-3;
        AM[i][j] = AangM[r][c] * 1.59362e-3 / 0.529177;
        A1[i][j] = Aang1[r][c] * 1.59362e-3 * 0.529177;
        C6[i][j] = Cang6[r][c] * 1.59362e-3 / ipow(0.529177, 6);
        C8[i][j] = Cang8[r][c] * 1.59362e-3 / ipow(0.529177, 8);
        C10[i][j] = Cang10[r][c] * 1.59362e-3 / ipow(0.529177, 10);
      }
  }

  static double gammln(double xx) {   // :496-517
    const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155,
                           .1208650973866179e-2, -.5395239384953e-5};
    const double stp = 2.5066282746310005;
    double x = xx, y = x;
    double tmp = x + 5.5;
    tmp = (x + 0.5) * pimdk_log(tmp) - tmp;
    double ser = 1.000000000190015;
    for (int j = 1; j <= 6; ++j) {
      y = y + 1.0;
      ser = ser + cof[j - 1] / y;
    }
    return tmp + pimdk_log(stp * ser / x);
  }
  static void gser(double& gamser, double a, double x, double& gln) {   // :432-459
    const int ITMAX = 100;
    const double EPS = 3.e-7;
    gln = gammln(a);
    if (x <= 0.0) {
      if (x < 0.0) throw std::runtime_error("gser");
      gamser = 0.0;
      return;
    }
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 1; n <= ITMAX; ++n) {
      ap = ap + 1.;
      del = del * x / ap;
      sum = sum + del;
      if (std::fabs(del) < std::fabs(sum) * EPS) {
        gamser = sum * pimdk_exp(-x + a * pimdk_log(x) - gln);
        return;
      }
    }
    throw std::runtime_error("gser");
  }
  static void gcf(double& gammcf, double a, double x, double& gln) {   // :461-494
    const int itmax = 100;
    const double eps = 3.e-7, fpmin = 1.e-30;
    gln = gammln(a);
    double b = x + 1. - a;
    double c = 1.0 / fpmin;
    double d = 1.0 / b;
    double h = d;
    for (int i = 1; i <= itmax; ++i) {
      double an = -i * (i - a);
      b = b + 2.;
      d = an * d + b;
      if (std::fabs(d) < fpmin) d = fpmin;
      c = b + an / c;
      if (std::fabs(c) < fpmin) c = fpmin;
      d = 1.0 / d;
      double del = d * c;
      h = h * del;
      if (std::fabs(del - 1.0) < eps) {
        gammcf = pimdk_exp(-x + a * pimdk_log(x) - gln) * h;
        return;
      }
    }
    throw std::runtime_error("gcf");
  }
  static double gammp(double a, double x) {   // :411-430
    if (x < 0.0 || a <= 0.0) throw std::runtime_error("gammp");
    double gammcf, gamser, gln;
    if (x < a + 1.) {
      gser(gamser, a, x, gln);
      return gamser;
    }
    gcf(gammcf, a, x, gln);
    return 1.0 - gammcf;
  }
  static double calcr(const double* x1, const double* x2) {   // :356-367
    double r = 0.0;
    for (int i = 0; i < 3; ++i) r = r + (x1[i] - x2[i]) * (x1[i] - x2[i]);
    return std::sqrt(r);
  }
  double tangtoennies(double r, int a, int b) const {   // :369-384 (a, b 0-based here)
    double eint = pimdk_exp(-beta0[a][b] * r) * (A0[a][b] + A1[a][b] * r + AM[a][b] / r);
    eint = eint + watercharge[a] * methanecharge[b] / (r);
    eint = eint + (C6[a][b] / ipow(r, 6)) * gammp(7.0, delta6[a][b] * r);
    eint = eint + (C8[a][b] / ipow(r, 8)) * gammp(9.0, delta8[a][b] * r);
    eint = eint + (C10[a][b] / ipow(r, 10)) * gammp(11.0, delta8[a][b] * r);
    return eint;
  }
  double gradtangtoennies(double r, int a, int b) const {   // :386-405
    double grad = -beta0[a][b] * pimdk_exp(-beta0[a][b] * r) * (A0[a][b] + A1[a][b] * r + AM[a][b] / r);
    grad = grad + pimdk_exp(-beta0[a][b] * r) * (A1[a][b] - AM[a][b] / ipow(r, 2));
    grad = grad - watercharge[a] * methanecharge[b] / (ipow(r, 2));
    grad = grad - 6.0 * (C6[a][b] / ipow(r, 7)) * gammp(7.0, delta6[a][b] * r);
    grad = grad - 8.0 * (C8[a][b] / ipow(r, 9)) * gammp(9.0, delta8[a][b] * r);
    grad = grad - 10.0 * (C10[a][b] / ipow(r, 11)) * gammp(11.0, delta8[a][b] * r);
    grad = grad + C6[a][b] * (ipow(delta6[a][b], 7)) * pimdk_exp(-delta6[a][b] * r) / pimdk_exp(gammln(7.0));
    grad = grad + C8[a][b] * (ipow(delta8[a][b], 9)) * pimdk_exp(-delta8[a][b] * r) / pimdk_exp(gammln(9.0));
    grad = grad + C10[a][b] * (ipow(delta8[a][b], 11)) * pimdk_exp(-delta8[a][b] * r) / pimdk_exp(gammln(11.0));
    return grad;
  }
  double wmrb(const double* x) const {   // :307-335, gradt = .false.
    double ereal = 0.0;
    for (int i = 1; i <= waterdof; ++i)
      for (int j = 1; j <= methanedof; ++j) {
        double r12 = calcr(x + 3 * (i - 1), x + 3 * (j - 1) + 24);
        ereal = ereal + tangtoennies(r12, i - 1, j - 1);
      }
    return ereal;
  }
  void wmrb_grad(const double* x, double* grad) const {   // :267-305
    for (int d = 0; d < 51; ++d) grad[d] = 0.0;
    for (int i = 1; i <= waterdof; ++i)
      for (int j = 1; j <= methanedof; ++j) {
        double r12 = calcr(x + 3 * (i - 1), x + 3 * (j - 1) + 24);
        for (int k = 1; k <= 3; ++k) {
          double rk = x[3 * (i - 1) + k - 1] - x[3 * (j - 1) + k + 24 - 1];
          grad[3 * (i - 1) + k - 1] = grad[3 * (i - 1) + k - 1] + (rk * gradtangtoennies(r12, i - 1, j - 1) / r12);
          grad[3 * (j - 1) + k + 24 - 1] = grad[3 * (j - 1) + k + 24 - 1] - (rk * gradtangtoennies(r12, i - 1, j - 1) / r12);
        }
      }
  }
};

}  // namespace oracle
