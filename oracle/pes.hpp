// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// CPU restatement of the three in-scope `module mcmod_mass` plugins:
//   mcmod_1d.f90:8-58, mcmod_2dtest.f90:11-61, mcmod_waterdimer_ccpol.f90:9-77, mcmod_so2.f90:10-84,
//   mcmod_watmeth.f90:10-62 (watmeth.hpp), mcmod_malon.f90:10-72 (malon.hpp).
// x and grad are Fortran (ndim,natom) column-major: index = (atom)*ndim + dim.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "ccpol_impl.hpp"
#include "tables.hpp"
#include "malon.hpp"
#include "watmeth.hpp"

namespace oracle {

enum PesKind { PES_1D = 1, PES_2DTEST = 2, PES_CCPOL = 3, PES_SO2 = 4, PES_WATMETH = 5, PES_MALON = 6 };

struct Pes {
  PesKind kind = PES_1D;
  int ndim = 1, natom = 1;
  double V0 = 0.0;
  // 1D (mcmod_1d.f90:9-12)
  double Vheight = 1.0, x0_1d = 1.0;
  // 2D (mcmod_2dtest.f90:16-24)
  int m = 6;
  double a0 = 2.0, b0 = 0.2, rho0 = 3.0;
  double wx[6], wy[6];
  // so2 (mcmod_so2.f90:10-14)
  double omegaforce = 10000.0, r0 = 20.0;
  // CCpol
  const CcpolTables* tab = nullptr;
  long potcount = 0;

  void init_1d() {
    kind = PES_1D;
    Vheight = 1.0;
    x0_1d = 1.0;
  }
  void init_2d() {
    kind = PES_2DTEST;
    ndim = 2;
    natom = 1;
    const double pi = 3.14159265358979;  // mcmod_2dtest.f90:14 (truncated literal)
    m = 6;
    a0 = 2.0;
    b0 = 0.2;
    rho0 = 3.0;
    for (int k = 1; k <= m; ++k) {
      wx[k - 1] = rho0 * std::cos((double)k * 2.0 * pi / (double)m);
      wy[k - 1] = rho0 * std::sin((double)k * 2.0 * pi / (double)m);
    }
    V0 = 0.0;  // never initialised by the reference: static storage -> 0
  }
  void init_so2() {
    kind = PES_SO2;
    ndim = 2;
    natom = 1;
    omegaforce = 10000.0;
    r0 = 20.0;
    V0 = 0.0;
  }
  // water-methane (mcmod_watmeth.f90:10-13)
  WaterMethane wm;
  void init_watmeth() {
    kind = PES_WATMETH;
    ndim = 3;
    natom = 17;
    V0 = 0.0;
  }
  // malonaldehyde (mcmod_malon.f90:10-13: V_init does nothing; the fit lives in pes' DATA statements)
  const Malonaldehyde* mal = nullptr;
  void init_malon(const Malonaldehyde* t) {
    kind = PES_MALON;
    ndim = 3;
    natom = 9;
    mal = t;
    V0 = 0.0;
  }
  void init_ccpol(const CcpolTables* t) {
    kind = PES_CCPOL;
    ndim = 3;
    natom = 6;
    tab = t;
    V0 = 0.0;  // mcmod_waterdimer_ccpol.f90:14
  }

  // function V(x)
  double V(const double* x) const {
    switch (kind) {
      case PES_1D: {  // mcmod_1d.f90:20   (ignores V0)
        double s = 0.0;
        for (int i = 0; i < ndim * natom; ++i) {
          double u = (x[i] / x0_1d) * (x[i] / x0_1d) - 1.0;
          s += Vheight * (u * u);
        }
        return s;
      }
      case PES_2DTEST: {  // mcmod_2dtest.f90:33-39
        double answer = 0.0;
        for (int k = 0; k < m; ++k) {
          double dx = x[0] - wx[k], dy = x[1] - wy[k];
          answer = answer - 0.5 * pimdk_exp(-a0 * (dx * dx + dy * dy));
          answer = answer - 0.5 * pimdk_exp(-b0 * (dx * dx + dy * dy));
        }
        return answer - V0;
      }
      case PES_SO2: {  // mcmod_so2.f90:22-31
        double r = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        double answer = 0.5 * (omegaforce * omegaforce) * ((r - r0) * (r - r0));
        return answer - V0;
      }
      case PES_WATMETH:  // mcmod_watmeth.f90:15-27 (V0 is not subtracted)
        return wm.wmrb(x);
      case PES_MALON:  // mcmod_malon.f90:15-23
        return mal->energy(x) - V0;
      case PES_CCPOL: {  // mcmod_waterdimer_ccpol.f90:18-37
        const double ang = 0.529177;
        double xtemp[18];
        for (int i = 0; i < 18; ++i) xtemp[i] = x[i] * ang;
        double Etot = ccpol<double>(*tab, xtemp);
        return (Etot / 627.510) - V0;
      }
    }
    throw std::runtime_error("bad PES kind");
  }

  // subroutine Vprime(x, grad).  CCpol: x is perturbed in place and NOT restored bit-exactly.
  void Vprime(double* x, double* grad) {
    switch (kind) {
      case PES_1D: {  // mcmod_1d.f90:31-32
        for (int i = 0; i < ndim * natom; ++i) {
          double g = ((x[i] / x0_1d) * (x[i] / x0_1d) - 1.0);
          grad[i] = g * 4.0 * Vheight * x[i] / (x0_1d * x0_1d);
        }
        return;
      }
      case PES_2DTEST: {  // mcmod_2dtest.f90:48-58 (24 exp, no CSE in the source)
        potcount++;
        double g1 = 0.0, g2 = 0.0;
        for (int k = 0; k < m; ++k) {
          double dx = x[0] - wx[k], dy = x[1] - wy[k];
          double u = dx * dx + dy * dy;
          g1 = g1 + a0 * dx * pimdk_exp(-a0 * u);
          g1 = g1 + b0 * dx * pimdk_exp(-b0 * u);
          g2 = g2 + a0 * dy * pimdk_exp(-a0 * u);
          g2 = g2 + b0 * dy * pimdk_exp(-b0 * u);
        }
        grad[0] = g1;
        grad[1] = g2;
        return;
      }
      case PES_SO2: {  // mcmod_so2.f90:42-44
        double r = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        grad[0] = (omegaforce * omegaforce) * x[0] * (1.0 - r0 / r);
        grad[1] = (omegaforce * omegaforce) * x[1] * (1.0 - r0 / r);
        return;
      }
      case PES_WATMETH:  // mcmod_watmeth.f90:30-40
        wm.wmrb_grad(x, grad);
        return;
      case PES_MALON:  // mcmod_malon.f90:26-40 (grad(i,j) = gradtemp(ndim*(j-1)+i): the same storage order)
        mal->gradient(x, grad);
        return;
      case PES_CCPOL: {  // mcmod_waterdimer_ccpol.f90:40-58
        const double eps = 1e-4;
        for (int i = 0; i < ndim; ++i)
          for (int j = 0; j < natom; ++j) {
            double& xij = x[j * ndim + i];
            xij = xij + eps;
            double potplus = V(x);
            xij = xij - 2.0 * eps;
            double potminus = V(x);
            xij = xij + eps;
            grad[j * ndim + i] = (potplus - potminus) / (2.0 * eps);
          }
        return;
      }
    }
  }

  // subroutine Vdoubleprime(x, hess): hess(ndim,natom,ndim,natom) column-major,
  // index(i,j,i2,j2) = i + ndim*(j + natom*(i2 + ndim*j2)) (0-based); hess(i,j,:,:) = d grad(:,:) / d x(i,j).
  //   1D    mcmod_1d.f90:37-57    central difference of the analytic gradient, eps = 1e-4, x perturbed in place
  //   2D    mcmod_2dtest.f90:63-86  "analytic", but the four elements are ASSIGNED inside the loop over the wells,
  //         so only the last well (k = m) survives, and the cross/diagonal terms reuse dudy/dudx as written.
  //         Restated literally: this is what the reference's detJ sees.
  //   CCpol mcmod_waterdimer_ccpol.f90:59-76  central difference, eps = 1e-5, of the finite-difference Vprime, whose
  //         own in-place perturbation drift of x carries through the whole double loop
  void Vdoubleprime(double* x, double* hess) {
    const int nd = ndim * natom;
    auto H = [&](int i, int j, int i2, int j2) -> double& { return hess[i + ndim * (j + natom * (i2 + ndim * j2))]; };
    if (kind == PES_2DTEST) {
      for (int q = 0; q < nd * nd; ++q) hess[q] = 0.0;
      for (int k = 0; k < m; ++k) {
        double u = (x[0] - wx[k]) * (x[0] - wx[k]) + (x[1] - wy[k]) * (x[1] - wy[k]);
        double dvdu = a0 * pimdk_exp(-a0 * u) + b0 * pimdk_exp(-b0 * u);
        double d2vdu2 = -(a0 * a0) * pimdk_exp(-a0 * u) - (b0 * b0) * pimdk_exp(-b0 * u);
        double dudx = x[0] - wx[k];
        double dudy = x[1] - wy[k];
        H(0, 0, 0, 0) = (d2vdu2 * dudx + dvdu) * dudx;
        H(1, 0, 0, 0) = (d2vdu2 * dudy + dvdu) * dudx;
        H(0, 0, 1, 0) = (d2vdu2 * dudy + dvdu) * dudx;
        H(1, 0, 1, 0) = (d2vdu2 * dudy + dvdu) * dudy;
      }
      return;
    }
    if (kind == PES_SO2) {  // mcmod_so2.f90:76-82, literally (the diagonal (1 - r0/r) term is not in the reference)
      double r = std::sqrt(x[0] * x[0] + x[1] * x[1]);
      double w2 = omegaforce * omegaforce, r3 = r * r * r;
      H(0, 0, 0, 0) = x[0] * x[0] * w2 * r0 / r3;
      H(0, 0, 1, 0) = x[0] * x[1] * w2 * r0 / r3;
      H(1, 0, 0, 0) = x[0] * x[1] * w2 * r0 / r3;
      H(1, 0, 1, 0) = x[1] * x[1] * w2 * r0 / r3;
      return;
    }
    if (kind == PES_MALON) {  // mcmod_malon.f90:43-70: the packed analytic Hessian spread to hess(i1,j1,i2,j2) and its transpose
      std::vector<double> hp(nd * (nd + 1) / 2);
      mal->hessian_packed(x, hp.data());
      int ij = 0;
      for (int d1 = 0; d1 < nd; ++d1)
        for (int d2 = 0; d2 <= d1; ++d2) {
          const int i1 = d1 % ndim, j1 = d1 / ndim, i2 = d2 % ndim, j2 = d2 / ndim;
          H(i1, j1, i2, j2) = hp[ij];
          H(i2, j2, i1, j1) = hp[ij];
          ++ij;
        }
      return;
    }
    const double eps = (kind == PES_CCPOL) ? 1e-5 : 1e-4;
    std::vector<double> gp(nd), gm(nd);
    for (int i = 0; i < ndim; ++i)
      for (int j = 0; j < natom; ++j) {
        double& xij = x[j * ndim + i];
        xij = xij + eps;
        Vprime(x, gp.data());
        xij = xij - 2.0 * eps;
        Vprime(x, gm.data());
        xij = xij + eps;
        for (int j2 = 0; j2 < natom; ++j2)
          for (int i2 = 0; i2 < ndim; ++i2) H(i, j, i2, j2) = (gp[j2 * ndim + i2] - gm[j2 * ndim + i2]) / (2.0 * eps);
      }
  }
};

}  // namespace oracle
