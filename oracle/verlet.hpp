// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of `module verletint` (verletmodule.f90) and of the ring-polymer algebra the
// hot path uses from `module instantonmod` (instantonmod.f90:17-151 UM*, :397-603 splines),
// keeping the reference's operation count: 10*ndof dense mat-vecs per Langevin step, 8*ndof per
// Andersen step, sin/cos evaluated per element in step_nm, one Vprime call per bead.
//
// Parity status: UNPINNED by the reference (it ships no tests or vectors for this module,
// SURVEY §8c).  Pinned here by mathematical identities in tests/ (T*T=I, spring-matrix
// eigenvalues, leggauss, CubicSpline, energy conservation).  dsymv is MKL in the reference
// (summation order unknown); restated as row dot products.  RNG: see philox.hpp.
//
// Arrays are Fortran column-major x(n,ndim,natom): idx(i,j,k) = (i-1) + n*((j-1) + ndim*(k-1)).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "pes.hpp"
#include "philox.hpp"

namespace oracle {

static const double PI_TRUNC = 3.14159265358979;  // instantonmod.f90:4

// gauleg, verletmodule.f90:124-160
inline void gauleg(double x1, double x2, double* x, double* w, int nintegral) {
  const double EPS = 3.e-14;
  int m = (nintegral + 1) / 2;
  double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (int i = 1; i <= m; ++i) {
    double z = std::cos(PI_TRUNC * (i - 0.25) / (nintegral + 0.5));
    double z1, pp;
    do {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= nintegral; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = nintegral * (z * p1 - p2) / (z * z - 1.0);
      z1 = z;
      z = z1 - p1 / pp;
    } while (std::fabs(z - z1) > EPS);
    x[i - 1] = xm - xl * z;
    x[nintegral - i] = xm + xl * z;
    w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
    w[nintegral - i] = w[i - 1];
  }
}

// spline (natural when yp > 0.99e30) + tridag, instantonmod.f90:397-481
inline void spline(const double* x, const double* y, int n, double yp1, double ypn, double* y2) {
  std::vector<double> a(n + 1), b(n + 1), c(n + 1), r(n + 1), gam(n + 1);
  // 1-based like the reference
  for (int i = 1; i <= n - 1; ++i) c[i] = x[i] - x[i - 1];
  for (int i = 1; i <= n - 1; ++i) r[i] = 6.0 * ((y[i] - y[i - 1]) / c[i]);
  for (int i = n - 1; i >= 2; --i) r[i] = r[i] - r[i - 1];  // r(2:n-1)=r(2:n-1)-r(1:n-2) (array semantics)
  for (int i = 2; i <= n - 1; ++i) a[i] = c[i - 1];
  for (int i = 2; i <= n - 1; ++i) b[i] = 2.0 * (c[i] + a[i]);
  b[1] = 1.0;
  b[n] = 1.0;
  if (yp1 > 0.99e30) {
    r[1] = 0.0;
    c[1] = 0.0;
  } else {
    r[1] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
    c[1] = 0.5;
  }
  if (ypn > 0.99e30) {
    r[n] = 0.0;
    a[n] = 0.0;
  } else {
    r[n] = (-3.0 / (x[n - 1] - x[n - 2])) * ((y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]) - ypn);
    a[n] = 0.5;
  }
  // tridag(a(2:n), b(1:n), c(1:n-1), r(1:n), u(1:n)); inside, a is re-based: a_t(j-1) = a(j)
  std::vector<double> u(n + 1);
  double bet = b[1];
  if (bet == 0.0) throw std::runtime_error("tridag_ser: Error at code stage 1");
  u[1] = r[1] / bet;
  for (int j = 2; j <= n; ++j) {
    gam[j] = c[j - 1] / bet;
    bet = b[j] - a[j] * gam[j];
    if (bet == 0.0) throw std::runtime_error("tridag_ser: Error at code stage 2");
    u[j] = (r[j] - a[j] * u[j - 1]) / bet;
  }
  for (int j = n - 1; j >= 1; --j) u[j] = u[j] - gam[j + 1] * u[j + 1];
  for (int j = 1; j <= n; ++j) y2[j - 1] = u[j];
}

// locate, instantonmod.f90:560-596 (returns 1-based j; 0 or n when out of range)
inline int locate(const double* xx, int n, double x) {
  bool ascnd = (xx[n - 1] >= xx[0]);
  int jl = 0, ju = n + 1;
  while (ju - jl > 1) {
    int jm = (ju + jl) / 2;
    if (ascnd == (x >= xx[jm - 1])) jl = jm;
    else ju = jm;
  }
  if (x == xx[0]) return 1;
  if (x == xx[n - 1]) return n - 1;
  return jl;
}

// splint / splin_grad, instantonmod.f90:500-556
inline double splint(const double* xa, const double* ya, const double* y2a, int n, double x) {
  int klo = std::max(std::min(locate(xa, n, x), n - 1), 1);
  int khi = klo + 1;
  double h = xa[khi - 1] - xa[klo - 1];
  if (h == 0.0) throw std::runtime_error("bad xa input in splint");
  double a = (xa[khi - 1] - x) / h, b = (x - xa[klo - 1]) / h;
  return a * ya[klo - 1] + b * ya[khi - 1] +
         ((a * a * a - a) * y2a[klo - 1] + (b * b * b - b) * y2a[khi - 1]) * (h * h) / 6.0;
}
inline double splin_grad(const double* xa, const double* ya, const double* y2a, int n, double x) {
  int klo = std::max(std::min(locate(xa, n, x), n - 1), 1);
  int khi = klo + 1;
  double h = xa[khi - 1] - xa[klo - 1];
  if (h == 0.0) throw std::runtime_error("bad xa input in splin_grad");
  double a = (xa[khi - 1] - x) / h, b = (x - xa[klo - 1]) / h;
  return ((ya[khi - 1] - ya[klo - 1]) / h) +
         ((1.0 - 3.0 * a * a) * y2a[klo - 1] + (3.0 * b * b - 1.0) * y2a[khi - 1]) * h / 6.0;
}

struct Verlet {
  // mcmod_mass / instantonmod / verletint module state
  int n = 0, ndim = 0, natom = 0, ndof = 0;
  std::vector<double> mass;  // mass(natom)
  double betan = 0, tau = 1.0, gamma = 1.0, dt = 1e-3;
  long NMC = 0, imin = 0, Noutput = 100000;
  bool cayley = false, fixedends = true;
  std::vector<double> transmatrix, beadvec, beadmass, lam, c1, c2;  // (n,n) (n,ndof) (natom,n) (n) (natom,n)x2
  Pes* pes = nullptr;
  // RNG contract
  uint64_t seed = 0;
  uint32_t traj_gid = 0;
  bool nan_trap = false;

  inline size_t IX(int i, int j, int k) const { return (size_t)(i - 1) + (size_t)n * ((j - 1) + (size_t)ndim * (k - 1)); }
  inline double& T(int i, int l) { return transmatrix[(size_t)(l - 1) * n + (i - 1)]; }
  inline double& BM(int j, int i) { return beadmass[(size_t)(i - 1) * natom + (j - 1)]; }
  inline double& BV(int i, int dof) { return beadvec[(size_t)(dof - 1) * n + (i - 1)]; }
  inline double& C1(int j, int i) { return c1[(size_t)(i - 1) * natom + (j - 1)]; }
  inline double& C2(int j, int i) { return c2[(size_t)(i - 1) * natom + (j - 1)]; }

  void setup(int n_, int ndim_, int natom_, const double* mass_, double betan_, Pes* p) {
    n = n_; ndim = ndim_; natom = natom_; ndof = ndim * natom; betan = betan_; pes = p;
    mass.assign(mass_, mass_ + natom);
    transmatrix.assign((size_t)n * n, 0.0);
    beadvec.assign((size_t)n * ndof, 0.0);
    beadmass.assign((size_t)natom * n, 0.0);
    lam.assign(n, 0.0);
  }

  // init_nm, verletmodule.f90:306-338.  a,b are (ndim,natom).
  void init_nm(const double* a, const double* b) {
    for (int i = 1; i <= n; ++i) {
      for (int j = 1; j <= natom; ++j) {
        lam[i - 1] = 2.0 * std::sin((double)i * PI_TRUNC / (double)(2 * n + 2)) / betan;
        BM(j, i) = mass[j - 1] * ((lam[i - 1] * tau) * (lam[i - 1] * tau));
      }
      for (int l = i; l <= n; ++l) {
        T(i, l) = std::sin((double)((long)i * l) * PI_TRUNC / (double)(n + 1)) * std::sqrt(2.0 / (double)(n + 1));
        T(l, i) = T(i, l);
        if (T(i, l) != T(i, l)) throw std::runtime_error("Nan!");
      }
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          int dofi = (k - 1) * ndim + j;
          double v = a[(k - 1) * ndim + (j - 1)] * std::sin((double)i * PI_TRUNC / (double)(n + 1)) +
                     b[(k - 1) * ndim + (j - 1)] * std::sin((double)((long)n * i) * PI_TRUNC / (double)(n + 1));
          v = v * std::sqrt(2.0 / (double)(n + 1));
          v = v / ((lam[i - 1] * betan) * (lam[i - 1] * betan));
          BV(i, dofi) = v;
        }
    }
  }

  // dsymv('U', n, 1, T, n, x, 1, 0, y, 1): y = T x
  void symv(const double* x, double* y) const {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      const double* row = &transmatrix[(size_t)i * n];  // symmetric: column i == row i
      for (int l = 0; l < n; ++l) s += row[l] * x[l];
      y[i] = s;
    }
  }
  // nmtransform_forward / backward, verletmodule.f90:254-286 (vectors have stride 1 here)
  void nm_forward(const double* xprop, double* qprop, int bead) {
    symv(xprop, qprop);
    if (bead > 0)
      for (int i = 1; i <= n; ++i) qprop[i - 1] = qprop[i - 1] - BV(i, bead);
  }
  void nm_backward(double* qprop, double* xprop, int bead) {
    if (bead > 0)
      for (int i = 1; i <= n; ++i) qprop[i - 1] = qprop[i - 1] + BV(i, bead);
    symv(qprop, xprop);
    if (bead > 0)
      for (int i = 1; i <= n; ++i) qprop[i - 1] = qprop[i - 1] - BV(i, bead);
  }

  // step_nm, verletmodule.f90:494-561 (transform = .true.)
  void step_nm(double time, double* x, double* p) {
    size_t tot = (size_t)n * ndof;
    std::vector<double> newpi(tot), q(tot), pip(tot);
    for (int i = 1; i <= ndim; ++i)
      for (int j = 1; j <= natom; ++j) {
        int dofi = (j - 1) * ndim + i;
        nm_forward(&p[IX(1, i, j)], &pip[IX(1, i, j)], 0);
        nm_forward(&x[IX(1, i, j)], &q[IX(1, i, j)], dofi);
      }
    for (int i = 1; i <= n; ++i)
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          size_t e = IX(i, j, k);
          double bm = BM(k, i);
          double omegak = std::sqrt(mass[k - 1] / bm) * lam[i - 1];
          if (cayley) {
            newpi[e] = pip[e] * (4.0 - (omegak * omegak) * (time * time)) - 4.0 * q[e] * bm * (omegak * omegak) * time;
            newpi[e] = newpi[e] / (4.0 + (omegak * omegak) * (time * time));
            q[e] = q[e] * (4.0 - (omegak * omegak) * (time * time)) + 4.0 * pip[e] * time / bm;
            q[e] = q[e] / (4.0 + (omegak * omegak) * (time * time));
          } else {
            newpi[e] = pip[e] * std::cos(time * omegak) - q[e] * omegak * bm * std::sin(omegak * time);
            q[e] = q[e] * std::cos(time * omegak) + pip[e] * std::sin(omegak * time) / (omegak * bm);
          }
          if (newpi[e] != newpi[e]) nan_trap = true;  // "NaN in 1st NM propagation"
        }
    pip = newpi;
    for (int i = 1; i <= ndim; ++i)
      for (int j = 1; j <= natom; ++j) {
        int dofi = (j - 1) * ndim + i;
        nm_backward(&pip[IX(1, i, j)], &p[IX(1, i, j)], 0);
        nm_backward(&q[IX(1, i, j)], &x[IX(1, i, j)], dofi);
      }
  }

  // step_v, verletmodule.f90:565-585 (recalculate = .true.)
  void step_v(double time, double* x, double* p) {
    std::vector<double> xb(ndof), g(ndof);
    for (int i = 1; i <= n; ++i) {
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) xb[(k - 1) * ndim + (j - 1)] = x[IX(i, j, k)];
      pes->Vprime(xb.data(), g.data());
      // Vprime works on the slice x(i,:,:) itself: in-place FD perturbation drift is kept
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) x[IX(i, j, k)] = xb[(k - 1) * ndim + (j - 1)];
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          size_t e = IX(i, j, k);
          p[e] = p[e] - g[(k - 1) * ndim + (j - 1)] * time;
          if (p[e] != p[e]) nan_trap = true;  // "NaN in pot propagation"
        }
    }
  }

  // c1,c2 as in propagate_pimd_pile, verletmodule.f90:381-386
  void setup_pile() {
    c1.assign((size_t)natom * n, 0.0);
    c2.assign((size_t)natom * n, 0.0);
    for (int i = 1; i <= n; ++i)
      for (int j = 1; j <= natom; ++j) {
        C1(j, i) = std::exp(-gamma * dt * lam[i - 1] * std::sqrt(mass[j - 1] / BM(j, i)));
        C2(j, i) = std::sqrt(1.0 - C1(j, i) * C1(j, i));
      }
  }

  // step_langevin, verletmodule.f90:634-667; noise for step `step` from the RNG contract
  void step_langevin(double* pprop, uint64_t step) {
    size_t tot = (size_t)n * ndof;
    std::vector<double> p(tot);
    for (int i = 1; i <= ndim; ++i)
      for (int j = 1; j <= natom; ++j) {
        int dofi = (j - 1) * ndim + i;
        nm_forward(&pprop[IX(1, i, j)], &p[IX(1, i, j)], 0);
        for (int k = 1; k <= n; ++k) {
          double pk = normal_at(seed, STREAM_LANGEVIN, step, traj_gid, (uint64_t)(dofi - 1) * n + (k - 1));
          size_t e = IX(k, i, j);
          p[e] = (C1(j, k) * C1(j, k)) * p[e] +
                 std::sqrt(BM(j, k) / betan) * C2(j, k) * std::sqrt(1.0 + C1(j, k) * C1(j, k)) * pk;
        }
      }
    for (int i = 1; i <= ndim; ++i)
      for (int j = 1; j <= natom; ++j) nm_backward(&p[IX(1, i, j)], &pprop[IX(1, i, j)], 0);
  }

  // time_step_pile :423-435 / time_step_nm :291-302
  void time_step_pile(double* x, double* p, uint64_t step) {
    step_v(dt, x, p);
    step_nm(0.5 * dt, x, p);
    step_langevin(p, step);
    step_nm(0.5 * dt, x, p);
  }
  void time_step_nm(double* x, double* p) {
    step_nm(0.5 * dt, x, p);
    step_v(dt, x, p);
    step_nm(0.5 * dt, x, p);
  }

  // momentum (re)sampling shared by init_path :102-115 and the Andersen kick :213-228
  void sample_momenta(double* p, int stream, uint64_t step) {
    std::vector<double> vel(n), tempp(n);
    for (int i = 1; i <= ndim; ++i)
      for (int k = 1; k <= natom; ++k) {
        int dofi = (k - 1) * ndim + i;
        double stdev = std::sqrt(1.0 / betan);
        for (int j = 1; j <= n; ++j) {
          double z = normal_at(seed, stream, step, traj_gid, (uint64_t)(dofi - 1) * n + (j - 1));
          vel[j - 1] = (0.0 + stdev * z) * std::sqrt(BM(k, j));
        }
        nm_backward(vel.data(), tempp.data(), 0);
        for (int j = 1; j <= n; ++j) p[IX(j, i, k)] = tempp[j - 1];
      }
  }

  // init_path, verletmodule.f90:32-119 (readhess=.false.).  path(npath,ndim,natom) + splines.
  void init_path(double xi, const double* lampath, const double* path, const double* splinepath, int npath,
                 double* x, double* p, uint64_t step = 0) {
    std::vector<double> ya(npath), y2(npath);
    for (int i = 1; i <= ndim; ++i)
      for (int j = 1; j <= natom; ++j) {
        size_t off = (size_t)npath * ((i - 1) + (size_t)ndim * (j - 1));
        for (int k = 1; k <= n; ++k) {
          double xieff = (double)(k - 1) * xi / (double)(n - 1);
          x[IX(k, i, j)] = splint(lampath, path + off, splinepath + off, npath, xieff);
        }
      }
    sample_momenta(p, STREAM_INIT, step);
  }

  // estimator contribution, verletmodule.f90:397-403
  double contr(const double* x, const double* dbdl) const {
    double c = 0.0;
    for (int j = 1; j <= ndim; ++j)
      for (int k = 1; k <= natom; ++k)
        c = c + mass[k - 1] * (-x[IX(n, j, k)]) * dbdl[(k - 1) * ndim + (j - 1)];
    return c;
  }

  // propagate_pimd_pile, verletmodule.f90:372-416 (restart<2, iprint=.false.)
  // dHdrlimit (namelist MCDATA, pimd_par.f90:45,88) and what init_path needs for the outlier re-initialisation :404-409
  double dHdrlimit = -1.0, rp_xi = 0.0;
  std::vector<double> rp_lam, rp_path, rp_spl;
  double propagate_pimd_pile(double* x, double* p, const double* dbdl) {
    setup_pile();
    double dHdr = 0.0;
    for (long ii = 1; ii <= NMC; ++ii) {
      time_step_pile(x, p, (uint64_t)ii);
      if (ii > imin) {
        const double c = contr(x, dbdl);
        if (std::fabs(c) < dHdrlimit || dHdrlimit < 0.0) {
          dHdr = dHdr + c;
        } else {   // "Over limit ... reinitialize path": the contribution is dropped, the momenta are drawn afresh
                   // (RNG contract: stream 0 at the current step; the initial init_path used step 0)
          init_path(rp_xi, rp_lam.data(), rp_path.data(), rp_spl.data(), (int)rp_lam.size(), x, p, (uint64_t)ii);
        }
      }
    }
    return dHdr / (double)(NMC - imin);
  }

  // propagate_pimd_nm, verletmodule.f90:190-250
  double propagate_pimd_nm(double* x, double* p, const double* dbdl) {
    long count = 0;
    double dHdr = 0.0;
    int rkick = poisson_norm(seed, 0, traj_gid, (double)Noutput);
    for (long ii = 1; ii <= NMC; ++ii) {
      count = count + 1;
      if (count >= rkick) {
        count = 0;
        sample_momenta(p, STREAM_ANDERSEN, (uint64_t)ii);
        rkick = poisson_norm(seed, (uint64_t)ii, traj_gid, (double)Noutput);
      }
      time_step_nm(x, p);
      if (ii > imin) dHdr = dHdr + contr(x, dbdl);
    }
    return dHdr / (double)(NMC - imin);
  }

  // UM, instantonmod.f90:17-46
  double UM(const double* x, const double* a, const double* b) {
    double um = 0.0;
    std::vector<double> xb(ndof);
    auto bead = [&](int i) {
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) xb[(k - 1) * ndim + (j - 1)] = x[IX(i, j, k)];
      return xb.data();
    };
    for (int i = 1; i <= n - 1; ++i) {
      double pot = pes->V(bead(i));
      um = um + pot;
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          double d = x[IX(i + 1, j, k)] - x[IX(i, j, k)];
          um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d * d);
        }
    }
    um = um + pes->V(bead(n));
    if (fixedends)
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          double d1 = x[IX(1, j, k)] - a[(k - 1) * ndim + (j - 1)];
          um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d1 * d1);
          double d2 = b[(k - 1) * ndim + (j - 1)] - x[IX(n, j, k)];
          um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d2 * d2);
        }
    return um;
  }

  // UMhessian, instantonmod.f90:155-217 (no inithess): answer(ndof+1, totdof) column-major, LAPACK lower band
  // storage as handed to DSBEVD('L', totdof, kd = ndof).  Literal restatement, including: the spring coupling
  // -1/betan**2 is written at answer(ndof+1, fulldof1) for beads i > 1 (so the bead 1 - bead 2 coupling is
  // absent and the last bead's entry lies outside the matrix), and every diagonal carries 2/betan**2.
  // x (n,ndim,natom) is perturbed in place by the PES's Vdoubleprime exactly like the reference's x(i,:,:).
  void UMhessian(double* x, bool singlewell, double* answer) {
    const int totdof = n * ndof, ld = ndof + 1;
    for (long q = 0; q < (long)ld * totdof; ++q) answer[q] = 0.0;
    std::vector<double> hess((size_t)ndof * ndof, 0.0), xb(ndof);
    auto Hs = [&](int j1, int k1, int j2, int k2) { return hess[(j1 - 1) + ndim * ((k1 - 1) + natom * ((j2 - 1) + ndim * (k2 - 1)))]; };
    for (int i = 1; i <= n; ++i) {
      if ((i == 1 && singlewell) || !singlewell) {
        for (int j = 1; j <= ndim; ++j)
          for (int k = 1; k <= natom; ++k) xb[(k - 1) * ndim + (j - 1)] = x[IX(i, j, k)];
        pes->Vdoubleprime(xb.data(), hess.data());
        for (int j = 1; j <= ndim; ++j)
          for (int k = 1; k <= natom; ++k) x[IX(i, j, k)] = xb[(k - 1) * ndim + (j - 1)];
      }
      for (int j1 = 1; j1 <= ndim; ++j1)
        for (int k1 = 1; k1 <= natom; ++k1)
          for (int j2 = 1; j2 <= ndim; ++j2)
            for (int k2 = 1; k2 <= natom; ++k2) {
              const int idof1 = (k1 - 1) * ndim + j1, idof2 = (k2 - 1) * ndim + j2;
              const int fulldof1 = ndof * (i - 1) + idof1, fulldof2 = ndof * (i - 1) + idof2;
              if (fulldof2 < fulldof1) continue;
              auto A = [&](int r, int c) -> double& { return answer[(r - 1) + (long)ld * (c - 1)]; };
              if (idof1 == idof2) {
                A(1, fulldof1) = 2.0 / (betan * betan) + Hs(j2, k2, j1, k1) / std::sqrt(mass[k1 - 1] * mass[k2 - 1]);
                if (i > 1) A(ndof + 1, fulldof1) = -1.0 / (betan * betan);
              } else {
                const int index = 1 + fulldof2 - fulldof1;
                if (index < 0) continue;
                A(index, fulldof1) = Hs(j2, k2, j1, k1) / std::sqrt(mass[k1 - 1] * mass[k2 - 1]);
              }
            }
    }
  }

  // spring part of the gradient shared by UMprime :59-77 and UMforceenergy :117-134
  double spring_grad(const double* x, const double* a, const double* b, int i, int j, int k) const {
    double m = mass[k - 1], bn2 = betan * betan;
    if (i == 1) {
      if (fixedends) return m * (2.0 * x[IX(1, j, k)] - a[(k - 1) * ndim + (j - 1)] - x[IX(2, j, k)]) / bn2;
      return m * (x[IX(1, j, k)] - x[IX(2, j, k)]) / bn2;
    } else if (i == n) {
      if (fixedends) return m * (2.0 * x[IX(n, j, k)] - x[IX(n - 1, j, k)] - b[(k - 1) * ndim + (j - 1)]) / bn2;
      return m * (x[IX(n, j, k)] - x[IX(n - 1, j, k)]) / bn2;
    }
    return m * (2.0 * x[IX(i, j, k)] - x[IX(i - 1, j, k)] - x[IX(i + 1, j, k)]) / bn2;
  }

  // UMprime, instantonmod.f90:50-89 (x is intent(in): Vprime sees a copy of the bead here)
  void UMprime(const double* x, double* answer, const double* a, const double* b) {
    std::vector<double> xb(ndof), g(ndof);
    for (int i = 1; i <= n; ++i) {
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) answer[IX(i, j, k)] = spring_grad(x, a, b, i, j, k);
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) xb[(k - 1) * ndim + (j - 1)] = x[IX(i, j, k)];
      pes->Vprime(xb.data(), g.data());
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) answer[IX(i, j, k)] = answer[IX(i, j, k)] + g[(k - 1) * ndim + (j - 1)];
    }
  }

  // UMforceenergy, instantonmod.f90:93-151 with potforce(x,grad,energy) = (Vprime, V) of the plugin
  double UMforceenergy(const double* x, double* answer, const double* a, const double* b) {
    std::vector<double> xb(ndof), g(ndof);
    double um = 0.0;
    for (int i = 1; i <= n; ++i) {
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) xb[(k - 1) * ndim + (j - 1)] = x[IX(i, j, k)];
      double energy = pes->V(xb.data());
      pes->Vprime(xb.data(), g.data());
      um = um + energy;
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k)
          if (i < n) {
            double d = x[IX(i + 1, j, k)] - x[IX(i, j, k)];
            um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d * d);
          }
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k)
          answer[IX(i, j, k)] = spring_grad(x, a, b, i, j, k) + g[(k - 1) * ndim + (j - 1)];
    }
    if (fixedends)
      for (int j = 1; j <= ndim; ++j)
        for (int k = 1; k <= natom; ++k) {
          double d1 = x[IX(1, j, k)] - a[(k - 1) * ndim + (j - 1)];
          um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d1 * d1);
          double d2 = b[(k - 1) * ndim + (j - 1)] - x[IX(n, j, k)];
          um = um + (0.5 * mass[k - 1] / (betan * betan)) * (d2 * d2);
        }
    return um;
  }
};

}  // namespace oracle
