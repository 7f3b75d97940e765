// Analytic gradient of the CCpol-8sf[2012, Radau f=1] water-dimer energy (opt-in mode PIMDK_MODE_ANALYTIC).
//
// The reference has no analytic gradient: Vprime (mcmod_waterdimer_ccpol.f90:40-58) is a central difference of 36 energies.
// This header differentiates the SAME energy expression (main_CCpol-8sf.f:210-380, proc_sapt5sf_new_ncd.f,
// proc_ccpol8s-dimer_xyz_ncd.f, H2O.pjt2.f) by the chain rule in reverse, at the cost of about two energies:
//
//   * the energy is invariant under rigid motions of the dimer, and align_on_z_axis (main_CCpol-8sf.f:443-573) is a rigid
//     motion, so V(x) equals the same model evaluated in the laboratory frame with each monomer embedded about its own
//     centre of mass; the alignment needs no derivative;
//   * site-pair sums (poten's 8 x 8 pairs, U0's 625 + 25 + 9 pairs) are sums of functions of one distance and, for
//     SAPT-5s'f, of the monomers' symmetry coordinates: their partial derivatives are written out by hand
//     (d/dr of a Tang-Toennies factor 1 - e^{-br} sum_{i<=n} (br)^i/i! is b e^{-br} (br)^n / n!);
//   * the embedded rigid monomer is a rigid body carried by the Radau frame (vecI, vecJ, vecI x vecJ) and the monomer's
//     centre of mass: its 8 + 25 sites are fixed linear combinations of those four vectors (coefficients obtained once by
//     running put_rigid / set_sites / fill_sites on the canonical frame), so their adjoints reduce to 12 numbers per monomer;
//   * the iterated induction energy (indN_iter) is differentiated at its fixed point:
//     dE = - sum_i mu_i . dE0_i - mu_A . dT_AB . mu_B;
//   * the small per-monomer pieces that remain (Radau frame, set_sites of the flexible monomer, PJT2) are propagated
//     forward with three tangents per atom (dual numbers Dn<3>) and contracted with the adjoints.
//
// Everything here is __host__ __device__ so that tests/ can run the same text on the CPU against the oracle's dual-number
// gradient (oracle/dual.hpp).  It is NOT the reference's arithmetic (no operation-order contract, contraction allowed):
// results agree with the finite-difference default to its truncation error (~3e-8 of max|grad|).
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/pimdk_detmath.h"
#include "ccpol_tables.h"

#if defined(__CUDACC__)
#define PIMDK_AG __host__ __device__ __forceinline__
#define PIMDK_AG_NOINLINE __host__ __device__ __noinline__
#else
#define PIMDK_AG inline
#define PIMDK_AG_NOINLINE inline
#endif

namespace pimdk {
namespace agrad {

// ---------------------------------------------------------------- forward-mode dual numbers ------
template <int N>
struct Dn {
  double v;
  double d[N];
  PIMDK_AG Dn() {}
  PIMDK_AG Dn(double x) : v(x) {
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
};
#define PIMDK_AG_LOOP for (int i = 0; i < N; ++i)
template <int N> PIMDK_AG Dn<N> operator+(const Dn<N>& a, const Dn<N>& b) { Dn<N> r; r.v = a.v + b.v; PIMDK_AG_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> PIMDK_AG Dn<N> operator-(const Dn<N>& a, const Dn<N>& b) { Dn<N> r; r.v = a.v - b.v; PIMDK_AG_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> PIMDK_AG Dn<N> operator*(const Dn<N>& a, const Dn<N>& b) { Dn<N> r; r.v = a.v * b.v; PIMDK_AG_LOOP r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> PIMDK_AG Dn<N> operator/(const Dn<N>& a, const Dn<N>& b) {
  Dn<N> r;
  const double ib = 1.0 / b.v;
  r.v = a.v * ib;
  PIMDK_AG_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N> PIMDK_AG Dn<N> operator-(const Dn<N>& a) { Dn<N> r; r.v = -a.v; PIMDK_AG_LOOP r.d[i] = -a.d[i]; return r; }
template <int N> PIMDK_AG Dn<N> operator+(const Dn<N>& a, double b) { Dn<N> r = a; r.v = a.v + b; return r; }
template <int N> PIMDK_AG Dn<N> operator+(double b, const Dn<N>& a) { return a + b; }
template <int N> PIMDK_AG Dn<N> operator-(const Dn<N>& a, double b) { Dn<N> r = a; r.v = a.v - b; return r; }
template <int N> PIMDK_AG Dn<N> operator-(double b, const Dn<N>& a) { Dn<N> r; r.v = b - a.v; PIMDK_AG_LOOP r.d[i] = -a.d[i]; return r; }
template <int N> PIMDK_AG Dn<N> operator*(const Dn<N>& a, double b) { Dn<N> r; r.v = a.v * b; PIMDK_AG_LOOP r.d[i] = a.d[i] * b; return r; }
template <int N> PIMDK_AG Dn<N> operator*(double b, const Dn<N>& a) { return a * b; }
template <int N> PIMDK_AG Dn<N> operator/(const Dn<N>& a, double b) { return a * (1.0 / b); }
template <int N> PIMDK_AG Dn<N> operator/(double b, const Dn<N>& a) { return Dn<N>(b) / a; }
template <int N> PIMDK_AG Dn<N> chain(const Dn<N>& a, double value, double slope) { Dn<N> r; r.v = value; PIMDK_AG_LOOP r.d[i] = slope * a.d[i]; return r; }
#undef PIMDK_AG_LOOP

PIMDK_AG double ag_sqrt(double x) { return sqrt(x); }
PIMDK_AG double ag_exp(double x) { return pimdk_exp(x); }
PIMDK_AG double ag_cos(double x) { return pimdk_cos(x); }
PIMDK_AG double ag_acos(double x) { return pimdk_acos(x); }
PIMDK_AG double ag_tanh(double x) { return pimdk_tanh(x); }
PIMDK_AG double ag_val(double x) { return x; }
template <int N> PIMDK_AG Dn<N> ag_sqrt(const Dn<N>& a) { const double s = sqrt(a.v); return chain(a, s, 0.5 / s); }
template <int N> PIMDK_AG Dn<N> ag_exp(const Dn<N>& a) { const double e = pimdk_exp(a.v); return chain(a, e, e); }
template <int N> PIMDK_AG Dn<N> ag_cos(const Dn<N>& a) { return chain(a, pimdk_cos(a.v), -pimdk_sin(a.v)); }
template <int N> PIMDK_AG Dn<N> ag_acos(const Dn<N>& a) { return chain(a, pimdk_acos(a.v), -1.0 / sqrt(1.0 - a.v * a.v)); }
template <int N> PIMDK_AG Dn<N> ag_tanh(const Dn<N>& a) { const double t = pimdk_tanh(a.v); return chain(a, t, 1.0 - t * t); }
template <int N> PIMDK_AG double ag_val(const Dn<N>& a) { return a.v; }

template <class R>
PIMDK_AG R ag_ipow(R x, int n) {   // binary powering like the oracle's ipow
  R result = R(1.0);
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

// ---------------------------------------------------------------- tables of the analytic mode ----
// Built on the host by build_grad_tab() (below) from the CcpolDev image; staged into shared memory by the kernels.
struct CcpolGradTab {
  double sapt_abc[8][3];    // embedded-rigid SAPT-5s'f site k (Angstrom) = COM + a I + b J + c K, K = I x J
  double s_rig[3];          // symmetry coordinates of the rigid monomer (round-off of zero)
  double cc_abc[25][3];     // CCpol-8s site k (bohr) = COM/a0 + a I + b J + c K
  double bin5[36][5];       // per bin of U0's sweep: beta, c(nl), c(nl+36), c(nl+72), c(nl+108)
  uint8_t pair_bin[625];    // bin of the site pair (nsA, nsB) -> [nsB*25 + nsA]
  uint8_t pad_[7];
};

constexpr double kA0 = 0.529177249;          // Angstrom per bohr inside CCpol / SAPT-5s'f
constexpr double kAngPlugin = 0.529177;      // the plugin's own factor (mcmod_waterdimer_ccpol.f90:20)
constexpr double kHar2Kcal = 627.510;

PIMDK_AG int site_type(int i) { return (0x43322110 >> (4 * i)) & 0xf; }   // set_sites :1748-1755, 0-based types

// ---------------------------------------------------------------- per-monomer leaves (templated) --
// COMcalc (main_CCpol-8sf.f:575-595)
template <class R>
PIMDK_AG void comcalc_t(const R* O, const R* H1, const R* H2, R* COM) {
  const double mO = 15.9949146221, mH = 1.0078250321;
  const double M = mO + mH + mH;
  for (int i = 0; i < 3; ++i) COM[i] = (mO * O[i] + mH * H1[i] + mH * H2[i]) / M;
}

// radau_f1_tst (main_CCpol-8sf.f:719-810): atoms relative to the monomer's centre of mass -> vecI (bisector), vecJ
template <class R>
PIMDK_AG void radau_f1_t(const R* r0, const R* r1, const R* r2, R* vecI, R* vecJ) {
  const double xmO = 15.9949146221, xmH = 1.0078250321;
  const double xm12 = 2.0 * xmH;
  const double xm = xm12 + xmO;
  const double alpha = sqrt(xmO / xm);
  const double b = (alpha - alpha * alpha) * xm / xm12;
  R q1[3], q2[3], bv[3], t2[3];
  for (int j = 0; j < 3; ++j) {
    q1[j] = r1[j] - b * r0[j];
    q2[j] = r2[j] - b * r0[j];
  }
  R xq1 = ag_sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2]);
  R xq2 = ag_sqrt(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2]);
  for (int j = 0; j < 3; ++j) bv[j] = q1[j] / xq1 + q2[j] / xq2;
  R sss = ag_sqrt(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2]);
  for (int j = 0; j < 3; ++j) vecI[j] = bv[j] / sss;
  R dot = vecI[0] * q2[0] + vecI[1] * q2[1] + vecI[2] * q2[2];
  for (int j = 0; j < 3; ++j) t2[j] = q2[j] - dot * vecI[j];
  R ttt = ag_sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
  for (int j = 0; j < 3; ++j) vecJ[j] = -(t2[j] / ttt);
}

// put_rigid (main_CCpol-8sf.f:391-435): the reference monomer carried by (vecI, vecJ), relative to the centre of mass
PIMDK_AG void put_rigid(const double* vi1, const double* vi2, double* O, double* H1, double* H2) {
  const double ds = 0.79170358110560535, dc = 0.61090542612139243, rOHref = 0.97162570027717354,
               com_shift = 0.66429466101803e-01;
  for (int j = 0; j < 3; ++j) {
    const double w1 = rOHref * (dc * vi1[j] + ds * vi2[j]);
    const double w2 = rOHref * (dc * vi1[j] - ds * vi2[j]);
    const double vshift = -com_shift * vi1[j];
    H1[j] = w1 + vshift;
    H2[j] = w2 + vshift;
    O[j] = 0.0 + vshift;
  }
}

// set_sites (proc_sapt5sf_new_ncd.f:1574-1758).  c[a][j]: atom a (O, H1, H2) in bohr; sites[k*3+j] (Angstrom), s[3].
template <class R>
PIMDK_AG void set_sites_t(const R (&c)[3][3], R* sites, R* s) {
  const double a0 = kA0, r0_ang = 0.9716257, theta0_deg = 104.69;
  const double sig2 = 0.371792435, sig3 = 0.2067213, sig4 = 0.125368076, sig5 = 0.2;
  const double shift = 9.01563628739252e-4;
  const double pi = pimdk_acos(-1.0);
  const double rad2d = 180.0 / pi;
  R v1[3], v2[3], vn1[3], vn2[3], v[3], vb[3], vp[3];
  for (int j = 0; j < 3; ++j) {
    v1[j] = c[1][j] - c[0][j];
    v2[j] = c[2][j] - c[0][j];
  }
  R xnv1 = ag_sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  R xnv2 = ag_sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  for (int j = 0; j < 3; ++j) {
    vn1[j] = v1[j] / xnv1;
    vn2[j] = v2[j] / xnv2;
    v[j] = vn1[j] + vn2[j];
  }
  R xnv = ag_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int j = 0; j < 3; ++j) vb[j] = v[j] / xnv;
  v[0] = v1[1] * v2[2] - v1[2] * v2[1];
  v[1] = v1[2] * v2[0] - v1[0] * v2[2];
  v[2] = v1[0] * v2[1] - v1[1] * v2[0];
  R xn = ag_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int j = 0; j < 3; ++j) vp[j] = v[j] / xn;
  const double r0 = r0_ang / a0;
  const double theta0 = theta0_deg / rad2d;
  const double cta = pimdk_cos(0.5 * theta0);
  R prodv1vb = v1[0] * vb[0] + v1[1] * vb[1] + v1[2] * vb[2];
  R prodv2vb = v2[0] * vb[0] + v2[1] * vb[1] + v2[2] * vb[2];
  R bunny = (0.5 * (prodv1vb + prodv2vb)) / (r0 * cta);
  const double xm16 = 15.994915, xm1 = 1.007825;
  const double sm = xm16 + 2.0 * xm1;
  for (int j = 0; j < 3; ++j) {
    sites[0 + j] = c[0][j] * a0;
    sites[3 + j] = c[1][j] * a0;
    sites[6 + j] = c[2][j] * a0;
    sites[9 + j] = (c[0][j] + (sig3 * vp[j] + sig2 * vb[j] * bunny)) * a0;
    sites[12 + j] = (c[0][j] + (sig2 * vb[j] * bunny - sig3 * vp[j])) * a0;
    sites[15 + j] = (c[0][j] + (sig5 * vp[j] - sig4 * vb[j] * bunny)) * a0;
    sites[18 + j] = (c[0][j] - (sig5 * vp[j] + sig4 * vb[j] * bunny)) * a0;
    R vsm = (xm16 * c[0][j] + xm1 * c[1][j] + xm1 * c[2][j]) / sm;
    sites[21 + j] = (vsm - shift * vb[j]) * a0;
  }
  R sprod = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
  R theta1 = ag_acos(sprod / (xnv1 * xnv2));
  R xnv1_ang = xnv1 * a0, xnv2_ang = xnv2 * a0;
  const double dsqrt2 = sqrt(2.0);
  s[0] = ((xnv1_ang - r0_ang) + (xnv2_ang - r0_ang)) / dsqrt2;
  s[1] = ag_sqrt(xnv1_ang * xnv2_ang) * (theta1 * rad2d - theta0_deg) / rad2d;
  s[2] = ((xnv1_ang - r0_ang) - (xnv2_ang - r0_ang)) / dsqrt2;
}

// POTS (H2O.pjt2.f:1-146, -r8 literals): Q1, Q2 in bohr, THETA in rad -> Hartree
template <class R>
PIMDK_AG_NOINLINE R pots_t(R Q1, R Q2, R THETA) {
  const double TOANG = 0.5291772, CMTOAU = 219474.624, X1 = 1.0, RHO1 = 75.50035308;
  const double FA2 = 18902.44193433, FA3 = 1893.99788146, FA4 = 4096.73443772, FA5 = -1959.60113289,
               FA6 = 4484.15893388, FA7 = 4044.55388819, FA8 = -4771.45043545;
  const double RZ = .95792059, A = 2.226;
  const double F1A1 = -6152.40141181, F2A1 = -2902.13912267, F3A1 = -5732.68460689, F4A1 = 953.88760833;
  const double F11 = 42909.88869093, F1A11 = -2767.19197173, F2A11 = -3394.24705517;
  const double F13 = -1031.93055205, F1A13 = 6023.83435258;
  const double F1A111 = 124.23529382, F2A111 = -1282.50661226;
  const double F113 = -1146.49109522, F1A113 = 9884.41685141, F2A113 = 3040.34021836;
  const double F1111 = 2040.96745268, F1113 = -422.03394198, FA1113 = -7238.09979404;
  const double F11111 = -4969.24544932, F111111 = 8108.49652354, F71 = 90.0;
  const double c1 = 50.0, c2 = 10.0, beta1 = 22.0, beta2 = 13.5, gammas = 0.05, gammaa = 0.10, delta = 0.85, rhh0 = 1.40;
  const double RHO = RHO1 * 3.141592654 / 180.0;
  R Y1 = X1 - ag_exp(-A * (TOANG * Q1 - RZ));
  R Y3 = X1 - ag_exp(-A * (TOANG * Q2 - RZ));
  R cth = ag_cos(THETA);
  R C = cth + pimdk_cos(RHO);
  R C2 = C * C, C3 = C2 * C, C4 = C2 * C2, C5 = C4 * C, C6 = C4 * C2;
  R V0 = (FA2 + FA3 * C + FA4 * C2 + FA6 * C4 + FA7 * C5) * C2 + (FA8 * C6 + FA5 * C3) * C2;
  R FE1 = F1A1 * C + F2A1 * C2 + F3A1 * C3 + F4A1 * C4;
  R FE11 = F11 + F1A11 * C + F2A11 * C2;
  R FE13 = F13 + F1A13 * C;
  R FE111 = F1A111 * C + F2A111 * C2;
  R FE113 = F113 + F1A113 * C + F2A113 * C2;
  R FE1113 = F1113 + FA1113 * C;
  R Y12 = Y1 * Y1, Y13 = Y12 * Y1, Y14 = Y12 * Y12, Y32 = Y3 * Y3, Y33 = Y32 * Y3, Y34 = Y32 * Y32;
  R V = V0 + FE1 * (Y1 + Y3) + FE11 * (Y12 + Y32) + FE13 * Y1 * Y3 + FE111 * (Y13 + Y33) + FE113 * (Y12 * Y3 + Y1 * Y32) +
        F1111 * (Y14 + Y34) + FE1113 * (Y13 * Y3 + Y1 * Y33) + F11111 * (Y14 * Y1 + Y34 * Y3) +
        F111111 * (Y14 * Y12 + Y34 * Y32) + F71 * (Y14 * Y13 + Y34 * Y33);
  const double sqrt2 = sqrt(2.0);
  const double xmup1 = sqrt2 / 3.0 + 0.5;
  const double xmum1 = xmup1 - X1;
  R term = 2.0 * xmum1 * xmup1 * Q1 * Q2 * cth;
  R a1 = xmup1 * Q1, a2 = xmum1 * Q2, b1 = xmum1 * Q1, b2 = xmup1 * Q2;
  R r1 = TOANG * ag_sqrt(a1 * a1 + a2 * a2 - term);
  R r2 = TOANG * ag_sqrt(b1 * b1 + b2 * b2 - term);
  R rhh = ag_sqrt(Q1 * Q1 + Q2 * Q2 - 2.0 * Q1 * Q2 * cth);
  R rbig = (r1 + r2) / sqrt2;
  R rlit = (r1 - r2) / sqrt2;
  R alpha = (X1 - ag_tanh(gammas * rbig * rbig)) * (X1 - ag_tanh(gammaa * rlit * rlit));
  R drhh = TOANG * (rhh - delta * rhh0);
  V = V + c1 * ag_exp(-(beta1 * alpha) * drhh) + c2 * ag_exp(-(beta2 * alpha) * drhh);
  return V / CMTOAU;
}

// One monomer's PJT2 energy (Hartree) and its gradient with respect to the monomer's 9 Cartesian coordinates in Angstrom
// (ccpol :226-262: r = |H - O| / 0.529177249, theta from the bond vectors)
PIMDK_AG double pjt2_monomer(const double* A9, double* g9) {
  double v1[3], v2[3];
  for (int j = 0; j < 3; ++j) {
    v1[j] = A9[3 + j] - A9[j];
    v2[j] = A9[6 + j] - A9[j];
  }
  const double r1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  const double r2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  const double ss = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
  const double ct = ss / (r1 * r2);
  const double th = pimdk_acos(ct);
  Dn<3> Q1(r1 / kA0), Q2(r2 / kA0), TH(th);
  Q1.d[0] = 1.0;
  Q2.d[1] = 1.0;
  TH.d[2] = 1.0;
  const Dn<3> V = pots_t<Dn<3> >(Q1, Q2, TH);
  // chain to the atoms: dr1/dH1 = v1/r1; dtheta/dv1 = -(v2/(r1 r2) - ct v1/r1^2)/sin(theta)
  const double st = sqrt(1.0 - ct * ct);
  for (int j = 0; j < 3; ++j) {
    const double dth1 = -(v2[j] / (r1 * r2) - ct * v1[j] / (r1 * r1)) / st;
    const double dth2 = -(v1[j] / (r1 * r2) - ct * v2[j] / (r2 * r2)) / st;
    const double gH1 = V.d[0] * (v1[j] / r1) / kA0 + V.d[2] * dth1;
    const double gH2 = V.d[1] * (v2[j] / r2) / kA0 + V.d[2] * dth2;
    g9[3 + j] = gH1;
    g9[6 + j] = gH2;
    g9[j] = -(gH1 + gH2);
  }
  return V.v;
}

// ---------------------------------------------------------------- damping -----------------
// Tang-Toennies factor d(n, beta r) (function d / damp) and its derivative with respect to r.  The order is a template
// parameter so that every i in term*br/i is a compile-time reciprocal; the reference's small-argument branch (series tail
// summed until term/sum < 1e-8) is a fixed 12 further terms here, which is at least as many as that criterion takes for the
// arguments that reach it (|d| < 1e-8 means br < 0.2 ... 0.9 for n = 1 ... 10).
template <int N>
PIMDK_AG void tt_damp_d(double beta, double r, double& dd, double& ddr) {
  const double br = beta * r;
  double sum = 1.0, term = 1.0;
#pragma unroll
  for (int i = 1; i <= N; ++i) {
    term = term * br * (1.0 / (double)i);
    sum = sum + term;
  }
  const double e = pimdk_exp(-br);
  ddr = beta * e * term;                       // b e^{-br} (br)^n / n!
  dd = 1.0 - e * sum;
  if (fabs(dd) < 1.0e-8) {
    double t = term, acc = 0.0;
#pragma unroll
    for (int i = N + 1; i <= N + 12; ++i) {
      t = t * br * (1.0 / (double)i);
      acc = acc + t;
    }
    dd = acc * e;
  }
  if (br == 0.0) { dd = 0.0; ddr = 0.0; }
}

// ---------------------------------------------------------------- SAPT-5s'f pair sum with adjoints ----
PIMDK_AG double flex_charge(const double* pa, double s1, double s2, double s3) {
  return pa[0] + pa[1] * s1 + pa[2] * s2 + pa[3] * s3 + pa[4] * s1 * s2 + pa[5] * s2 * s3 + pa[6] * s1 * s1 +
         pa[7] * s2 * s2 + pa[8] * s3 * s3;
}
// adds w * dq/d(s1, s2, s3') to g[0..2] (s3' is the signed third coordinate the charge was evaluated with)
PIMDK_AG void flex_charge_adj(const double* pa, double s1, double s2, double s3, double w, double* g) {
  g[0] += w * (pa[1] + pa[4] * s2 + 2.0 * pa[6] * s1);
  g[1] += w * (pa[2] + pa[4] * s1 + pa[5] * s3 + 2.0 * pa[7] * s2);
  g[2] += w * (pa[3] + pa[5] * s2 + 2.0 * pa[8] * s3);
}

// One site pair of poten (potparts + linear terms, proc_sapt5sf_new_ncd.f:130-213, 238-729), its value and its partial
// derivatives: dr (distance), dx[3] / dy[3] (s1, s2, s3 of A / of B — with respect to the monomers' OWN s3, the sign and
// squaring rules of the site applied), dqa, dqb (the two flexible charges).
struct PairOut {
  double e, dr, dx[3], dy[3], dqa, dqb;
};
template <class Tab>
PIMDK_AG void sapt_pair_adj(const Tab& T, int ia, int ib, double r, const double* sa, const double* sb, double qa,
                            double qb, PairOut& o) {
  const int ta = site_type(ia), tb = site_type(ib);
  const int pt = tb * kNType + ta;
  const int flags = T.pairflags[pt];
  o.e = o.dr = o.dqa = o.dqb = 0.0;
  for (int k = 0; k < 3; ++k) o.dx[k] = o.dy[k] = 0.0;
  if (flags == 0) return;
  const double* pb = &T.parab[pt * kNParab];
#define PB(k) pb[(k)-1]
  const double x1 = sa[0], x2 = sa[1], y1 = sb[0], y2 = sb[1];
  // third coordinate as the pair sees it, and its derivative with respect to the monomer's s3
  const double s3a = (ia == 2) ? -sa[2] : sa[2], s3b = (ib == 2) ? -sb[2] : sb[2];
  const double x3 = (ta != 1) ? s3a * s3a : s3a, y3 = (tb != 1) ? s3b * s3b : s3b;
  const double dx3 = (ta != 1) ? 2.0 * sa[2] : ((ia == 2) ? -1.0 : 1.0);
  const double dy3 = (tb != 1) ? 2.0 * sb[2] : ((ib == 2) ? -1.0 : 1.0);
  double gx3 = 0.0, gy3 = 0.0;                  // dE/dx3, dE/dy3
  double b, al, dbx, dby, dalx, daly;
  if (ta == tb) {
    b = PB(1) + PB(41) * (x3 + y3) + PB(46) * (x3 * x3 + y3 * y3);
    al = PB(2) + PB(43) * (x3 + y3) + PB(48) * (x3 * x3 + y3 * y3);
    dbx = PB(41) + 2.0 * PB(46) * x3; dby = PB(41) + 2.0 * PB(46) * y3;
    dalx = PB(43) + 2.0 * PB(48) * x3; daly = PB(43) + 2.0 * PB(48) * y3;
  } else if (ta < tb) {
    b = PB(1) + PB(41) * x3 + PB(42) * y3 + PB(46) * x3 * x3 + PB(47) * y3 * y3;
    al = PB(2) + PB(43) * x3 + PB(44) * y3 + PB(48) * x3 * x3 + PB(49) * y3 * y3;
    dbx = PB(41) + 2.0 * PB(46) * x3; dby = PB(42) + 2.0 * PB(47) * y3;
    dalx = PB(43) + 2.0 * PB(48) * x3; daly = PB(44) + 2.0 * PB(49) * y3;
  } else {
    b = PB(1) + PB(41) * y3 + PB(42) * x3 + PB(47) * x3 * x3 + PB(46) * y3 * y3;
    al = PB(2) + PB(43) * y3 + PB(44) * x3 + PB(48) * y3 * y3 + PB(49) * x3 * x3;
    dbx = PB(42) + 2.0 * PB(47) * x3; dby = PB(41) + 2.0 * PB(46) * y3;
    dalx = PB(44) + 2.0 * PB(49) * x3; daly = PB(43) + 2.0 * PB(48) * y3;
  }
  const double beta = fabs(b), sgb = b < 0.0 ? -1.0 : 1.0;
  const double rinv = 1.0 / r;
  if (flags & 2) {                               // damped electrostatics d(1, dmp1 r) qa qb / r
    double d1, d1r;
    tt_damp_d<1>(PB(6), r, d1, d1r);
    const double t = d1 * rinv;
    o.e += t * qa * qb;
    o.dqa += t * qb;
    o.dqb += t * qa;
    o.dr += qa * qb * (d1r - t) * rinv;
  }
  if (flags & 4) {                               // damped dispersion - d_n C_n / r^n, n = 6, 8, 10
    const double pm = (ta == tb) ? 0.0 : ((ta < tb) ? 1.0 : -1.0);
    const double r2i = rinv * rinv;
    double rni = r2i * r2i * r2i;                // r^-6
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int n = 6 + 2 * q;
      double dn, dnr;
      if (q == 0) tt_damp_d<6>(PB(7), r, dn, dnr);
      else if (q == 1) tt_damp_d<8>(PB(8), r, dn, dnr);
      else tt_damp_d<10>(PB(9), r, dn, dnr);
      const double cn = PB(3 + q) + PB(11 + q) * (x3 + y3) + PB(14 + q) * (x1 + y1) + PB(17 + q) * (x2 + y2) +
                        PB(20 + q) * (x3 * y3) + PB(23 + q) * (x1 * y1) + PB(26 + q) * (x2 * y2) +
                        pm * (PB(29 + q) * (x3 - y3) + PB(32 + q) * (x1 - y1) + PB(35 + q) * (x2 - y2));
      const double w = -dn * rni;                // dE/dC_n
      o.e += w * cn;
      o.dr += -cn * (dnr - (double)n * dn * rinv) * rni;
      gx3 += w * (PB(11 + q) + PB(20 + q) * y3 + pm * PB(29 + q));
      gy3 += w * (PB(11 + q) + PB(20 + q) * x3 - pm * PB(29 + q));
      o.dx[0] += w * (PB(14 + q) + PB(23 + q) * y1 + pm * PB(32 + q));
      o.dy[0] += w * (PB(14 + q) + PB(23 + q) * x1 - pm * PB(32 + q));
      o.dx[1] += w * (PB(17 + q) + PB(26 + q) * y2 + pm * PB(35 + q));
      o.dy[1] += w * (PB(17 + q) + PB(26 + q) * x2 - pm * PB(35 + q));
      rni *= r2i;
    }
  }
  if ((flags & 1) && beta > 0.0) {               // exponential terms: sum_k C_k val_k, val_k = e^{alpha - beta r} r^k
    const double val0 = pimdk_exp(al) * pimdk_exp(-beta * r);
    const double val[4] = {val0, val0 * r, val0 * r * r, val0 * r * r * r};
    double C[4] = {1.0, PB(38), PB(39), PB(40)};
    const double sym[10] = {x1 + y1, x2 + y2, x3 + y3, x1 * x2 + y1 * y2, x2 * x3 + y2 * y3, x1 * x1 + y1 * y1, x2 * x2 + y2 * y2,
                            x1 * y1, x2 * y2, x3 * y3};
    const double* cs = &T.c[T.itu_s[pt] - 1];
    double U[10];
    for (int g = 0; g < 10; ++g) {
      double u = 0.0;
      for (int k = 0; k < 4; ++k) {
        u += cs[4 * g + k] * val[k];
        C[k] += cs[4 * g + k] * sym[g];
      }
      U[g] = u;
    }
    double W[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (ta != tb) {
      const double sgn = (ta < tb) ? 1.0 : -1.0;
      const double asy[7] = {x1 - y1, x2 - y2, x3 - y3, x1 * x2 - y1 * y2, x2 * x3 - y2 * y3, x1 * x1 - y1 * y1, x2 * x2 - y2 * y2};
      const double* ca = &T.c[T.itu_a[pt] - 1];
      for (int g = 0; g < 7; ++g) {
        double w = 0.0;
        for (int k = 0; k < 4; ++k) {
          w += ca[4 * g + k] * val[k];
          C[k] += sgn * ca[4 * g + k] * asy[g];
        }
        W[g] = sgn * w;
      }
    }
    const double eexp = C[0] * val[0] + C[1] * val[1] + C[2] * val[2] + C[3] * val[3];
    o.e += eexp;
    o.dr += val0 * ((C[1] + r * (2.0 * C[2] + 3.0 * C[3] * r)) - beta * (C[0] + r * (C[1] + r * (C[2] + r * C[3]))));
    const double gal = eexp, gb = -r * eexp * sgb;       // dE/dalpha, dE/db
    gx3 += gal * dalx + gb * dbx;
    gy3 += gal * daly + gb * dby;
    o.dx[0] += U[0] + U[3] * x2 + 2.0 * U[5] * x1 + U[7] * y1 + (W[0] + W[3] * x2 + 2.0 * W[5] * x1);
    o.dx[1] += U[1] + U[3] * x1 + U[4] * x3 + 2.0 * U[6] * x2 + U[8] * y2 + (W[1] + W[3] * x1 + W[4] * x3 + 2.0 * W[6] * x2);
    gx3 += U[2] + U[4] * x2 + U[9] * y3 + (W[2] + W[4] * x2);
    o.dy[0] += U[0] + U[3] * y2 + 2.0 * U[5] * y1 + U[7] * x1 - (W[0] + W[3] * y2 + 2.0 * W[5] * y1);
    o.dy[1] += U[1] + U[3] * y1 + U[4] * y3 + 2.0 * U[6] * y2 + U[8] * x2 - (W[1] + W[3] * y1 + W[4] * y3 + 2.0 * W[6] * y2);
    gy3 += U[2] + U[4] * y2 + U[9] * x3 - (W[2] + W[4] * y2);
  }
#undef PB
  o.dx[2] = gx3 * dx3;
  o.dy[2] = gy3 * dy3;
}

// dipind, per-monomer part (proc_sapt5sf_new_ncd.f:1363-1470): dipole sum over the 8 sites and polarisability.
// The sign of s3 flips cumulatively from site 3 on (:1400-1402): +, +, -, +, -, +, -, +.
PIMDK_AG double dipind_sign(int i) { return (i >= 2 && (i & 1) == 0) ? -1.0 : 1.0; }
template <class Tab>
PIMDK_AG void dipind_monomer(const Tab& T, const double* sites, const double* s, double* dm, double& polis) {
  dm[0] = dm[1] = dm[2] = 0.0;
  for (int i = 0; i < 8; ++i) {
    const double* pa = &T.param[site_type(i) * kNParam];
    const double q = flex_charge(pa, s[0], s[1], dipind_sign(i) * s[2]) / 18.22262373;
    for (int k = 0; k < 3; ++k) dm[k] += q * sites[i * 3 + k] / kA0;
  }
  const double* pa = &T.param[0];
  polis = pa[9] + pa[10] * s[0] + pa[11] * s[1] + pa[12] * s[2] + pa[13] * s[0] * s[1] + pa[14] * s[1] * s[2] +
          pa[15] * s[0] * s[0] + pa[16] * s[1] * s[1] + pa[17] * s[2] * s[2];
}
// adjoint of the above: (adj_dm[3], adj_polis) -> += adj_sites[24], adj_s[3]
template <class Tab>
PIMDK_AG void dipind_monomer_adj(const Tab& T, const double* sites, const double* s, const double* adm, double apol,
                                 double* asites, double* as) {
  for (int i = 0; i < 8; ++i) {
    const double* pa = &T.param[site_type(i) * kNParam];
    const double sg = dipind_sign(i);
    const double q = flex_charge(pa, s[0], s[1], sg * s[2]) / 18.22262373;
    double aq = 0.0;
    for (int k = 0; k < 3; ++k) {
      asites[i * 3 + k] += adm[k] * q / kA0;
      aq += adm[k] * sites[i * 3 + k] / kA0;
    }
    double g[3] = {0.0, 0.0, 0.0};
    flex_charge_adj(pa, s[0], s[1], sg * s[2], aq / 18.22262373, g);
    as[0] += g[0];
    as[1] += g[1];
    as[2] += sg * g[2];
  }
  const double* pa = &T.param[0];
  as[0] += apol * (pa[10] + pa[13] * s[1] + 2.0 * pa[15] * s[0]);
  as[1] += apol * (pa[11] + pa[13] * s[0] + pa[14] * s[2] + 2.0 * pa[16] * s[1]);
  as[2] += apol * (pa[12] + pa[14] * s[1] + 2.0 * pa[17] * s[2]);
}
// dipind, pair part (:1471-1533, R = 0) with adjoints.  u = T(v) m = 3 v (v.m)/r^5 - m/r^3, v = Oa - Ob (both tensors use the
// same v, as the reference calls TTTprod(Oa, Ob, ...) twice).
PIMDK_AG double dipind_pair_adj(double par, const double* Oa, const double* Ob, const double* dma, const double* dmb,
                                double pA, double pB, double* aOa, double* aOb, double* adma, double* admb, double& apA, double& apB) {
  const double k0 = -0.5 * (kA0 * kA0) * ((kA0 * kA0) * (kA0 * kA0)) * kHar2Kcal;
  double v[3];
  for (int k = 0; k < 3; ++k) v[k] = Oa[k] - Ob[k];
  const double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double r = sqrt(r2);
  double dmp, dmpr;
  tt_damp_d<6>(par, r, dmp, dmpr);
  const double r3i = 1.0 / (r2 * r), r5i = r3i / r2, r7i = r5i / r2;
  double etot = 0.0, av[3] = {0.0, 0.0, 0.0};
  for (int w = 0; w < 2; ++w) {
    const double* m = w ? dmb : dma;
    double* am = w ? admb : adma;
    const double pol = w ? pB : pA;
    const double vm = v[0] * m[0] + v[1] * m[1] + v[2] * m[2];
    double u[3];
    for (int k = 0; k < 3; ++k) u[k] = 3.0 * v[k] * vm * r5i - m[k] * r3i;
    const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    const double uv = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    const double um = u[0] * m[0] + u[1] * m[1] + u[2] * m[2];
    etot += pol * uu;
    (w ? apB : apA) = k0 * dmp * uu;
    for (int k = 0; k < 3; ++k) {
      am[k] = k0 * dmp * pol * 2.0 * (3.0 * v[k] * uv * r5i - u[k] * r3i);
      av[k] += pol * 2.0 * (3.0 * u[k] * vm * r5i + 3.0 * uv * m[k] * r5i - 15.0 * uv * vm * v[k] * r7i + 3.0 * um * v[k] * r5i);
    }
  }
  for (int k = 0; k < 3; ++k) {
    const double g = k0 * (dmp * av[k] + etot * dmpr * v[k] / r);
    aOa[k] = g;
    aOb[k] = -g;
  }
  return k0 * etot * dmp;
}

// poten (+ dipind) for one item: sites/s of A and B (Angstrom) -> energy (kcal/mol) and the adjoints
//   adj[0..23] d/d sitesA, adj[24..47] d/d sitesB, adj[48..50] d/d sA, adj[51..53] d/d sB
template <class Tab>
PIMDK_AG double sapt_item_adj(const Tab& T, const double* sitesA, const double* sA, const double* sitesB, const double* sB,
                              double* adj) {
  for (int k = 0; k < 54; ++k) adj[k] = 0.0;
  // dipole-induction term
  double dma[3], dmb[3], pA, pB;
  dipind_monomer(T, sitesA, sA, dma, pA);
  dipind_monomer(T, sitesB, sB, dmb, pB);
  double aOa[3], aOb[3], adma[3], admb[3], apA, apB;
  double e = dipind_pair_adj(T.parab[10 - 1], sitesA, sitesB, dma, dmb, pA, pB, aOa, aOb, adma, admb, apA, apB);
  for (int k = 0; k < 3; ++k) {
    adj[k] += aOa[k];
    adj[24 + k] += aOb[k];
  }
  dipind_monomer_adj(T, sitesA, sA, adma, apA, adj, adj + 48);
  dipind_monomer_adj(T, sitesB, sB, admb, apB, adj + 24, adj + 51);
  // site-pair sum
  double qa[8], qb[8], aqa[8], aqb[8];
  for (int i = 0; i < 8; ++i) {
    qa[i] = flex_charge(&T.param[site_type(i) * kNParam], sA[0], sA[1], (i == 2) ? -sA[2] : sA[2]);
    qb[i] = flex_charge(&T.param[site_type(i) * kNParam], sB[0], sB[1], (i == 2) ? -sB[2] : sB[2]);
    aqa[i] = aqb[i] = 0.0;
  }
  for (int ia = 0; ia < 8; ++ia)
    for (int ib = 0; ib < 8; ++ib) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = sitesA[ia * 3 + k] - sitesB[ib * 3 + k];
      const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      PairOut o;
      sapt_pair_adj(T, ia, ib, r, sA, sB, qa[ia], qb[ib], o);
      e += o.e;
      const double f = o.dr / r;
      for (int k = 0; k < 3; ++k) {
        adj[ia * 3 + k] += f * d[k];
        adj[24 + ib * 3 + k] -= f * d[k];
        adj[48 + k] += o.dx[k];
        adj[51 + k] += o.dy[k];
      }
      aqa[ia] += o.dqa;
      aqb[ib] += o.dqb;
    }
  for (int i = 0; i < 8; ++i) {
    const double sg = (i == 2) ? -1.0 : 1.0;
    double g[3] = {0.0, 0.0, 0.0};
    flex_charge_adj(&T.param[site_type(i) * kNParam], sA[0], sA[1], sg * sA[2], aqa[i], g);
    adj[48] += g[0]; adj[49] += g[1]; adj[50] += sg * g[2];
    double h[3] = {0.0, 0.0, 0.0};
    flex_charge_adj(&T.param[site_type(i) * kNParam], sB[0], sB[1], sg * sB[2], aqb[i], h);
    adj[51] += h[0]; adj[52] += h[1]; adj[53] += sg * h[2];
  }
  return e;
}

// ---------------------------------------------------------------- CCpol-8s rigid model: pair functions ----
// sweep pair (U0 :166-186 + the linear combination :105-110): E = e^{-beta R} (c0 + c1 R + c2 R^2 + c3 R^3)
PIMDK_AG void sweep_pair(const double* b5, double R, double& e, double& dedR) {
  const double ex = pimdk_exp_nonpos(-b5[0] * R);   // beta >= 0 (checked when the tables are built), R >= 0
  const double p = b5[1] + R * (b5[2] + R * (b5[3] + R * b5[4]));
  const double dp = b5[2] + R * (2.0 * b5[3] + 3.0 * b5[4] * R);
  e = ex * p;
  dedR = ex * (dp - b5[0] * p);
}
// damped electrostatics of one charged pair (:190-200)
PIMDK_AG void elst_pair(double d1, double qq, double R, double& e, double& dedR) {
  double f, fr;
  tt_damp_d<1>(d1, R, f, fr);
  e = f * qq / R;
  dedR = qq * (fr - f / R) / R;
}
// damped dispersion of one atom pair (:201-214): - f6 C6/R^6 - f8 C8/R^8 - f10 C10/R^10
PIMDK_AG void disp_pair(const double* dmp3, const double* c3, double R, double& e, double& dedR) {
  const double ri = 1.0 / R, r2i = ri * ri;
  double rni = r2i * r2i * r2i;
  e = 0.0;
  dedR = 0.0;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int n = 6 + 2 * q;
    double f, fr;
    if (q == 0) tt_damp_d<6>(dmp3[0], R, f, fr);
    else if (q == 1) tt_damp_d<8>(dmp3[1], R, f, fr);
    else tt_damp_d<10>(dmp3[2], R, f, fr);
    e -= f * c3[q] * rni;
    dedR -= c3[q] * (fr - (double)n * f * ri) * rni;
    rni *= r2i;
  }
}
// indN_iter for two molecules (:235-372): polarisable centres Rp[2][3], permanent fields E0[2][3] -> Eind (Hartree, the
// reference's last-sweep value) and the converged induced dipoles mu = G2.
PIMDK_AG double ind_solve(const double (&Rp)[2][3], const double (&E0)[2][3], double (&mu)[2][3], int* noconv) {
  const double pol = 9.922;
  double v[3];
  for (int k = 0; k < 3; ++k) {
    v[k] = Rp[0][k] - Rp[1][k];
    mu[0][k] = mu[1][k] = 0.0;
  }
  const double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double r3i = 1.0 / (r2 * sqrt(r2)), r5i = r3i / r2;
  double change = 10.0, Eind = 0.0;
  int it = 0;
  while (change > 1.0e-20 && it < 200) {
    Eind = 0.0;
    change = 0.0;
    for (int i = 0; i < 2; ++i) {
      const int j = 1 - i;
      const double vm = v[0] * mu[j][0] + v[1] * mu[j][1] + v[2] * mu[j][2];   // (v.mu) is even in v: same for both directions
      double E1[3];
      for (int k = 0; k < 3; ++k) E1[k] = E0[i][k] + (3.0 * v[k] * vm * r5i - mu[j][k] * r3i);
      double dot = 0.0;
      for (int k = 0; k < 3; ++k) {
        const double p = pol * E1[k];
        change += (mu[i][k] - p) * (mu[i][k] - p);
        mu[i][k] = p;
        dot += E1[k] * E0[i][k];
      }
      Eind += -0.5 * pol * dot;
    }
    ++it;
  }
  if (it >= 200 && noconv) *noconv = 1;
  return Eind;
}
// adjoint of the induction energy at its fixed point: dE = - sum_i mu_i . dE0_i - mu_0 . dT(v) . mu_1
//   aE0[i][k] = dE/dE0_i,k ; aV[k] = dE/dv_k (v = Rp_0 - Rp_1)
PIMDK_AG void ind_adj(const double (&Rp)[2][3], const double (&mu)[2][3], double (&aE0)[2][3], double* aV) {
  double v[3];
  for (int k = 0; k < 3; ++k) {
    v[k] = Rp[0][k] - Rp[1][k];
    aE0[0][k] = -mu[0][k];
    aE0[1][k] = -mu[1][k];
  }
  const double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double r5i = 1.0 / (r2 * r2 * sqrt(r2)), r7i = r5i / r2;
  const double v0 = v[0] * mu[0][0] + v[1] * mu[0][1] + v[2] * mu[0][2];
  const double v1 = v[0] * mu[1][0] + v[1] * mu[1][1] + v[2] * mu[1][2];
  const double mm = mu[0][0] * mu[1][0] + mu[0][1] * mu[1][1] + mu[0][2] * mu[1][2];
  for (int k = 0; k < 3; ++k)
    aV[k] = -(3.0 * (mu[0][k] * v1 + mu[1][k] * v0) * r5i - 15.0 * v0 * v1 * v[k] * r7i + 3.0 * mm * v[k] * r5i);
}

// ---------------------------------------------------------------- host: tables of the analytic mode ----
// rigid-body coefficients: run the plain constructions on the canonical frame I = x, J = y (K = z), COM = 0
inline const char* build_grad_tab(const CcpolDev& T, CcpolGradTab* g) {
  const double I[3] = {1.0, 0.0, 0.0}, J[3] = {0.0, 1.0, 0.0};
  double O[3], H1[3], H2[3];
  put_rigid(I, J, O, H1, H2);                     // Angstrom, relative to the centre of mass
  {
    double c[3][3], sites[24], s[3];
    for (int j = 0; j < 3; ++j) {
      c[0][j] = O[j] / kA0;
      c[1][j] = H1[j] / kA0;
      c[2][j] = H2[j] / kA0;
    }
    set_sites_t<double>(c, sites, s);
    for (int k = 0; k < 8; ++k)
      for (int j = 0; j < 3; ++j) g->sapt_abc[k][j] = sites[k * 3 + j];
    for (int j = 0; j < 3; ++j) g->s_rig[j] = s[j];
  }
  {  // fill_sites (proc_ccpol8s-dimer_xyz_ncd.f:487-548) on the same rigid monomer, bohr
    const double dv1pv2 = 1.99230765895, dv1mv2 = 2.907303924565;
    double o[3], h1[3], h2[3], com[3], ex[3], ey[3], ez[3];
    for (int j = 0; j < 3; ++j) {
      o[j] = O[j] / kA0;
      h1[j] = H1[j] / kA0;
      h2[j] = H2[j] / kA0;
    }
    comcalc_t<double>(o, h1, h2, com);
    for (int j = 0; j < 3; ++j) {
      const double v1 = h1[j] - com[j], v2 = h2[j] - com[j];
      ez[j] = -(v1 + v2) / dv1pv2;
      ex[j] = (v2 - v1) / dv1mv2;
    }
    ey[0] = ez[1] * ex[2] - ez[2] * ex[1];
    ey[1] = ez[2] * ex[0] - ez[0] * ex[2];
    ey[2] = ez[0] * ex[1] - ez[1] * ex[0];
    for (int k = 0; k < 25; ++k)
      for (int j = 0; j < 3; ++j)
        g->cc_abc[k][j] = ex[j] * T.sites[k * 3 + 0] + ey[j] * T.sites[k * 3 + 1] + ez[j] * T.sites[k * 3 + 2] + com[j];
  }
  for (int b = 0; b < 25; ++b)
    for (int a = 0; a < 25; ++a) {
      const int ib = T.ind_beta[b * 25 + a];
      int indlin = ib - 98;
      if (indlin < 0) indlin += 65;
      if (indlin < 1 || indlin > 36) return "ind_beta maps outside the 36 bins";
      g->pair_bin[b * 25 + a] = (uint8_t)(indlin - 1);
      g->bin5[indlin - 1][0] = T.params[ib - 1];
      for (int p = 0; p < 4; ++p) g->bin5[indlin - 1][1 + p] = T.cc[indlin - 1 + 36 * p];
    }
  return "";
}

}  // namespace agrad
}  // namespace pimdk
