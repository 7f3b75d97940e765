// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement, in the reference's own operation order, of one CCpol-8sf energy
//   ccpol / CCpol_xyz / align_on_z_axis / COMcalc3 / radau_f1_tst / eck_rad_tst / put_rigid
//                                   main_CCpol-8sf.f:210-810
//   driver_potss_sapt5sf / poten / potparts(+_old) / d / dipind / TTTprod / scalp / set_sites
//                                   proc_sapt5sf_new_ncd.f:1-729, 733-1224, 1230-1261, 1363-1758
//   ccpol8s_dimer / U0 / indN_iter / efield_bohr / distan / damp / fill_sites / cross / COMcalc
//                                   proc_ccpol8s-dimer_xyz_ncd.f:2-573
//   POTS (PJT2 monomer)             H2O.pjt2.f:1-146
// Everything is a template on the scalar type R so that the same text runs with R=double
// (parity, CPU baseline) and with R=oracle::Counted (exact operation census, opcount.hpp).
// The reference is compiled -r8 -i8 (makefile:5): every real literal is FP64.
//
// Choices where the Fortran leaves the bits to the compiler (documented in DESIGN.md):
//  * exp/pow/sin/cos/acos/tanh come from include/pimdk_detmath.h (see its header for why).
//  * x**k with integer k: binary powering (ipow below);  x**2.d0 -> x*x.
//  * `rin` passed to dipind is never assigned in driver_potss_sapt5sf
//    (proc_sapt5sf_new_ncd.f:43-46) -> restated as 0.
#pragma once
#include <cmath>

#include "../include/pimdk_detmath.h"
#include "tables.hpp"

namespace oracle {

// Elementary functions: the shared deterministic math policy (include/pimdk_detmath.h) stands in
// for the reference's (unknown) Intel libm; sqrt, fabs are IEEE; atan is used by the Eckart
// embedding only (surfaces other than 3/10).
using std::fabs;
using std::sqrt;
inline double exp(double x) { return pimdk_exp(x); }
inline double pow(double x, double y) { return pimdk_pow(x, y); }
inline double sin(double x) { return pimdk_sin(x); }
inline double cos(double x) { return pimdk_cos(x); }
inline double acos(double x) { return pimdk_acos(x); }
inline double atan(double x) { return pimdk_atan(x); }
inline double tanh(double x) { return pimdk_tanh(x); }

template <class R>
inline R ipow(R x, int n) {
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

// ---------------------------------------------------------------- SAPT-5s'f --------------
// function d / function damp  (proc_sapt5sf_new_ncd.f:1230-1261, proc_ccpol8s...:451-485; identical)
template <class R>
R tt_damp(int n, R beta, R r) {
  R br = beta * r;
  if (br == R(0.0)) return R(0.0);
  R sum = R(1.0), term = R(1.0);
  for (int i = 1; i <= n; ++i) {
    term = term * br / R((double)i);
    sum = sum + term;
  }
  R dd = R(1.0) - exp(-br) * sum;
  if (fabs(dd) < R(1.0e-8)) {
    dd = R(0.0);
    for (int i = n + 1; i <= 1000; ++i) {
      term = term * br / R((double)i);
      dd = dd + term;
      if (term / dd < R(1.0e-8)) break;
    }
    dd = dd * exp(-br);
  }
  return dd;
}

// TTTprod, proc_sapt5sf_new_ncd.f:1541-1558
template <class R>
void TTTprod(const R* Ri, const R* Rj, const R* u, R rij, R* v) {
  R ddd = pow(rij, R(0.66666666666666666));
  R scal = R(0.0);
  for (int i = 0; i < 3; ++i) {
    v[i] = Ri[i] - Rj[i];
    scal = scal + v[i] * u[i];
  }
  for (int i = 0; i < 3; ++i) v[i] = (R(3.0) * v[i] * scal * ddd - u[i]) * rij;
}

// set_sites, proc_sapt5sf_new_ncd.f:1574-1758.  carta(i,j): atom i (O,H1,H2), component j; bohr.
template <class R>
void set_sites(const R carta[3][3], R sitea[8][3], R sa[3], int itypea[8]) {
  const R a0 = R(0.529177249);
  const R r0_ang = R(0.9716257);
  const R theta0_deg = R(104.69);
  const R sig2 = R(0.371792435), sig3 = R(0.2067213), sig4 = R(0.125368076), sig5 = R(0.2);
  const R shift = R(9.01563628739252e-4);
  const R pi = acos(R(-1.0));
  const R rad2d = R(180.0) / pi;
  R v1[3], vn1[3], v2[3], vn2[3], v[3], vb[3], vp[3], vsm[3];

  for (int j = 0; j < 3; ++j) {
    sitea[0][j] = carta[0][j];
    sitea[1][j] = carta[1][j];
    sitea[2][j] = carta[2][j];
  }
  for (int j = 0; j < 3; ++j) v1[j] = carta[1][j] - carta[0][j];
  R xnv1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  for (int j = 0; j < 3; ++j) vn1[j] = v1[j] / xnv1;
  R xnv1_ang = xnv1 * a0;
  for (int j = 0; j < 3; ++j) v2[j] = carta[2][j] - carta[0][j];
  R xnv2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  for (int j = 0; j < 3; ++j) vn2[j] = v2[j] / xnv2;
  R xnv2_ang = xnv2 * a0;

  for (int j = 0; j < 3; ++j) v[j] = vn1[j] + vn2[j];
  R xnv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int j = 0; j < 3; ++j) vb[j] = v[j] / xnv;

  v[0] = v1[1] * v2[2] - v1[2] * v2[1];
  v[1] = v1[2] * v2[0] - v1[0] * v2[2];
  v[2] = v1[0] * v2[1] - v1[1] * v2[0];
  R xn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int j = 0; j < 3; ++j) vp[j] = v[j] / xn;

  R r0 = r0_ang / a0;
  R theta0 = theta0_deg / rad2d;
  R cta = cos(R(0.5) * theta0);
  R prodv1vb = v1[0] * vb[0] + v1[1] * vb[1] + v1[2] * vb[2];
  R prodv2vb = v2[0] * vb[0] + v2[1] * vb[1] + v2[2] * vb[2];
  R bunny = (R(0.5) * (prodv1vb + prodv2vb)) / (r0 * cta);  // LAY_CLAUDE = 0

  for (int j = 0; j < 3; ++j) {
    R vd1a = sig3 * vp[j] + sig2 * vb[j] * bunny;
    sitea[3][j] = carta[0][j] + vd1a;
    R vd1b = -sig3 * vp[j] + sig2 * vb[j] * bunny;
    sitea[4][j] = carta[0][j] + vd1b;
    R vd2a = sig5 * vp[j] - sig4 * vb[j] * bunny;
    sitea[5][j] = carta[0][j] + vd2a;
    R vd2b = -sig5 * vp[j] - sig4 * vb[j] * bunny;
    sitea[6][j] = carta[0][j] + vd2b;
  }
  const R xm16 = R(15.994915), xm1 = R(1.007825);
  R sm = xm16 + R(2.0) * xm1;
  for (int j = 0; j < 3; ++j) vsm[j] = (xm16 * carta[0][j] + xm1 * carta[1][j] + xm1 * carta[2][j]) / sm;
  for (int j = 0; j < 3; ++j) sitea[7][j] = vsm[j] - shift * vb[j];

  for (int ia = 0; ia < 8; ++ia)
    for (int i = 0; i < 3; ++i) sitea[ia][i] = sitea[ia][i] * a0;

  R sprod = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
  R ccos = sprod / (xnv1 * xnv2);
  R theta1 = acos(ccos);
  R theta1_deg = theta1 * rad2d;
  R dsqrt2 = sqrt(R(2.0));
  sa[0] = ((xnv1_ang - r0_ang) + (xnv2_ang - r0_ang)) / dsqrt2;
  sa[1] = sqrt(xnv1_ang * xnv2_ang) * (theta1_deg - theta0_deg) / rad2d;
  sa[2] = ((xnv1_ang - r0_ang) - (xnv2_ang - r0_ang)) / dsqrt2;

  static const int types[8] = {1, 2, 2, 3, 3, 4, 4, 5};
  for (int i = 0; i < 8; ++i) itypea[i] = types[i];
}

// potparts / potparts_old, proc_sapt5sf_new_ncd.f:238-729 / 733-1224.
// values[] is 1-based like the Fortran (values[0] unused).
template <class R>
void potparts(const CcpolTables& T, bool old_variant, R rij, int ia, int ib, int& numt, int numtm[2],
              R* values, const R sa[3], const R sb[3], const int itypea[8], const int itypeb[8]) {
  const int ta = itypea[ia - 1], tb = itypeb[ib - 1];
  auto PB = [&](int k) { return R(T.PARAB(k, ta, tb)); };
  auto PA = [&](int k) { return R(T.PARAM(k, ta)); };
  auto PBb = [&](int k) { return R(T.PARAM(k, tb)); };

  R beta = PB(1);
  R alpha = PB(2);
  R a = exp(alpha);  // :277 (recomputed at :367)
  R c6 = PB(3), c8 = PB(4), c10 = PB(5);
  R dmp1 = PB(6), dmp6 = PB(7), dmp8 = PB(8), dmp10 = PB(9);
  R a1 = PB(38), a2 = PB(39), a3 = PB(40);
  R qa = PA(1), qb = PBb(1);

  R s1 = sa[0], s2 = sa[1], s3 = sa[2], s4 = sb[0], s5 = sb[1], s6 = sb[2];
  R signa = R(1.0), signb = R(1.0);
  if (ia == 3) signa = R(-1.0);
  if (ib == 3) signb = R(-1.0);
  s3 = signa * s3;
  s6 = signb * s6;
  qa = qa + PA(2) * s1 + PA(3) * s2 + PA(4) * s3 + PA(5) * s1 * s2 + PA(6) * s2 * s3 + PA(7) * s1 * s1 +
       PA(8) * s2 * s2 + PA(9) * s3 * s3;
  qb = qb + PBb(2) * s4 + PBb(3) * s5 + PBb(4) * s6 + PBb(5) * s4 * s5 + PBb(6) * s5 * s6 +
       PBb(7) * s4 * s4 + PBb(8) * s5 * s5 + PBb(9) * s6 * s6;
  if (ta != 2) s3 = s3 * s3;
  if (tb != 2) s6 = s6 * s6;
  if (ta == tb) {
    beta = beta + PB(41) * (s3 + s6);
    beta = beta + PB(46) * (s3 * s3 + s6 * s6);
  } else if (ta < tb) {
    beta = beta + PB(41) * s3;
    beta = beta + PB(42) * s6;
    beta = beta + PB(46) * s3 * s3;
    beta = beta + PB(47) * s6 * s6;
  } else {
    beta = beta + PB(41) * s6;
    beta = beta + PB(42) * s3;
    beta = beta + PB(47) * s3 * s3;
    beta = beta + PB(46) * s6 * s6;
  }
  beta = fabs(beta);
  if (ta == tb) {
    alpha = alpha + PB(43) * (s3 + s6);
    alpha = alpha + PB(48) * (s3 * s3 + s6 * s6);
  } else if (ta < tb) {
    alpha = alpha + PB(43) * s3;
    alpha = alpha + PB(44) * s6;
    alpha = alpha + PB(48) * s3 * s3;
    alpha = alpha + PB(49) * s6 * s6;
  } else {
    alpha = alpha + PB(43) * s6;
    alpha = alpha + PB(44) * s3;
    alpha = alpha + PB(48) * s6 * s6;
    alpha = alpha + PB(49) * s3 * s3;
  }
  a = exp(alpha);
  R d1 = tt_damp(1, dmp1, rij);
  R d6 = tt_damp(6, dmp6, rij);
  R d8 = tt_damp(8, dmp8, rij);
  R d10 = tt_damp(10, dmp10, rij);
  c6 = c6 + PB(11) * (s3 + s6) + PB(14) * (s1 + s4) + PB(17) * (s2 + s5) + PB(20) * (s3 * s6) +
       PB(23) * (s1 * s4) + PB(26) * (s2 * s5);
  c8 = c8 + PB(12) * (s3 + s6) + PB(15) * (s1 + s4) + PB(18) * (s2 + s5) + PB(21) * (s3 * s6) +
       PB(24) * (s1 * s4) + PB(27) * (s2 * s5);
  c10 = c10 + PB(13) * (s3 + s6) + PB(16) * (s1 + s4) + PB(19) * (s2 + s5) + PB(22) * (s3 * s6) +
        PB(25) * (s1 * s4) + PB(28) * (s2 * s5);
  R c6as = R(0.0), c8as = R(0.0), c10as = R(0.0);
  if (ta != tb) {
    c6as = c6as + PB(29) * (s3 - s6) + PB(32) * (s1 - s4) + PB(35) * (s2 - s5);
    c8as = c8as + PB(30) * (s3 - s6) + PB(33) * (s1 - s4) + PB(36) * (s2 - s5);
    c10as = c10as + PB(31) * (s3 - s6) + PB(34) * (s1 - s4) + PB(37) * (s2 - s5);
    if (ta > tb) {
      c6as = -c6as;
      c8as = -c8as;
      c10as = -c10as;
    }
  }
  c6 = c6 + c6as;
  c8 = c8 + c8as;
  c10 = c10 + c10as;

  if (beta > R(0.0)) {
    numtm[0] = 40;
    numtm[1] = (ta == tb) ? 0 : 28;
    numt = numtm[0] + numtm[1] + 1;
    R val[4];
    val[0] = a * exp(-beta * rij);
    val[1] = val[0] * rij;
    val[2] = val[1] * rij;
    val[3] = val[2] * rij;
    values[numt] = val[0] + a1 * val[1] + a2 * val[2] + a3 * val[3] + d1 * qa * qb / rij -
                   d6 * c6 / ipow(rij, 6) - d8 * c8 / ipow(rij, 8) - d10 * c10 / ipow(rij, 10);
    // symmetric block 1..40 (identical text in the "no H" and "H involved" branches, :453-716)
    R sym[10] = {s1 + s4,           s2 + s5,           s3 + s6,           s1 * s2 + s4 * s5,
                 s2 * s3 + s5 * s6, s1 * s1 + s4 * s4, s2 * s2 + s5 * s5, s1 * s4,
                 s2 * s5,           s3 * s6};
    for (int g = 0; g < 10; ++g)
      for (int k = 0; k < 4; ++k) values[1 + 4 * g + k] = sym[g] * val[k];
    if (old_variant) {  // potparts_old: values(15:16) use s4*s4 (:972-973 of the file)
      R w = s1 * s2 + s4 * s4;
      values[15] = w * val[2];
      values[16] = w * val[3];
    }
    if (ta != tb) {
      R asy[7] = {s1 - s4, s2 - s5, s3 - s6, s1 * s2 - s4 * s5, s2 * s3 - s5 * s6, s1 * s1 - s4 * s4,
                  s2 * s2 - s5 * s5};
      for (int g = 0; g < 7; ++g)
        for (int k = 0; k < 4; ++k) {
          R w = asy[g];
          if (old_variant && g == 3 && k >= 2) w = s1 * s2 - s4 * s4;
          values[41 + 4 * g + k] = (ta < tb) ? w * val[k] : -w * val[k];
        }
    }
  } else {
    numt = 1;
    numtm[0] = 0;
    numtm[1] = 0;
    values[numt] = d1 * qa * qb / rij - d6 * c6 / ipow(rij, 6) - d8 * c8 / ipow(rij, 8) -
                   d10 * c10 / ipow(rij, 10);
  }
}

// dipind, proc_sapt5sf_new_ncd.f:1363-1533
template <class R>
R dipind(const CcpolTables& T, R Rin, const R sa[3], const R sb[3], const R siteat[8][3],
         const R sitebt[8][3], const int itypea[8], const int itypeb[8]) {
  const R a0 = R(0.529177249), har2kcal = R(627.510);
  R dma[3] = {R(0.0), R(0.0), R(0.0)}, dmb[3] = {R(0.0), R(0.0), R(0.0)}, u[3];
  R s1 = sa[0], s2 = sa[1], s3 = sa[2];
  R signa = R(1.0), polisa = R(0.0);
  for (int ia = 1; ia <= 8; ++ia) {
    if (ia == 3) signa = R(-1.0);
    s3 = signa * s3;  // cumulative, :1400-1402
    const int t = itypea[ia - 1];
    R qa = R(T.PARAM(1, t)) + R(T.PARAM(2, t)) * s1 + R(T.PARAM(3, t)) * s2 + R(T.PARAM(4, t)) * s3 +
           R(T.PARAM(5, t)) * s1 * s2 + R(T.PARAM(6, t)) * s2 * s3 + R(T.PARAM(7, t)) * s1 * s1 +
           R(T.PARAM(8, t)) * s2 * s2 + R(T.PARAM(9, t)) * s3 * s3;
    qa = qa / R(18.22262373);
    for (int i = 0; i < 3; ++i) dma[i] = dma[i] + qa * (siteat[ia - 1][i]) / a0;
    if (ia == 1)
      polisa = R(T.PARAM(10, t)) + R(T.PARAM(11, t)) * s1 + R(T.PARAM(12, t)) * s2 +
               R(T.PARAM(13, t)) * s3 + R(T.PARAM(14, t)) * s1 * s2 + R(T.PARAM(15, t)) * s2 * s3 +
               R(T.PARAM(16, t)) * s1 * s1 + R(T.PARAM(17, t)) * s2 * s2 + R(T.PARAM(18, t)) * s3 * s3;
  }
  R s4 = sb[0], s5 = sb[1], s6 = sb[2];
  R signb = R(1.0), polisb = R(0.0);
  for (int ib = 1; ib <= 8; ++ib) {
    if (ib == 3) signb = R(-1.0);
    s6 = signb * s6;
    const int t = itypeb[ib - 1];
    R qb = R(T.PARAM(1, t)) + R(T.PARAM(2, t)) * s4 + R(T.PARAM(3, t)) * s5 + R(T.PARAM(4, t)) * s6 +
           R(T.PARAM(5, t)) * s4 * s5 + R(T.PARAM(6, t)) * s5 * s6 + R(T.PARAM(7, t)) * s4 * s4 +
           R(T.PARAM(8, t)) * s5 * s5 + R(T.PARAM(9, t)) * s6 * s6;
    qb = qb / R(18.22262373);
    for (int i = 0; i < 3; ++i) {
      R Rtemp = (i == 2) ? Rin : R(0.0);
      dmb[i] = dmb[i] + qb * (sitebt[ib - 1][i] - Rtemp) / a0;
    }
    if (ib == 1)
      polisb = R(T.PARAM(10, t)) + R(T.PARAM(11, t)) * s4 + R(T.PARAM(12, t)) * s5 +
               R(T.PARAM(13, t)) * s6 + R(T.PARAM(14, t)) * s4 * s5 + R(T.PARAM(15, t)) * s5 * s6 +
               R(T.PARAM(16, t)) * s4 * s4 + R(T.PARAM(17, t)) * s5 * s5 + R(T.PARAM(18, t)) * s6 * s6;
  }
  R dlen = R(0.0);
  for (int i = 0; i < 3; ++i) {
    R pom = sitebt[0][i] - siteat[0][i];
    dlen = dlen + pom * pom;
  }
  dlen = sqrt(dlen);
  R dmpind = tt_damp(6, R(T.PARAB(10, 1, 1)), dlen);
  dlen = pow(dlen, R(-3.0));
  TTTprod(siteat[0], sitebt[0], dma, dlen, u);
  R energy_a_on_b = polisa * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  TTTprod(siteat[0], sitebt[0], dmb, dlen, u);
  R energy_b_on_a = polisb * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  R energy = energy_a_on_b + energy_b_on_a;
  energy = R(-0.5) * (ipow(a0, 6)) * har2kcal * energy * dmpind;
  return energy;
}

// driver_potss_sapt5sf + poten, proc_sapt5sf_new_ncd.f:1-222.  carta/cartb in Angstrom on entry,
// divided by a0 in place (:36-41).
template <class R>
R sapt5sf(const CcpolTables& T, R carta[3][3], R cartb[3][3]) {
  const R a0 = R(0.529177249);
  for (int ii = 0; ii < 3; ++ii)
    for (int jj = 0; jj < 3; ++jj) {
      carta[ii][jj] = carta[ii][jj] / a0;
      cartb[ii][jj] = cartb[ii][jj] / a0;
    }
  R siteat[8][3], sitebt[8][3], sa[3], sb[3];
  int itypea[8], itypeb[8];
  set_sites(carta, siteat, sa, itypea);
  set_sites(cartb, sitebt, sb, itypeb);

  int itypus[7][7][3] = {};  // itypus(ntypemax,ntypemax,2), 1-based
  int iii = 1;
  R val = R(0.0);
  R values[101];
  int numt, numtm[2];
  for (int ia = 1; ia <= 8; ++ia) {
    for (int ib = 1; ib <= 8; ++ib) {
      R valp = R(0.0);
      R diff, ttt = R(0.0);
      for (int i = 0; i < 3; ++i) {
        diff = siteat[ia - 1][i] - sitebt[ib - 1][i];
        ttt = ttt + diff * diff;
      }
      R rij = sqrt(ttt);
      potparts(T, T.ipotparts == 0, rij, ia, ib, numt, numtm, values, sa, sb, itypea, itypeb);
      valp = valp + values[numt];  // ntpot=124161 > 10: last value is the fixed part (:164-169)
      const int ta = itypea[ia - 1], tb = itypeb[ib - 1];
      int itsmax = (ta != tb) ? 3 : 1;
      for (int its = 1; its <= itsmax; its += 2) {
        int itsm = its < 2 ? its : 2;
        int itu = itypus[ta][tb][itsm];
        if (itu == 0) {
          itypus[ta][tb][itsm] = iii;
          itypus[tb][ta][itsm] = iii;
          itu = iii;
          iii = iii + numtm[itsm - 1];
        }
        for (int i = 1; i <= numtm[itsm - 1]; ++i) {
          int itu1 = itu + i - 1;
          if (itsm == 1)
            valp = valp + R(T.c[itu1 - 1]) * values[i];
          else
            valp = valp + R(T.c[itu1 - 1]) * values[i + numtm[0]];
        }
      }
      val = val + valp;
    }
  }
  R fcind = dipind(T, R(0.0), sa, sb, siteat, sitebt, itypea, itypeb);
  val = val + fcind;
  return val;
}

// ---------------------------------------------------------------- CCpol-8s rigid ---------
// COMcalc / COMcalc3 (identical arithmetic), proc_ccpol8s...:560-572, main_CCpol-8sf.f:575-592
template <class R>
void comcalc(const R* O1, const R* H1, const R* H2, R* COM) {
  const R mO = R(15.9949146221), mH = R(1.0078250321);
  R M = mO + mH + mH;
  for (int i = 0; i < 3; ++i) COM[i] = (mO * O1[i] + mH * H1[i] + mH * H2[i]) / M;
}

// fill_sites, proc_ccpol8s-dimer_xyz_ncd.f:487-548
template <class R>
void fill_sites(const CcpolTables& T, const R* O, const R* H1, const R* H2, R rsites[25][3]) {
  const R dv1pv2 = R(1.99230765895), dv1mv2 = R(2.907303924565);
  R COM[3], v1[3], v2[3], ex[3], ey[3], ez[3];
  comcalc(O, H1, H2, COM);
  for (int j = 0; j < 3; ++j) {
    v1[j] = H1[j] - COM[j];
    v2[j] = H2[j] - COM[j];
  }
  for (int j = 0; j < 3; ++j) {
    ez[j] = -(v1[j] + v2[j]);
    ex[j] = v2[j] - v1[j];
  }
  for (int j = 0; j < 3; ++j) {
    ez[j] = ez[j] / dv1pv2;
    ex[j] = ex[j] / dv1mv2;
  }
  ey[0] = ez[1] * ex[2] - ez[2] * ex[1];  // cross(ey,ez,ex)
  ey[1] = ez[2] * ex[0] - ez[0] * ex[2];
  ey[2] = ez[0] * ex[1] - ez[1] * ex[0];
  for (int kk = 1; kk <= 25; ++kk)
    for (int j = 0; j < 3; ++j)
      rsites[kk - 1][j] = ex[j] * R(T.SITES(1, kk)) + ey[j] * R(T.SITES(2, kk)) + ez[j] * R(T.SITES(3, kk));
  for (int kk = 0; kk < 25; ++kk)
    for (int j = 0; j < 3; ++j) rsites[kk][j] = rsites[kk][j] + COM[j];
}

// efield_bohr, proc_ccpol8s-dimer_xyz_ncd.f:380-421
template <class R>
void efield_bohr(const CcpolTables& T, const R* veci, const R sitebt[25][3], R* e) {
  R sep[25][3], sepl[25];
  for (int is = 0; is < 25; ++is)
    if (T.chrg[is] != 0.0) {
      sepl[is] = R(0.0);
      for (int k = 0; k < 3; ++k) {
        sep[is][k] = veci[k] - sitebt[is][k];
        sepl[is] = sepl[is] + sep[is][k] * sep[is][k];
      }
      sepl[is] = pow(sepl[is], R(-1.5));
    }
  for (int k = 0; k < 3; ++k) e[k] = R(0.0);
  const R a0 = R(1.0);
  for (int is = 0; is < 25; ++is)
    if (T.chrg[is] != 0.0)
      for (int k = 0; k < 3; ++k) e[k] = e[k] + a0 * a0 * R(T.chrg[is]) * sep[is][k] * sepl[is];
}

// indN_iter with N=2, proc_ccpol8s-dimer_xyz_ncd.f:235-372.  Returns false on non-convergence.
template <class R>
bool indN_iter2(const CcpolTables& T, const R sitesA[25][3], const R sitesB[25][3], R& energy,
                int* sweeps = nullptr) {
  const R pol = R(9.922), sig = R(0.367911875040999981), plen = R(1.1216873242), dmpfct = R(1.0);
  const int maxit = 200;
  const R(*mol[2])[3] = {sitesA, sitesB};
  R Rp[2][3], G2[2][3] = {}, E0[2][3], epom[3];
  for (int i = 0; i < 2; ++i)
    for (int ii = 0; ii < 3; ++ii) {
      R pom = R(0.5) * (mol[i][1][ii] + mol[i][2][ii]);
      Rp[i][ii] = mol[i][0][ii] + sig * (pom - mol[i][0][ii]) / plen;
    }
  R dist = R(0.0);
  for (int ii = 0; ii < 3; ++ii) dist = dist + (Rp[0][ii] - Rp[1][ii]) * (Rp[0][ii] - Rp[1][ii]);
  dist = pow(dist, R(-1.5));
  for (int i = 0; i < 2; ++i) {
    for (int k = 0; k < 3; ++k) E0[i][k] = R(0.0);
    int j = 1 - i;
    efield_bohr(T, Rp[i], mol[j], epom);
    for (int k = 0; k < 3; ++k) E0[i][k] = E0[i][k] + epom[k];
  }
  const R thr_iter = R(1.0e-20);
  R change = R(10.0);
  int isteps = 0;
  energy = R(0.0);
  while (change > thr_iter && isteps < maxit) {
    energy = R(0.0);
    change = R(0.0);
    for (int i = 0; i < 2; ++i) {
      R E1[3] = {E0[i][0], E0[i][1], E0[i][2]};
      int j = 1 - i;
      TTTprod(Rp[i], Rp[j], G2[j], dist, epom);
      for (int k = 0; k < 3; ++k) E1[k] = E1[k] + dmpfct * epom[k];
      R polE1[3] = {pol * E1[0], pol * E1[1], pol * E1[2]};
      change = (G2[i][0] - polE1[0]) * (G2[i][0] - polE1[0]) + (G2[i][1] - polE1[1]) * (G2[i][1] - polE1[1]) +
               (G2[i][2] - polE1[2]) * (G2[i][2] - polE1[2]) + change;
      for (int k = 0; k < 3; ++k) G2[i][k] = polE1[k];
      energy = R(-0.5) * pol * (E1[0] * E0[i][0] + E1[1] * E0[i][1] + E1[2] * E0[i][2]) + energy;
    }
    isteps = isteps + 1;
  }
  if (sweeps) *sweeps = isteps;
  return isteps < maxit;
}

// U0, proc_ccpol8s-dimer_xyz_ncd.f:118-233
template <class R>
void U0(const CcpolTables& T, const R sitesA[25][3], const R sitesB[25][3], R& a0out, R aj[145]) {
  const int nlin = 144;
  for (int i = 1; i <= nlin; ++i) aj[i] = R(0.0);
  R E_ele = R(0.0), E_ind = R(0.0);
  for (int nsA = 1; nsA <= 25; ++nsA)
    for (int nsB = 1; nsB <= 25; ++nsB) {
      R d = R(0.0);
      for (int i = 0; i < 3; ++i) {
        R r12 = sitesA[nsA - 1][i] - sitesB[nsB - 1][i];
        d = d + r12 * r12;
      }
      R Rr = sqrt(d);
      const int ij = CcpolTables::IJ(nsA, nsB);
      if (T.ind_beta[ij] != 0) {
        R beta = R(T.params[T.ind_beta[ij] - 1]);
        R eks = exp(-beta * Rr);
        int indlin = T.ind_beta[ij] - 98;
        if (indlin < 0) indlin = indlin + 65;
        int ind0 = indlin, ind1 = ind0 + 36, ind2 = ind1 + 36, ind3 = ind2 + 36;
        aj[ind0] = aj[ind0] + eks;
        aj[ind1] = aj[ind1] + eks * Rr;
        aj[ind2] = aj[ind2] + eks * Rr * Rr;
        aj[ind3] = aj[ind3] + eks * Rr * Rr * Rr;
      }
      if (T.ind_charge[nsA - 1] * T.ind_charge[nsB - 1] != 0) {
        R qA = R(T.params[T.ind_charge[nsA - 1] - 1]);
        R qB = R(T.params[T.ind_charge[nsB - 1] - 1]);
        R d1 = R(T.params[T.ind_d1[ij] - 1]);
        R f1 = tt_damp(1, d1, Rr);
        E_ele = E_ele + f1 * qA * qB / Rr;
      }
      if (T.ind_d6[ij] != 0) {
        R d6 = R(T.params[T.ind_d6[ij] - 1]);
        R d8 = R(T.params[T.ind_d8[ij] - 1]);
        R d10 = R(T.params[T.ind_d10[ij] - 1]);
        R C6 = R(T.params[T.ind_c6[ij] - 1]);
        R C8 = R(T.params[T.ind_c8[ij] - 1]);
        R C10 = R(T.params[T.ind_c10[ij] - 1]);
        R f6 = tt_damp(6, d6, Rr);
        R f8 = tt_damp(8, d8, Rr);
        R f10 = tt_damp(10, d10, Rr);
        R R2 = Rr * Rr;
        R R6 = R2 * R2 * R2;
        R R8 = R6 * R2;
        R R10 = R8 * R2;
        E_ind = E_ind - f6 * C6 / R6 - f8 * C8 / R8 - f10 * C10 / R10;
      }
    }
  a0out = E_ele + E_ind;
}

// ccpol8s_dimer (imode=0), proc_ccpol8s-dimer_xyz_ncd.f:60-116.  Inputs in Angstrom, converted in place.
template <class R>
R ccpol8s_dimer(const CcpolTables& T, R* Oa, R* Ha1, R* Ha2, R* Ob, R* Hb1, R* Hb2, bool* converged) {
  const R bohr2a = R(0.529177249), h2kcal = R(627.510);
  for (int j = 0; j < 3; ++j) {
    Oa[j] = Oa[j] / bohr2a;
    Ha1[j] = Ha1[j] / bohr2a;
    Ha2[j] = Ha2[j] / bohr2a;
    Ob[j] = Ob[j] / bohr2a;
    Hb1[j] = Hb1[j] / bohr2a;
    Hb2[j] = Hb2[j] / bohr2a;
  }
  R sA[25][3], sB[25][3];
  fill_sites(T, Oa, Ha1, Ha2, sA);
  fill_sites(T, Ob, Hb1, Hb2, sB);
  R Eind;
  bool ok = indN_iter2(T, sA, sB, Eind);
  if (converged) *converged = ok;
  R a0, aj[145];
  U0(T, sA, sB, a0, aj);
  R E = Eind;
  for (int nl = 1; nl <= 144; ++nl) E = E + R(T.cc[nl - 1]) * aj[nl];
  E = E + a0;
  return E * h2kcal;
}

// ---------------------------------------------------------------- PJT2 monomer -----------
// POTS, H2O.pjt2.f:1-146 (literals are default REAL promoted by -r8)
template <class R>
R pots(R Q1, R Q2, R THETA, bool r8 = true) {
  // r8=false reproduces a build WITHOUT -r8: default-REAL literals and sqrt(2.0) are single precision.
  // The golden valm(1:10) of main_CCpol-8sf.f:182-183 were generated that way.
  auto L = [&](double x) { return r8 ? R(x) : R((double)(float)x); };
  const R TOANG = L(0.5291772), CMTOAU = L(219474.624), X1 = L(1.0);
  const R RHO1 = L(75.50035308);
  const R FA2 = L(18902.44193433), FA3 = L(1893.99788146), FA4 = L(4096.73443772), FA5 = L(-1959.60113289),
          FA6 = L(4484.15893388), FA7 = L(4044.55388819), FA8 = L(-4771.45043545), FA9 = L(0.0), FA10 = L(0.0);
  const R RZ = L(.95792059), A = L(2.226);
  const R F1A1 = L(-6152.40141181), F2A1 = L(-2902.13912267), F3A1 = L(-5732.68460689), F4A1 = L(953.88760833);
  const R F11 = L(42909.88869093), F1A11 = L(-2767.19197173), F2A11 = L(-3394.24705517);
  const R F13 = L(-1031.93055205), F1A13 = L(6023.83435258);
  const R F111 = L(0.0), F1A111 = L(124.23529382), F2A111 = L(-1282.50661226);
  const R F113 = L(-1146.49109522), F1A113 = L(9884.41685141), F2A113 = L(3040.34021836);
  const R F1111 = L(2040.96745268), FA1111 = L(0.0), F1113 = L(-422.03394198), FA1113 = L(-7238.09979404);
  const R FA1133 = L(0.0), F11111 = L(-4969.24544932), F111111 = L(8108.49652354), F71 = L(90.0);
  const R c1 = L(50.0), c2 = L(10.0), beta1 = L(22.0), beta2 = L(13.5), gammas = L(0.05), gammaa = L(0.10),
          delta = L(0.85), rhh0 = L(1.40);
  const R RHO = RHO1 * L(3.141592654) / L(180.0);
  const R FA11 = L(0.0);
  const R F1A3 = F1A1, F2A3 = F2A1, F3A3 = F3A1, F4A3 = F4A1, F33 = F11, F1A33 = F1A11, F2A33 = F2A11;
  const R F333 = F111, F1A333 = F1A111, F2A333 = F2A111, F133 = F113, F1A133 = F1A113, F2A133 = F2A113;
  const R F3333 = F1111, FA3333 = FA1111, F1333 = F1113, FA1333 = FA1113, F33333 = F11111,
          F333333 = F111111, F73 = F71;

  R DR = TOANG * Q1 - RZ;
  R DS = TOANG * Q2 - RZ;
  R Y1 = X1 - exp(-A * DR);
  R Y3 = X1 - exp(-A * DS);
  R CORO = cos(THETA) + cos(RHO);
  auto P = [&](R x, int n) { return ipow(x, n); };
  R V0 = (FA2 + FA3 * CORO + FA4 * P(CORO, 2) + FA6 * P(CORO, 4) + FA7 * P(CORO, 5)) * P(CORO, 2);
  V0 = V0 + (FA8 * P(CORO, 6) + FA5 * P(CORO, 3) + FA9 * P(CORO, 7) + FA10 * P(CORO, 8)) * P(CORO, 2);
  V0 = V0 + (FA11 * P(CORO, 9)) * P(CORO, 2);
  R FE1 = F1A1 * CORO + F2A1 * P(CORO, 2) + F3A1 * P(CORO, 3) + F4A1 * P(CORO, 4);
  R FE3 = F1A3 * CORO + F2A3 * P(CORO, 2) + F3A3 * P(CORO, 3) + F4A3 * P(CORO, 4);
  R FE11 = F11 + F1A11 * CORO + F2A11 * P(CORO, 2);
  R FE33 = F33 + F1A33 * CORO + F2A33 * P(CORO, 2);
  R FE13 = F13 + F1A13 * CORO;
  R FE111 = F111 + F1A111 * CORO + F2A111 * P(CORO, 2);
  R FE333 = F333 + F1A333 * CORO + F2A333 * P(CORO, 2);
  R FE113 = F113 + F1A113 * CORO + F2A113 * P(CORO, 2);
  R FE133 = F133 + F1A133 * CORO + F2A133 * P(CORO, 2);
  R FE1111 = F1111 + FA1111 * CORO;
  R FE3333 = F3333 + FA3333 * CORO;
  R FE1113 = F1113 + FA1113 * CORO;
  R FE1333 = F1333 + FA1333 * CORO;
  R FE1133 = FA1133 * CORO;
  R FE11111 = F11111, FE33333 = F33333, FE111111 = F111111, FE333333 = F333333, FE71 = F71, FE73 = F73;
  R V = V0 + FE1 * Y1 + FE3 * Y3 + FE11 * P(Y1, 2) + FE33 * P(Y3, 2) + FE13 * Y1 * Y3 + FE111 * P(Y1, 3) +
        FE333 * P(Y3, 3) + FE113 * P(Y1, 2) * Y3 + FE133 * Y1 * P(Y3, 2) + FE1111 * P(Y1, 4) +
        FE3333 * P(Y3, 4) + FE1113 * P(Y1, 3) * Y3 + FE1333 * Y1 * P(Y3, 3) + FE1133 * P(Y1, 2) * P(Y3, 2) +
        FE11111 * P(Y1, 5) + FE33333 * P(Y3, 5) + FE111111 * P(Y1, 6) + FE333333 * P(Y3, 6) +
        FE71 * P(Y1, 7) + FE73 * P(Y3, 7);
  R sqrt2 = r8 ? sqrt(R(2.0)) : L(1.4142135623730951);  // (double)sqrtf(2.0f) when !r8
  R xmup1 = sqrt2 / R(3.0) + R(0.5);
  R xmum1 = xmup1 - X1;
  R term = R(2.0) * xmum1 * xmup1 * Q1 * Q2 * cos(THETA);
  R r1 = TOANG * sqrt(P(xmup1 * Q1, 2) + P(xmum1 * Q2, 2) - term);
  R r2 = TOANG * sqrt(P(xmum1 * Q1, 2) + P(xmup1 * Q2, 2) - term);
  R rhh = sqrt(P(Q1, 2) + P(Q2, 2) - R(2.0) * Q1 * Q2 * cos(THETA));
  R rbig = (r1 + r2) / sqrt2;
  R rlit = (r1 - r2) / sqrt2;
  R alpha = (X1 - tanh(gammas * P(rbig, 2))) * (X1 - tanh(gammaa * P(rlit, 2)));
  R alpha1 = beta1 * alpha;
  R alpha2 = beta2 * alpha;
  R drhh = TOANG * (rhh - delta * rhh0);
  V = V + c1 * exp(-alpha1 * drhh) + c2 * exp(-alpha2 * drhh);
  V = V / CMTOAU;
  return V;
}

// ---------------------------------------------------------------- frame / embedding ------
// align_on_z_axis, main_CCpol-8sf.f:443-573.  Mutates all six atoms; returns Rcom.
template <class R>
R align_on_z_axis(R* O1A, R* H1A, R* H2A, R* O1B, R* H1B, R* H2B) {
  const R thr = R(1.0e-9);
  R xyzA[3][3], xyzB[3][3], xyzAA[3][3], xyzBB[3][3], xyzAAA[3][3], xyzBBB[3][3];
  R comA[3], comB[3], s[3], s1[3];
  for (int j = 0; j < 3; ++j) {
    xyzA[0][j] = O1A[j];
    xyzA[1][j] = H1A[j];
    xyzA[2][j] = H2A[j];
    xyzB[0][j] = O1B[j];
    xyzB[1][j] = H1B[j];
    xyzB[2][j] = H2B[j];
  }
  comcalc(xyzA[0], xyzA[1], xyzA[2], comA);
  comcalc(xyzB[0], xyzB[1], xyzB[2], comB);
  R sss = R(0.0);
  for (int j = 0; j < 3; ++j) sss = sss + (comB[j] - comA[j]) * (comB[j] - comA[j]);
  R Rcom = sqrt(sss);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      xyzA[i][j] = xyzA[i][j] - comA[j];
      xyzB[i][j] = xyzB[i][j] - comA[j];
    }
    comB[i] = comB[i] - comA[i];
  }
  for (int i = 0; i < 3; ++i) comA[i] = comA[i] - comA[i];
  R ss = sqrt(comB[0] * comB[0] + comB[1] * comB[1]);
  if (ss < thr) {
    if (comB[2] < R(0.0)) {
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
          xyzA[i][j] = -xyzA[i][j];
          xyzB[i][j] = -xyzB[i][j];
        }
        comA[i] = -comA[i];
        comB[i] = -comB[i];
      }
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        xyzAAA[i][j] = xyzA[i][j];
        xyzBBB[i][j] = xyzB[i][j];
      }
  } else {
    R xnorm = sqrt(comB[0] * comB[0] + comB[1] * comB[1]);
    s[0] = comB[1] / xnorm;
    s[1] = -comB[0] / xnorm;
    s[2] = R(0.0);
    R rr = sqrt(comB[0] * comB[0] + comB[1] * comB[1] + comB[2] * comB[2]);
    R ccos = comB[2] / rr;
    R ssin = sqrt(R(1.0) - ccos * ccos);
    s1[0] = -comB[0] / xnorm;
    s1[1] = -comB[1] / xnorm;
    s1[2] = R(0.0);
    for (int i = 0; i < 3; ++i) {
      xyzAA[i][0] = xyzA[i][0] * s[0] + xyzA[i][1] * s[1];
      xyzAA[i][1] = xyzA[i][0] * s1[0] + xyzA[i][1] * s1[1];
      xyzAA[i][2] = xyzA[i][2];
    }
    for (int i = 0; i < 3; ++i) {
      xyzBB[i][0] = xyzB[i][0] * s[0] + xyzB[i][1] * s[1];
      xyzBB[i][1] = xyzB[i][0] * s1[0] + xyzB[i][1] * s1[1];
      xyzBB[i][2] = xyzB[i][2];
    }
    for (int i = 0; i < 3; ++i) {
      xyzAAA[i][0] = xyzAA[i][0];
      xyzAAA[i][1] = xyzAA[i][1] * ccos + xyzAA[i][2] * ssin;
      xyzAAA[i][2] = -xyzAA[i][1] * ssin + xyzAA[i][2] * ccos;
    }
    for (int i = 0; i < 3; ++i) {
      xyzBBB[i][0] = xyzBB[i][0];
      xyzBBB[i][1] = xyzBB[i][1] * ccos + xyzBB[i][2] * ssin;
      xyzBBB[i][2] = -xyzBB[i][1] * ssin + xyzBB[i][2] * ccos;
    }
  }
  for (int j = 0; j < 3; ++j) {
    O1A[j] = xyzAAA[0][j];
    H1A[j] = xyzAAA[1][j];
    H2A[j] = xyzAAA[2][j];
    O1B[j] = xyzBBB[0][j];
    H1B[j] = xyzBBB[1][j];
    H2B[j] = xyzBBB[2][j];
  }
  return Rcom;
}

// common head of eck_rad_tst / radau_f1_tst: Radau vectors q1,q2 (main_CCpol-8sf.f:731-770)
template <class R>
void radau_vectors(const R* r0, const R* r1, const R* r2, R q1[3], R q2[3], R& xq1, R& xq2, R& sss) {
  const R xmO = R(15.9949146221), xmH = R(1.0078250321);
  R xm12 = R(2.0) * xmH;
  R xm = xm12 + xmO;
  R alpha = sqrt(xmO / xm);
  R b = (alpha - alpha * alpha) * xm / xm12;
  for (int j = 0; j < 3; ++j) {
    q1[j] = r1[j] - b * r0[j];
    q2[j] = r2[j] - b * r0[j];
  }
  xq1 = R(0.0);
  xq2 = R(0.0);
  sss = R(0.0);
  for (int j = 0; j < 3; ++j) {
    xq1 = xq1 + q1[j] * q1[j];
    xq2 = xq2 + q2[j] * q2[j];
    sss = sss + q1[j] * q2[j];
  }
  xq1 = sqrt(xq1);
  xq2 = sqrt(xq2);
}

// radau_f1_tst, main_CCpol-8sf.f:719-810
template <class R>
void radau_f1_tst(const R* r0, const R* r1, const R* r2, R vecI[3], R vecJ[3]) {
  R q1[3], q2[3], xq1, xq2, sss, bv[3], temp2[3];
  radau_vectors(r0, r1, r2, q1, q2, xq1, xq2, sss);
  R theta_r = acos(sss / (xq1 * xq2));  // computed by the reference, unused here
  (void)theta_r;
  sss = R(0.0);
  for (int j = 0; j < 3; ++j) {
    R pom1 = q1[j] / xq1, pom2 = q2[j] / xq2;
    bv[j] = pom1 + pom2;
    sss = sss + bv[j] * bv[j];
  }
  sss = sqrt(sss);
  for (int j = 0; j < 3; ++j) {
    bv[j] = bv[j] / sss;
    vecI[j] = bv[j];
  }
  sss = R(0.0);
  for (int j = 0; j < 3; ++j) sss = sss + vecI[j] * q2[j];
  R ttt = R(0.0);
  for (int j = 0; j < 3; ++j) {
    temp2[j] = q2[j] - sss * vecI[j];
    ttt = ttt + temp2[j] * temp2[j];
  }
  ttt = sqrt(ttt);
  for (int j = 0; j < 3; ++j) {
    temp2[j] = temp2[j] / ttt;
    vecJ[j] = -temp2[j];
  }
}

// eck_rad_tst, main_CCpol-8sf.f:597-716 (needed only for isurf other than 3 and 10)
template <class R>
void eck_rad_tst(const R* r0, const R* r1, const R* r2, R vecI[3], R vecJ[3]) {
  const R xq1e = R(0.95111822), xq2e = R(0.95111822), theta_r_e = R(1.88412851);
  R q1[3], q2[3], xq1, xq2, sss, temp1[3], temp2[3];
  radau_vectors(r0, r1, r2, q1, q2, xq1, xq2, sss);
  R theta_r = acos(sss / (xq1 * xq2));
  R eta_e = R(0.5) * theta_r_e;
  R ang = theta_r - theta_r_e + eta_e;
  sss = (xq2e * xq2 * sin(ang) + xq1e * xq1 * sin(eta_e)) / (xq2e * xq2 * cos(ang) + xq1e * xq1 * cos(eta_e));
  R eta = atan(sss);
  sss = R(0.0);
  for (int j = 0; j < 3; ++j) {
    temp1[j] = q1[j] / xq1;
    sss = sss + temp1[j] * q2[j];
  }
  R ttt = R(0.0);
  for (int j = 0; j < 3; ++j) {
    temp2[j] = q2[j] - sss * temp1[j];
    ttt = ttt + temp2[j] * temp2[j];
  }
  ttt = sqrt(ttt);
  for (int j = 0; j < 3; ++j) temp2[j] = temp2[j] / ttt;
  for (int j = 0; j < 3; ++j) {
    vecI[j] = cos(eta) * temp1[j] + sin(eta) * temp2[j];
    vecJ[j] = -sin(eta) * temp1[j] + cos(eta) * temp2[j];
  }
}

// put_rigid, main_CCpol-8sf.f:391-435
template <class R>
void put_rigid(const R vi1[3], const R vi2[3], R O[3], R H1[3], R H2[3]) {
  const R ds = R(0.79170358110560535), dc = R(0.61090542612139243), rOHref = R(0.97162570027717354),
          com_shift = R(0.66429466101803e-01);
  R w1[3], w2[3], vshift[3], Opos[3] = {R(0.0), R(0.0), R(0.0)};
  for (int j = 0; j < 3; ++j) {
    w1[j] = dc * vi1[j] + ds * vi2[j];
    w2[j] = dc * vi1[j] - ds * vi2[j];
    vshift[j] = -com_shift * vi1[j];
  }
  for (int j = 0; j < 3; ++j) {
    w1[j] = rOHref * w1[j];
    w2[j] = rOHref * w2[j];
  }
  for (int j = 0; j < 3; ++j) {
    w1[j] = w1[j] + vshift[j];
    w2[j] = w2[j] + vshift[j];
    Opos[j] = Opos[j] + vshift[j];
  }
  for (int j = 0; j < 3; ++j) {
    O[j] = Opos[j];
    H1[j] = w1[j];
    H2[j] = w2[j];
  }
}

// CCpol_xyz, main_CCpol-8sf.f:273-380.  Atoms in Angstrom, mutated like the reference.
template <class R>
R ccpol_xyz(const CcpolTables& T, R* Oa, R* Ha1, R* Ha2, R* Ob, R* Hb1, R* Hb2, bool* converged) {
  R carta[3][3], cartb[3][3], cartaa[3][3], cartbb[3][3];
  R vecIa[3], vecJa[3], vecIb[3], vecJb[3];
  R Oaa[3], Haa1[3], Haa2[3], Obb[3], Hbb1[3], Hbb2[3];
  R Rcom = align_on_z_axis(Oa, Ha1, Ha2, Ob, Hb1, Hb2);
  for (int jj = 0; jj < 3; ++jj) {
    carta[0][jj] = Oa[jj];
    carta[1][jj] = Ha1[jj];
    carta[2][jj] = Ha2[jj];
    cartb[0][jj] = Ob[jj];
    cartb[1][jj] = Hb1[jj];
    cartb[2][jj] = Hb2[jj];
  }
  if (T.iembed == 1) eck_rad_tst(Oa, Ha1, Ha2, vecIa, vecJa);
  if (T.iembed == 2) radau_f1_tst(Oa, Ha1, Ha2, vecIa, vecJa);
  put_rigid(vecIa, vecJa, Oaa, Haa1, Haa2);
  Ob[2] = Ob[2] - Rcom;
  Hb1[2] = Hb1[2] - Rcom;
  Hb2[2] = Hb2[2] - Rcom;
  if (T.iembed == 1) eck_rad_tst(Ob, Hb1, Hb2, vecIb, vecJb);
  if (T.iembed == 2) radau_f1_tst(Ob, Hb1, Hb2, vecIb, vecJb);
  Ob[2] = Ob[2] + Rcom;
  Hb1[2] = Hb1[2] + Rcom;
  Hb2[2] = Hb2[2] + Rcom;
  put_rigid(vecIb, vecJb, Obb, Hbb1, Hbb2);
  Obb[2] = Obb[2] + Rcom;
  Hbb1[2] = Hbb1[2] + Rcom;
  Hbb2[2] = Hbb2[2] + Rcom;
  for (int jj = 0; jj < 3; ++jj) {
    cartaa[0][jj] = Oaa[jj];
    cartaa[1][jj] = Haa1[jj];
    cartaa[2][jj] = Haa2[jj];
    cartbb[0][jj] = Obb[jj];
    cartbb[1][jj] = Hbb1[jj];
    cartbb[2][jj] = Hbb2[jj];
  }
  if (converged) *converged = true;
  if (T.icc == 1) {
    R val = sapt5sf(T, carta, cartb);
    R vall = sapt5sf(T, cartaa, cartbb);
    R Erigid = ccpol8s_dimer(T, Oaa, Haa1, Haa2, Obb, Hbb1, Hbb2, converged);
    return Erigid + (val - vall);
  }
  return sapt5sf(T, carta, cartb);
}

// ccpol, main_CCpol-8sf.f:210-270.  xyz = Oa,Ha1,Ha2,Ob,Hb1,Hb2 (3 each), Angstrom, mutated.
template <class R>
R ccpol(const CcpolTables& T, R xyz[18], bool* converged = nullptr) {
  R *Oa = xyz, *Ha1 = xyz + 3, *Ha2 = xyz + 6, *Ob = xyz + 9, *Hb1 = xyz + 12, *Hb2 = xyz + 15;
  R Etot = ccpol_xyz(T, Oa, Ha1, Ha2, Ob, Hb1, Hb2, converged);
  if (T.iemonomer == 1) {
    R rA1 = R(0.0), rA2 = R(0.0), rB1 = R(0.0), rB2 = R(0.0), ssA = R(0.0), ssB = R(0.0);
    for (int j = 0; j < 3; ++j) {
      rA1 = rA1 + (Ha1[j] - Oa[j]) * (Ha1[j] - Oa[j]);
      rA2 = rA2 + (Ha2[j] - Oa[j]) * (Ha2[j] - Oa[j]);
      rB1 = rB1 + (Hb1[j] - Ob[j]) * (Hb1[j] - Ob[j]);
      rB2 = rB2 + (Hb2[j] - Ob[j]) * (Hb2[j] - Ob[j]);
      ssA = ssA + (Ha1[j] - Oa[j]) * (Ha2[j] - Oa[j]);
      ssB = ssB + (Hb1[j] - Ob[j]) * (Hb2[j] - Ob[j]);
    }
    rA1 = sqrt(rA1);
    rA2 = sqrt(rA2);
    rB1 = sqrt(rB1);
    rB2 = sqrt(rB2);
    R thA = acos(ssA / (rA1 * rA2));
    R thB = acos(ssB / (rB1 * rB2));
    const R a0 = R(0.529177249), h2kcal = R(627.510);
    rA1 = rA1 / a0;
    rA2 = rA2 / a0;
    rB1 = rB1 / a0;
    rB2 = rB2 / a0;
    R vA = pots(rA1, rA2, thA, T.pjt2_r8 != 0);
    R vB = pots(rB1, rB2, thB, T.pjt2_r8 != 0);
    Etot = Etot + (vA + vB) * h2kcal;
  }
  return Etot;
}

}  // namespace oracle
