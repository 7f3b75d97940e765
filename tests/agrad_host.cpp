// TEST INFRASTRUCTURE.  The analytic-gradient mathematics of pimd_tunneling_b200/csrc/ccpol_grad.cuh compiled for the
// HOST (the header is __host__ __device__), wired together per bead with plain loops, so that the formulas the CUDA kernels
// run can be checked on the CPU against the oracle's dual-number gradient (oracle/dual.hpp) without a GPU.
// Built by tests/agrad_lib.py:  g++ -O2 -shared -fPIC tests/agrad_host.cpp pimd_tunneling_b200/csrc/ccpol_tables.cpp
#include <cstring>
#include <string>

#include "../pimd_tunneling_b200/csrc/ccpol_grad.cuh"

using namespace pimdk;
using namespace pimdk::agrad;

static CcpolHost g_host;
static CcpolDev g_dev;
static CcpolGradTab g_grad;
static std::string g_err;

namespace {

inline void cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// forward values of one monomer: A9 (Angstrom) -> COM, I, J, K, flexible sites + s
struct Mono {
  double com[3], I[3], J[3], K[3];
  double sites[24], s[3];       // flexible
  double rsites[24];            // embedded-rigid SAPT sites (Angstrom)
  double cc[25][3];             // CCpol-8s sites (bohr)
};
void prep(const double* A9, Mono& m) {
  comcalc_t<double>(A9, A9 + 3, A9 + 6, m.com);
  double rel[9];
  for (int a = 0; a < 3; ++a)
    for (int j = 0; j < 3; ++j) rel[a * 3 + j] = A9[a * 3 + j] - m.com[j];
  radau_f1_t<double>(rel, rel + 3, rel + 6, m.I, m.J);
  cross(m.I, m.J, m.K);
  double c[3][3];
  for (int a = 0; a < 3; ++a)
    for (int j = 0; j < 3; ++j) c[a][j] = A9[a * 3 + j] / kA0;
  set_sites_t<double>(c, m.sites, m.s);
  for (int k = 0; k < 8; ++k)
    for (int j = 0; j < 3; ++j)
      m.rsites[k * 3 + j] = m.com[j] + g_grad.sapt_abc[k][0] * m.I[j] + g_grad.sapt_abc[k][1] * m.J[j] + g_grad.sapt_abc[k][2] * m.K[j];
  for (int k = 0; k < 25; ++k)
    for (int j = 0; j < 3; ++j)
      m.cc[k][j] = m.com[j] / kA0 + g_grad.cc_abc[k][0] * m.I[j] + g_grad.cc_abc[k][1] * m.J[j] + g_grad.cc_abc[k][2] * m.K[j];
}

// adjoint of a rigid body's sites -> (aI, aJ, aK, aCOM): site = scale*COM + a I + b J + c K
void rigid_adj(const double (*abc)[3], int nsite, const double* asite, double scale, double* aI, double* aJ, double* aK, double* aC) {
  for (int k = 0; k < nsite; ++k)
    for (int j = 0; j < 3; ++j) {
      aI[j] += abc[k][0] * asite[k * 3 + j];
      aJ[j] += abc[k][1] * asite[k * 3 + j];
      aK[j] += abc[k][2] * asite[k * 3 + j];
      aC[j] += scale * asite[k * 3 + j];
    }
}

// CCpol-8s rigid model (ccpol8s_dimer without the unit conversion): energy in Hartree, adjoints of the 2 x 25 sites
double rigid_model(const Mono& A, const Mono& B, double (*aA)[3], double (*aB)[3], int* noconv) {
  const CcpolDev& T = g_dev;
  for (int k = 0; k < 25; ++k)
    for (int j = 0; j < 3; ++j) aA[k][j] = aB[k][j] = 0.0;
  double E = 0.0;
  auto add_force = [&](int a, int b, double dedR, const double* d, double R) {
    for (int j = 0; j < 3; ++j) {
      const double f = dedR * d[j] / R;
      aA[a][j] += f;
      aB[b][j] -= f;
    }
  };
  for (int a = 0; a < 25; ++a)
    for (int b = 0; b < 25; ++b) {
      double d[3];
      for (int j = 0; j < 3; ++j) d[j] = A.cc[a][j] - B.cc[b][j];
      const double R = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      double e, de;
      sweep_pair(g_grad.bin5[g_grad.pair_bin[b * 25 + a]], R, e, de);
      E += e;
      add_force(a, b, de, d, R);
      if (a < 5 && b < 5 && (int)T.ind_charge[a] * (int)T.ind_charge[b] != 0) {
        elst_pair(T.params[T.ind_d1[b * 5 + a] - 1], T.params[T.ind_charge[a] - 1] * T.params[T.ind_charge[b] - 1], R, e, de);
        E += e;
        add_force(a, b, de, d, R);
      }
      if (a < 3 && b < 3 && T.ind_d6[b * 3 + a] != 0) {
        const int q = b * 3 + a;
        const double dm[3] = {T.params[T.ind_d6[q] - 1], T.params[T.ind_d8[q] - 1], T.params[T.ind_d10[q] - 1]};
        const double c3[3] = {T.params[T.ind_c6[q] - 1], T.params[T.ind_c8[q] - 1], T.params[T.ind_c10[q] - 1]};
        disp_pair(dm, c3, R, e, de);
        E += e;
        add_force(a, b, de, d, R);
      }
    }
  // induction
  const double sig = 0.367911875040999981, plen = 1.1216873242;
  const double w0 = 1.0 - sig / plen, w12 = 0.5 * sig / plen;
  double Rp[2][3], E0[2][3], mu[2][3];
  for (int m = 0; m < 2; ++m) {
    const double (*s)[3] = m ? B.cc : A.cc;
    for (int j = 0; j < 3; ++j) Rp[m][j] = s[0][j] + sig * (0.5 * (s[1][j] + s[2][j]) - s[0][j]) / plen;
  }
  for (int i = 0; i < 2; ++i) {
    const double (*s)[3] = i ? A.cc : B.cc;     // the OTHER monomer's charged sites
    for (int j = 0; j < 3; ++j) E0[i][j] = 0.0;
    for (int q = 0; q < 5; ++q) {
      double d[3];
      for (int j = 0; j < 3; ++j) d[j] = Rp[i][j] - s[q][j];
      const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const double r3i = 1.0 / (r2 * sqrt(r2));
      for (int j = 0; j < 3; ++j) E0[i][j] += T.chrg[q] * d[j] * r3i;
    }
  }
  E += ind_solve(Rp, E0, mu, noconv);
  double aE0[2][3], aV[3], aRp[2][3];
  ind_adj(Rp, mu, aE0, aV);
  for (int j = 0; j < 3; ++j) {
    aRp[0][j] = aV[j];
    aRp[1][j] = -aV[j];
  }
  for (int i = 0; i < 2; ++i) {
    const double (*s)[3] = i ? A.cc : B.cc;
    double (*as)[3] = i ? aA : aB;
    for (int q = 0; q < 5; ++q) {
      double d[3];
      for (int j = 0; j < 3; ++j) d[j] = Rp[i][j] - s[q][j];
      const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const double r3i = 1.0 / (r2 * sqrt(r2)), r5i = r3i / r2;
      const double ad = aE0[i][0] * d[0] + aE0[i][1] * d[1] + aE0[i][2] * d[2];
      for (int j = 0; j < 3; ++j) {
        const double g = T.chrg[q] * (aE0[i][j] * r3i - 3.0 * d[j] * ad * r5i);   // d(E0 . a)/d d_j
        aRp[i][j] += g;
        as[q][j] -= g;
      }
    }
  }
  for (int m = 0; m < 2; ++m) {
    double (*as)[3] = m ? aB : aA;
    for (int j = 0; j < 3; ++j) {
      as[0][j] += w0 * aRp[m][j];
      as[1][j] += w12 * aRp[m][j];
      as[2][j] += w12 * aRp[m][j];
    }
  }
  return E;
}

// gradient contributions of one monomer's atoms (Angstrom) from the adjoints of its flexible sites / s and of its
// rigid-body frame, by three tangents per atom through comcalc, radau_f1 and set_sites
void back(const double* A9, const double* asite, const double* as, const double* aI, const double* aJ, const double* aK,
          const double* aC, double* g9) {
  typedef Dn<3> D;
  for (int atom = 0; atom < 3; ++atom) {
    D X[9];
    for (int k = 0; k < 9; ++k) X[k] = D(A9[k]);
    for (int t = 0; t < 3; ++t) X[atom * 3 + t].d[t] = 1.0;
    D com[3], rel[9], I[3], J[3], K[3];
    comcalc_t<D>(X, X + 3, X + 6, com);
    for (int a = 0; a < 3; ++a)
      for (int j = 0; j < 3; ++j) rel[a * 3 + j] = X[a * 3 + j] - com[j];
    radau_f1_t<D>(rel, rel + 3, rel + 6, I, J);
    K[0] = I[1] * J[2] - I[2] * J[1];
    K[1] = I[2] * J[0] - I[0] * J[2];
    K[2] = I[0] * J[1] - I[1] * J[0];
    D c[3][3], sites[24], s[3];
    for (int a = 0; a < 3; ++a)
      for (int j = 0; j < 3; ++j) c[a][j] = X[a * 3 + j] / kA0;
    set_sites_t<D>(c, sites, s);
    for (int t = 0; t < 3; ++t) {
      double g = 0.0;
      for (int k = 0; k < 24; ++k) g += asite[k] * sites[k].d[t];
      for (int k = 0; k < 3; ++k) g += as[k] * s[k].d[t] + aI[k] * I[k].d[t] + aJ[k] * J[k].d[t] + aK[k] * K[k].d[t] + aC[k] * com[k].d[t];
      g9[atom * 3 + t] = g;
    }
  }
}

}  // namespace

extern "C" {

const char* agh_last_error() { return g_err.c_str(); }

int agh_load(const char* dir, int isurf, int iemon) {
  const char* m = load_ccpol_tables(dir, isurf, &g_host);
  if (m[0]) { g_err = m; return 1; }
  m = build_ccpol_dev(g_host, iemon, &g_dev);
  if (m[0]) { g_err = m; return 1; }
  m = build_grad_tab(g_dev, &g_grad);
  if (m[0]) { g_err = m; return 1; }
  return 0;
}

// SAPT-5s'f (poten + dipind) of two flexible monomers, Angstrom -> kcal/mol and gradient (18)
int agh_sapt(const double* a9, const double* b9, double* val, double* grad18) {
  Mono A, B;
  prep(a9, A);
  prep(b9, B);
  double adj[54];
  *val = sapt_item_adj(g_dev, A.sites, A.s, B.sites, B.s, adj);
  const double z[3] = {0.0, 0.0, 0.0};
  back(a9, adj, adj + 48, z, z, z, z, grad18);
  back(b9, adj + 24, adj + 51, z, z, z, z, grad18 + 9);
  return 0;
}

// V (Hartree, V0 not subtracted) and dV/dx at x(3,6) in bohr: the whole analytic path
int agh_energy_gradient(const double* x18, int iemon, double* V, double* grad18) {
  double Aa[18];
  for (int i = 0; i < 18; ++i) Aa[i] = x18[i] * kAngPlugin;
  Mono M[2];
  prep(Aa, M[0]);
  prep(Aa + 9, M[1]);
  double g[18];
  for (int i = 0; i < 18; ++i) g[i] = 0.0;
  double emon = 0.0;
  if (iemon) {
    double gm[9];
    for (int m = 0; m < 2; ++m) {
      emon += pjt2_monomer(Aa + 9 * m, gm) * kHar2Kcal;
      for (int k = 0; k < 9; ++k) g[9 * m + k] += gm[k] * kHar2Kcal;
    }
  }
  // flexible and embedded-rigid SAPT-5s'f
  double adjF[54], adjR[54];
  const double val = sapt_item_adj(g_dev, M[0].sites, M[0].s, M[1].sites, M[1].s, adjF);
  const double vall = sapt_item_adj(g_dev, M[0].rsites, g_grad.s_rig, M[1].rsites, g_grad.s_rig, adjR);
  // rigid model
  double aA[25][3], aB[25][3];
  int noconv = 0;
  double erig = 0.0;
  if (g_dev.icc) erig = rigid_model(M[0], M[1], aA, aB, &noconv) * kHar2Kcal;
  for (int m = 0; m < 2; ++m) {
    double aI[3] = {0, 0, 0}, aJ[3] = {0, 0, 0}, aK[3] = {0, 0, 0}, aC[3] = {0, 0, 0};
    double ar[24];
    const double sgn = g_dev.icc ? -1.0 : 0.0;      // Etot = Erigid + (val - vall); SAPT alone: val
    for (int k = 0; k < 24; ++k) ar[k] = sgn * adjR[24 * m + k];
    rigid_adj(g_grad.sapt_abc, 8, ar, 1.0, aI, aJ, aK, aC);
    if (g_dev.icc) {
      double ac[75];
      const double (*src)[3] = m ? aB : aA;
      for (int k = 0; k < 25; ++k)
        for (int j = 0; j < 3; ++j) ac[k * 3 + j] = src[k][j] * kHar2Kcal;
      rigid_adj(g_grad.cc_abc, 25, ac, 1.0 / kA0, aI, aJ, aK, aC);
    }
    double g9[9];
    back(Aa + 9 * m, adjF + 24 * m, adjF + 48 + 3 * m, aI, aJ, aK, aC, g9);
    for (int k = 0; k < 9; ++k) g[9 * m + k] += g9[k];
  }
  const double Etot = (g_dev.icc ? erig + (val - vall) : val) + emon;
  *V = Etot / kHar2Kcal;
  for (int i = 0; i < 18; ++i) grad18[i] = g[i] * kAngPlugin / kHar2Kcal;
  return noconv ? 2 : 0;
}

}  // extern "C"
