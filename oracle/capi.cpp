// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Flat C entry points over the C++ restatement so that tests/ and bench.py's cpu_baseline leg
// can drive it through ctypes.  Nothing under pimd_tunneling_b200/ may link or load this.
#include <cstring>
#include <string>

#include "ccpol_impl.hpp"
#include "dual.hpp"
#include "opcount.hpp"
#include "pes.hpp"
#include "tables.hpp"
#include "verlet.hpp"

using namespace oracle;

static CcpolTables g_tab;
static bool g_tab_loaded = false;
static Pes g_pes;
static Verlet g_v;
static std::string g_err;

#define ORC_TRY(body)                 \
  try {                               \
    body;                             \
    return 0;                         \
  } catch (const std::exception& e) { \
    g_err = e.what();                 \
    return 1;                         \
  }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- CCpol tables -------------------------------------------------------------------------
static Malonaldehyde g_mal;
int orc_malon_load(const char* tbl) { ORC_TRY(g_mal.load(tbl)) }
int orc_ccpol_load_text(const char* dir, int isurf, int iemon) {
  ORC_TRY(load_text(dir, isurf, iemon, g_tab); g_tab_loaded = true)
}
int orc_ccpol_load_packed(const char* sapt_tbl, const char* cc8s_tbl, int isurf, int iemon) {
  ORC_TRY(load_packed(sapt_tbl, cc8s_tbl, isurf, iemon, g_tab); g_tab_loaded = true)
}
// raw image of the tables (for text-vs-packed loader equality tests)
long orc_ccpol_tables_image(unsigned char* dst, long cap) {
  long sz = (long)sizeof(CcpolTables);
  if (dst && cap >= sz) std::memcpy(dst, &g_tab, sz);
  return sz;
}
void orc_ccpol_set_pjt2_r8(int r8) { g_tab.pjt2_r8 = r8; }
// `ccpol` (main_CCpol-8sf.f:210): 6 atoms x 3, Angstrom -> kcal/mol.  Input is copied, not mutated.
int orc_ccpol_energy_ang(const double* xyz18, double* E) {
  if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
  double w[18];
  std::memcpy(w, xyz18, sizeof(w));
  bool conv = true;
  *E = ccpol<double>(g_tab, w, &conv);
  if (!conv) { g_err = "No convergence in indN_iter"; return 2; }
  return 0;
}
// pieces, for finer-grained GPU parity tests
int orc_pots(double q1, double q2, double theta, double* v) { *v = pots<double>(q1, q2, theta, g_tab.pjt2_r8 != 0); return 0; }
int orc_sapt5sf_ang(const double* a9, const double* b9, double* val) {
  double ca[3][3], cb[3][3];
  std::memcpy(ca, a9, sizeof(ca));
  std::memcpy(cb, b9, sizeof(cb));
  *val = sapt5sf<double>(g_tab, ca, cb);
  return 0;
}
int orc_ccpol8s_dimer_ang(const double* xyz18, double* e, int* sweeps) {
  double w[18];
  std::memcpy(w, xyz18, sizeof(w));
  bool conv = true;
  *e = ccpol8s_dimer<double>(g_tab, w, w + 3, w + 6, w + 9, w + 12, w + 15, &conv);
  if (sweeps) *sweeps = conv ? 0 : -1;
  return conv ? 0 : 2;
}
// exact operation census of one `ccpol` energy (opcount.hpp); counts[] order = Counted::names()
int orc_ccpol_opcount(const double* xyz18, double* counts, int ncounts, double* E) {
  if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
  Counted w[18];
  for (int i = 0; i < 18; ++i) w[i] = Counted(xyz18[i]);
  Counted::reset();
  Counted e = ccpol<Counted>(g_tab, w);
  if (E) *E = e.v;
  for (int i = 0; i < ncounts && i < Counted::NKIND; ++i) counts[i] = (double)Counted::cnt[i];
  return 0;
}
// V (Hartree) and its ANALYTIC gradient (Hartree/bohr) at x(3,6) in bohr by forward-mode dual numbers (dual.hpp)
// through the same templates: V = ccpol(x * 0.529177)/627.510 (mcmod_waterdimer_ccpol.f90:18-37, V0 not subtracted).
int orc_ccpol_analytic_gradient(const double* x_bohr18, double* V, double* grad18) {
  if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
  Dual w[18];
  for (int i = 0; i < 18; ++i) {
    Dual xi(x_bohr18[i]);
    xi.d[i] = 1.0;
    w[i] = xi * Dual(0.529177);
  }
  bool conv = true;
  Dual e = ccpol<Dual>(g_tab, w, &conv) / Dual(627.510);
  if (!conv) { g_err = "No convergence in indN_iter"; return 2; }
  if (V) *V = e.v;
  for (int i = 0; i < 18; ++i) grad18[i] = e.d[i];
  return 0;
}
// the same for the two site models alone (stage-wise yardsticks for the analytic-gradient kernels):
// SAPT-5s'f(carta, cartb) in kcal/mol and its gradient with respect to the 18 Angstrom coordinates
int orc_sapt5sf_dual(const double* a9, const double* b9, double* val, double* grad18) {
  if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
  Dual ca[3][3], cb[3][3];
  for (int i = 0; i < 9; ++i) {
    ca[i / 3][i % 3] = Dual(a9[i]);
    ca[i / 3][i % 3].d[i] = 1.0;
    cb[i / 3][i % 3] = Dual(b9[i]);
    cb[i / 3][i % 3].d[9 + i] = 1.0;
  }
  Dual e = sapt5sf<Dual>(g_tab, ca, cb);
  *val = e.v;
  for (int i = 0; i < 18; ++i) grad18[i] = e.d[i];
  return 0;
}
// ccpol8s_dimer (kcal/mol) of six atoms in Angstrom and its gradient
int orc_ccpol8s_dual(const double* xyz18, double* val, double* grad18) {
  if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
  Dual w[18];
  for (int i = 0; i < 18; ++i) {
    w[i] = Dual(xyz18[i]);
    w[i].d[i] = 1.0;
  }
  bool conv = true;
  Dual e = ccpol8s_dimer<Dual>(g_tab, w, w + 3, w + 6, w + 9, w + 12, w + 15, &conv);
  *val = e.v;
  for (int i = 0; i < 18; ++i) grad18[i] = e.d[i];
  return conv ? 0 : 2;
}
int orc_opcount_kinds() { return Counted::NKIND; }
const char* orc_opcount_name(int i) { return Counted::name(i); }

// ---- plugin layer (module mcmod_mass) ------------------------------------------------------
int orc_pes_select(const char* name) {
  std::string s(name);
  if (s == "1d") { g_pes = Pes(); g_pes.init_1d(); return 0; }
  if (s == "2dtest") { g_pes = Pes(); g_pes.init_2d(); return 0; }
  if (s == "so2") { g_pes = Pes(); g_pes.init_so2(); return 0; }
  if (s == "watmeth") { g_pes = Pes(); g_pes.init_watmeth(); return 0; }
  if (s == "malon") {
    if (!g_mal.loaded) { g_err = "malonaldehyde tables not loaded (orc_malon_load)"; return 1; }
    g_pes = Pes();
    g_pes.init_malon(&g_mal);
    return 0;
  }
  if (s == "ccpol8sf") {
    if (!g_tab_loaded) { g_err = "tables not loaded"; return 1; }
    g_pes = Pes();
    g_pes.init_ccpol(&g_tab);
    return 0;
  }
  g_err = "unknown PES " + s;
  return 1;
}
void orc_pes_set_dims(int ndim, int natom) { g_pes.ndim = ndim; g_pes.natom = natom; }
void orc_pes_set_V0(double v0) { g_pes.V0 = v0; }
void orc_pes_set_so2(double omegaforce, double r0) { g_pes.omegaforce = omegaforce; g_pes.r0 = r0; }
double orc_V(const double* x) { return g_pes.V(x); }
void orc_Vprime(double* x, double* grad) { g_pes.Vprime(x, grad); }
void orc_Vdoubleprime(double* x, double* hess) { g_pes.Vdoubleprime(x, hess); }
// batch: x(ndim,natom,nbatch); v/grad may be NULL; x is updated in place when grad is requested
// (the FD drift of mcmod_waterdimer_ccpol.f90:48-52 is kept, as step_v sees it)
int orc_pes_eval(long nbatch, double* x, double* v, double* grad) {
  int nd = g_pes.ndim * g_pes.natom;
  ORC_TRY(for (long b = 0; b < nbatch; ++b) {
    if (v) v[b] = g_pes.V(x + b * nd);
    if (grad) g_pes.Vprime(x + b * nd, grad + b * nd);
  })
}

// ---- module verletint / instantonmod -------------------------------------------------------
int orc_nm_setup(int n, int ndim, int natom, const double* mass, double betan, double tau, double gamma,
                 double dt, int cayley, int fixedends) {
  g_v = Verlet();
  g_v.setup(n, ndim, natom, mass, betan, &g_pes);
  g_v.tau = tau;
  g_v.gamma = gamma;
  g_v.dt = dt;
  g_v.cayley = cayley != 0;
  g_v.fixedends = fixedends != 0;
  return 0;
}
int orc_init_nm(const double* a, const double* b) { ORC_TRY(g_v.init_nm(a, b)) }
void orc_get_nm(double* T, double* lam, double* beadmass, double* beadvec) {
  if (T) std::memcpy(T, g_v.transmatrix.data(), g_v.transmatrix.size() * 8);
  if (lam) std::memcpy(lam, g_v.lam.data(), g_v.lam.size() * 8);
  if (beadmass) std::memcpy(beadmass, g_v.beadmass.data(), g_v.beadmass.size() * 8);
  if (beadvec) std::memcpy(beadvec, g_v.beadvec.data(), g_v.beadvec.size() * 8);
}
void orc_set_rng(unsigned long long seed, unsigned int gid) { g_v.seed = seed; g_v.traj_gid = gid; }
void orc_nm_forward(const double* x, double* q, int bead) { g_v.nm_forward(x, q, bead); }
void orc_nm_backward(double* q, double* x, int bead) { g_v.nm_backward(q, x, bead); }
void orc_step_nm(double time, double* x, double* p) { g_v.step_nm(time, x, p); }
void orc_step_v(double time, double* x, double* p) { g_v.step_v(time, x, p); }
void orc_step_langevin(double* p, unsigned long long step) { g_v.setup_pile(); g_v.step_langevin(p, step); }
void orc_sample_momenta(double* p, int stream, unsigned long long step) { g_v.sample_momenta(p, stream, step); }
int orc_init_path(double xi, const double* lampath, const double* path, const double* splinepath, int npath,
                  double* x, double* p) {
  ORC_TRY(g_v.init_path(xi, lampath, path, splinepath, npath, x, p))
}
// thermostat: 1 Andersen (propagate_pimd_nm), 2 Langevin (propagate_pimd_pile)
int orc_propagate(int thermostat, double* x, double* p, const double* dbdl, long NMC, long imin, long Noutput,
                  double* dHdr) {
  g_v.NMC = NMC;
  g_v.imin = imin;
  g_v.Noutput = Noutput;
  g_v.nan_trap = false;
  ORC_TRY(*dHdr = (thermostat == 1) ? g_v.propagate_pimd_nm(x, p, dbdl) : g_v.propagate_pimd_pile(x, p, dbdl);
          if (g_v.nan_trap) throw std::runtime_error("NaN in propagation"))
}
// dHdrlimit and the path init_path re-initialises from (verletmodule.f90:404-409); limit < 0 switches the guard off
void orc_set_dhdrlimit(double limit, double xi, const double* lampath, const double* path, const double* splinepath, int npath) {
  g_v.dHdrlimit = limit;
  g_v.rp_xi = xi;
  if (limit >= 0.0) {
    const size_t nd = (size_t)g_v.ndim * g_v.natom;
    g_v.rp_lam.assign(lampath, lampath + npath);
    g_v.rp_path.assign(path, path + npath * nd);
    g_v.rp_spl.assign(splinepath, splinepath + npath * nd);
  }
}
int orc_poisson(unsigned long long seed, unsigned long long step, unsigned int gid, double lambda) {
  return poisson_norm(seed, step, gid, lambda);
}
double orc_normal(unsigned long long seed, int stream, unsigned long long step, unsigned int gid,
                  unsigned long long idx) {
  return normal_at(seed, stream, step, gid, idx);
}
double orc_UM(const double* x, const double* a, const double* b) { return g_v.UM(x, a, b); }
void orc_UMhessian(double* x, int singlewell, double* answer) { g_v.UMhessian(x, singlewell != 0, answer); }
void orc_UMprime(const double* x, double* g, const double* a, const double* b) { g_v.UMprime(x, g, a, b); }
double orc_UMforceenergy(const double* x, double* g, const double* a, const double* b) {
  return g_v.UMforceenergy(x, g, a, b);
}
void orc_gauleg(double x1, double x2, double* x, double* w, int n) { gauleg(x1, x2, x, w, n); }
int orc_spline(const double* x, const double* y, int n, double yp1, double ypn, double* y2) {
  ORC_TRY(spline(x, y, n, yp1, ypn, y2))
}
double orc_splint(const double* xa, const double* ya, const double* y2a, int n, double x) {
  return splint(xa, ya, y2a, n, x);
}
// gammp of watermethane.f90:411-430 (Numerical Recipes, EPS = 3e-7), for the accuracy check against scipy
double orc_wm_gammp(double a, double x) { return WaterMethane::gammp(a, x); }
double orc_splin_grad(const double* xa, const double* ya, const double* y2a, int n, double x) {
  return splin_grad(xa, ya, y2a, n, x);
}

}  // extern "C"
