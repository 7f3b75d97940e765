// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// CPU restatement of the reference's CCpol-8sf parameter tables (COMMON /ddaattaa/,
// main_CCpol-8sf.f:6-11) and of the three parsers that fill it:
//   data1            proc_sapt5sf_new_ncd.f:1266-1356   (SAPT-5s'f file, unit 55)
//   ccpol8s_dimer(-1) proc_ccpol8s-dimer_xyz_ncd.f:40-58 (data_CCpol8s, unit 7)
//   read_cc_data     main_CCpol-8sf.f:822-923            (data_ccdata, unit 8)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.
#pragma once
#include <string>

namespace oracle {

struct CcpolTables {
  // Fortran column-major images, 1-based accessors below.
  double param[18 * 6];        // param(18,6)
  double parab[84 * 6 * 6];    // parab(84,6,6)
  double c[1000];              // SAPT-5s'f linear coefficients
  int    numlin;
  double cc[2000];             // CCpol-8s linear coefficients
  int    nlin0;
  double params[1000];         // CCpol-8s nonlinear parameters
  int    nparsall;
  double chrg[25];
  double sites[3 * 25];        // sites(3,25)
  int ind_charge[25];
  int ind_beta[25 * 25], ind_d1[25 * 25], ind_d6[25 * 25], ind_d8[25 * 25], ind_d10[25 * 25];
  int ind_c6[25 * 25], ind_c8[25 * 25], ind_c10[25 * 25];
  // surface switches set by init_ccpol (main_CCpol-8sf.f:28-107)
  int iembed, ipotparts, icc, iemonomer;
  // 1 (default): PJT2 literals are FP64 as under the reference makefile's -r8; 0: single-precision
  // literals, the build the golden valm(1:10) came from (used only to pin the oracle)
  int pjt2_r8;

  double& PARAM(int k, int t) { return param[(t - 1) * 18 + (k - 1)]; }
  double  PARAM(int k, int t) const { return param[(t - 1) * 18 + (k - 1)]; }
  double& PARAB(int k, int t1, int t2) { return parab[((t2 - 1) * 6 + (t1 - 1)) * 84 + (k - 1)]; }
  double  PARAB(int k, int t1, int t2) const { return parab[((t2 - 1) * 6 + (t1 - 1)) * 84 + (k - 1)]; }
  double  SITES(int j, int i) const { return sites[(i - 1) * 3 + (j - 1)]; }
  static int IJ(int i, int j) { return (j - 1) * 25 + (i - 1); }  // ind_x(i,j), column-major 25x25
};

// name of the SAPT data file selected by isurf (main_CCpol-8sf.f:28-107); also sets
// iembed/ipotparts/icc.  Returns false for isurf outside 1..10.
bool surface_switches(int isurf, CcpolTables& t, std::string& saptfile);

// Parse the reference's own text files found in `dir` (bare names, as the reference
// opens them from the CWD).  Throws std::runtime_error on malformed input (the
// reference's `stop 010..040`).
void load_text(const std::string& dir, int isurf, int iemon, CcpolTables& t);

// Packed form written by tools/pack_ccpol_tables.py ("key count\n values...").
void load_packed(const std::string& sapt_tbl, const std::string& ccpol8s_tbl, int isurf, int iemon,
                 CcpolTables& t);

}  // namespace oracle
