// Device image of the CCpol-8sf parameter tables (what the reference keeps in COMMON /ddaattaa/,
// main_CCpol-8sf.f:6-11), trimmed to what the hot path reads and laid out for shared-memory
// staging: every access in the kernels is warp-uniform, so a table read is one broadcast LDS.
#pragma once
#include <cstddef>
#include <cstdint>

namespace pimdk {

constexpr int kNType = 5;      // SAPT-5s'f site types actually used (1..5; type 6 never occurs)
constexpr int kNParab = 84;
constexpr int kNParam = 18;

struct CcpolDev {
  // ---- CCpol-8s rigid model (the only part the rigid stage stages into shared memory) ----
  double cc[144];                            // CCpol-8s linear coefficients
  double params[134];                        // CCpol-8s nonlinear parameters (1-based in the file)
  double sites[75];                          // sites(3,25) body-frame coordinates
  double chrg[5];                            // induction charges of sites 1..5
  double V0;                                 // module variable V0 of mcmod_mass
  uint8_t ind_beta[625];                     // ind_beta(nsA,nsB) -> [nsB*25 + nsA] (0-based)
  uint8_t ind_charge[5];
  uint8_t ind_d1[25];                        // 5x5, [b*5+a]
  uint8_t ind_d6[9], ind_d8[9], ind_d10[9], ind_c6[9], ind_c8[9], ind_c10[9];  // 3x3, [b*3+a]
  // CCpol-8s site classes: runs of consecutive sites with identical ind_beta rows (8 classes of
  // sizes 1,2,2,4,4,4,4,4 for data_ccdata); cls_start[c]..cls_start[c+1]-1 are the sites of class c
  uint8_t cls_start[26];
  uint8_t iembed;                            // 1 Eckart, 2 Radau f=1 (main_CCpol-8sf.f:28-107)
  uint8_t icc;                               // 1: Erigid + (val - vall); 0: SAPT-5s'f alone
  uint8_t pad0_[3];
  int32_t ncls;
  int32_t iemonomer;
  double bin_beta[37];                       // beta of each bin (params(ind_beta) of its site pairs); [36] unused
  // The 36 bins as tasks of the item-per-lane sweep (a warp walks one bin for 32 energies), largest first:
  //   bits 0..4 first site of class ca, 5..7 its size, 8..12 first site of class cb, 13..15 its size, 16..21 bin
  uint32_t tbins[36];
  // ---- SAPT-5s'f flexible model ----
  alignas(16) double param[kNParam * kNType];  // param(k,t) -> [(t-1)*18 + k-1]; 16-byte aligned: staging granule
  double parab[kNParab * kNType * kNType];   // parab(k,ta,tb)     -> [((tb-1)*5+(ta-1))*84 + k-1]
  alignas(16) double c[568];                 // SAPT-5s'f linear coefficients (read as 16-byte pairs)
  // static image of poten's first-encounter index map itypus (proc_sapt5sf_new_ncd.f:181-203):
  // first linear coefficient (1-based) of the symmetric / antisymmetric block of a type pair,
  // 0 when the pair type carries no exponential.
  int16_t itu_s[kNType * kNType];
  int16_t itu_a[kNType * kNType];
  // static per type-pair facts (bit 0: carries exponential/linear terms, bit 1: damped electrostatics
  // (dmp1 != 0), bit 2: damped dispersion (any of dmp6/8/10 != 0)).  A term whose damping parameter is
  // exactly zero evaluates to +-0 in the reference (function d returns 0 for br == 0,
  // proc_sapt5sf_new_ncd.f:1234-1237) and adding +-0 changes no bits, so the kernels skip it.
  uint8_t pairflags[kNType * kNType];
  uint8_t potparts_old;                      // ipotparts = 0: potparts_old (surfaces 8, 9)
  uint8_t pad1_[2];
};
// The tables as kernel parameters.  Every table read of the pair-sum, rigid and sweep kernels is warp-uniform; taken from
// the constant bank it is an operand fetched through the uniform datapath (LDCU / c[0][..]) instead of a shared-memory load
// with a vector-register address: no staging prologue, no barrier, fewer vector registers (the pair-sum kernel loses its
// spills) and an idle load/store pipe.  Members carry CcpolDev's names so that the device functions below take either.
struct SaptParams {      // SAPT-5s'f flexible model (22.3 KB of the 32 KB a kernel may take as parameters)
  double param[kNParam * kNType];
  double parab[kNParab * kNType * kNType];
  alignas(16) double c[568];
  int16_t itu_s[kNType * kNType], itu_a[kNType * kNType];
  uint8_t pairflags[kNType * kNType];
};
struct RigidParams {     // CCpol-8s rigid model (3.6 KB)
  double cc[144];
  double params[134];
  double sites[75];
  double chrg[5];
  double bin_beta[37];
  uint32_t tbins[36];
  uint8_t ind_charge[5];
  uint8_t ind_d1[25];
  uint8_t ind_d6[9], ind_d8[9], ind_d10[9], ind_c6[9], ind_c8[9], ind_c10[9];
};
template <class A, class B, int N>
inline void copy_members(A (&dst)[N], const B (&src)[N]) {
  for (int i = 0; i < N; ++i) dst[i] = src[i];
}
inline void fill_params(const CcpolDev& h, SaptParams* s, RigidParams* r) {
  copy_members(s->param, h.param);
  copy_members(s->parab, h.parab);
  copy_members(s->c, h.c);
  copy_members(s->itu_s, h.itu_s);
  copy_members(s->itu_a, h.itu_a);
  copy_members(s->pairflags, h.pairflags);
  copy_members(r->cc, h.cc);
  copy_members(r->params, h.params);
  copy_members(r->sites, h.sites);
  copy_members(r->chrg, h.chrg);
  copy_members(r->bin_beta, h.bin_beta);
  copy_members(r->tbins, h.tbins);
  copy_members(r->ind_charge, h.ind_charge);
  copy_members(r->ind_d1, h.ind_d1);
  copy_members(r->ind_d6, h.ind_d6);
  copy_members(r->ind_d8, h.ind_d8);
  copy_members(r->ind_d10, h.ind_d10);
  copy_members(r->ind_c6, h.ind_c6);
  copy_members(r->ind_c8, h.ind_c8);
  copy_members(r->ind_c10, h.ind_c10);
}

// bytes of the leading rigid-model block = offset of the first SAPT member (a multiple of 16)
#define PIMDK_RIGID_TABLE_BYTES (offsetof(::pimdk::CcpolDev, param))

// Host-side full tables (same content as the reference's COMMON block) and loaders.
struct CcpolHost {
  double param[18 * 6];
  double parab[84 * 6 * 6];
  double c[1000];
  int numlin;
  double cc[2000];
  int nlin0;
  double params[1000];
  int nparsall;
  double chrg[25];
  double sites[75];
  int ind_charge[25];
  int ind_beta[625], ind_d1[625], ind_d6[625], ind_d8[625], ind_d10[625], ind_c6[625], ind_c8[625], ind_c10[625];
  int isurf, iembed, ipotparts, icc;   // surface switches of init_ccpol (main_CCpol-8sf.f:28-107)
};

// Loads the tables of surface `isurf` (1..10, main_CCpol-8sf.f:14-107; the plugin uses 3) from `dir`: first the
// reference's own text files (data_SAPT5spf*_20*, data_CCpol8s, data_ccdata — the names the reference opens from its
// CWD), else the packed *.tbl files shipped in pimd_tunneling_b200/data.  Returns "" on success, else a message.
const char* load_ccpol_tables(const char* dir, int isurf, CcpolHost* out);
// Builds the device image (static index map, byte-sized index tables).  Returns error or "".
const char* build_ccpol_dev(const CcpolHost& h, int iemonomer, CcpolDev* out);

}  // namespace pimdk
