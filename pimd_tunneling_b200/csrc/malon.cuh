// Malonaldehyde surface (pes_malonaldehyde.f90, module malonaldehyde; plugin mcmod_malon.f90:1-73): a constant plus 9 Morse
// terms, 80 one-, 1728 two- and 1741 three-dimensional Gaussians over the 36 interatomic distances of the nine atoms
// (C C O C O H H H H), with an analytic Cartesian gradient (B matrix) and Hessian.  x is (3, 9) in bohr, energy in
// Hartree above equilibrium.  The fit's numbers are the DATA statements of pes_malonaldehyde.f90:45-9347, shipped as
// pimd_tunneling_b200/data/malonaldehyde.tbl (tools/pack_malon_tables.py).  Arithmetic in the reference's operation order
// (built with -fmad=false); dexp is the shared math policy's exp.
#pragma once
#include <cstdint>

#include "../../include/pimdk_detmath.h"

namespace pimdk {

constexpr int kMalAtoms = 9, kMalDist = 36, kMalDof = 27;
constexpr int kMalMorse = 9, kMalG1 = 80, kMalG2 = 1728, kMalG3 = 1741;
constexpr int kMalHint = (kMalDist + 1) * kMalDist / 2;   // packed lower triangle of the internal Hessian

// Device image: one record per term, 16-byte aligned so that a (warp-uniform) record is read with 16-byte loads; the distance
// indices are 0-based and packed into one word per term (byte j = index of the term's j-th distance).
struct MalonTab {
  double shift, pad_;
  alignas(16) double morse[4 * kMalMorse];   // re, alpha, de, 0
  alignas(16) double g1d[4 * kMalG1];        // x0, alpha, d, shift
  alignas(16) double g2d[6 * kMalG2];        // x0(2), alpha(2), d, shift
  alignas(16) double g3d[8 * kMalG3];        // x0(3), alpha(3), d, shift
  uint32_t imorse[kMalMorse + 3], ig1d[kMalG1], ig2d[kMalG2], ig3d[kMalG3 + 3];
};
static_assert(sizeof(MalonTab) % 16 == 0, "uploaded and read in 16-byte granules");

// Reads the packed table file; returns "" or a message (malon_tables.cpp)
const char* load_malon_tab(const char* dir, MalonTab* out);

// v_gauss / f_gauss / h_gauss share their first lines (pes_malonaldehyde.f90:9660-9665, :9686-9691, :9715-9720)
template <int ND>
PIMDK_HD double mal_gauss_arg(const double* r, const double* x0, const double* alpha) {
  double v = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) v = v + ((r[i] - x0[i]) * (r[i] - x0[i])) * alpha[i];
  return v * 0.5;
}

// index of distance (i, j), j < i, in the reference's loop order (:9356-9364), 0-based
PIMDK_HD int mal_pair(int i, int j) { return i * (i - 1) / 2 + j; }

}  // namespace pimdk
