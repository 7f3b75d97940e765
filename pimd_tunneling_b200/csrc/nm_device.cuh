// Device pieces shared by the streamed kernels (nm_kernels.cu) and the fused warp-per-ring-polymer kernel
// (fused_small.cu): the same expressions in both, so the two paths agree bit for bit.
#pragma once
#include "kernels.h"
#include "philox.cuh"

namespace pimdk {

__device__ __forceinline__ double beadvec_at(const NmTables& nm, const double* __restrict__ a,
                                             const double* __restrict__ b, long traj, int dof, int k) {
  // beadvec(k,dof) = [a sin(k pi/(n+1)) + b sin(n k pi/(n+1))] * sqrt(2/(n+1)) / (lam_k betan)**2
  double v = a[dof] * nm.sA[k] + b[traj * nm.ndof + dof] * nm.sB[k];
  v = v * nm.norm;
  return v / nm.lamb2[k];
}

// ---- elementwise normal-mode update -------------------------------------------------------
enum { OP_KICK = 1, OP_ROT1 = 2, OP_LANGEVIN = 4, OP_ROT2 = 8 };

__device__ __forceinline__ void rotate(const NmTables& nm, int ak, double& P, double& Q) {
  const double bm = nm.bmass[ak];
  if (nm.cayley) {
    const double om = nm.omega[ak], time = nm.time;
    const double w2t2 = (om * om) * (time * time);
    double newpi = P * (4.0 - w2t2) - 4.0 * Q * bm * (om * om) * time;
    newpi = newpi / (4.0 + w2t2);
    double q = Q * (4.0 - w2t2) + 4.0 * P * time / bm;
    q = q / (4.0 + w2t2);
    P = newpi;
    Q = q;
  } else {
    const double cw = nm.cosw[ak], sw = nm.sinw[ak];
    const double newpi = P * cw - Q * nm.omega[ak] * bm * sw;
    Q = Q * cw + P * sw / nm.wbm[ak];
    P = newpi;
  }
}


}  // namespace pimdk
