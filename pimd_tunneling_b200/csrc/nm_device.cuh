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

// splint / locate (instantonmod.f90:500-524, 560-596): the spline path at reaction coordinate xv for one dof
// (ya, y2: that dof's path(:) and splinepath(:)); used by init_path and by the dHdrlimit re-initialisation
__device__ __forceinline__ double splint_at(const double* __restrict__ lampath, const double* __restrict__ ya,
                                            const double* __restrict__ y2, int npath, double xv) {
  const bool ascnd = lampath[npath - 1] >= lampath[0];
  int jl = 0, ju = npath + 1;
  while (ju - jl > 1) {
    const int jm = (ju + jl) / 2;
    if (ascnd == (xv >= lampath[jm - 1])) jl = jm;
    else ju = jm;
  }
  int loc = jl;
  if (xv == lampath[0]) loc = 1;
  else if (xv == lampath[npath - 1]) loc = npath - 1;
  int klo = loc < npath - 1 ? loc : npath - 1;
  if (klo < 1) klo = 1;
  const int khi = klo + 1;
  const double h = lampath[khi - 1] - lampath[klo - 1];
  const double aa = (lampath[khi - 1] - xv) / h, bb = (xv - lampath[klo - 1]) / h;
  return aa * ya[klo - 1] + bb * ya[khi - 1] + ((aa * aa * aa - aa) * y2[klo - 1] + (bb * bb * bb - bb) * y2[khi - 1]) * (h * h) / 6.0;
}

// ---- elementwise normal-mode update -------------------------------------------------------
enum { OP_KICK = 1, OP_ROT1 = 2, OP_LANGEVIN = 4, OP_ROT2 = 8, OP_ANDERSEN = 16, OP_CLOCK = 32 };

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
