// Bead energy and gradient of the model surfaces: mcmod_1d.f90:20,31-32, mcmod_2dtest.f90:33-39,48-58 and
// mcmod_so2.f90:17-48 (harmonic ring), in the reference's operation order (compile with -fmad=false).  Shared by the
// streamed kernel (pes_simple.cu) and the fused warp-per-ring-polymer kernel (fused_small.cu).
#pragma once
#include "../../include/pimdk_detmath.h"
#include "kernels.h"

namespace pimdk {

// x[ndof] -> v (if want_v), g[ndof] (if want_g).  The reference's 2D Vprime evaluates 24 exponentials without
// CSE (mcmod_2dtest.f90:52-55); the values are identical, so each is computed once here.
template <int MAXDOF>
__device__ __forceinline__ void simple_pes_eval(int kind, const SimplePesParams& P, const double* x, double* v,
                                                double* g, bool want_v, bool want_g) {
  if (kind == PES_1D) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < MAXDOF; ++d) {
      if (d < P.ndof) {
        const double xi = x[d];
        const double r = xi / P.x0;
        const double u = r * r - 1.0;
        if (want_v) s += P.Vheight * (u * u);
        if (want_g) g[d] = u * 4.0 * P.Vheight * xi / (P.x0 * P.x0);
      }
    }
    if (want_v) *v = s;
  } else if (kind == PES_SO2) {
    // mcmod_so2.f90:22-23: r = sqrt(x**2 + y**2); V = 0.5*omegaforce**2*(r-r0)**2 - V0
    // :42-44: grad = omegaforce**2 * x * (1 - r0/r)
    const double x1 = x[0], x2 = x[MAXDOF > 1 ? 1 : 0];
    const double r = sqrt(x1 * x1 + x2 * x2);
    const double w2 = P.omegaforce * P.omegaforce;
    if (want_v) *v = 0.5 * w2 * ((r - P.r0) * (r - P.r0)) - P.V0;
    if (want_g) {
      g[0] = w2 * x1 * (1.0 - P.r0 / r);
      if (MAXDOF > 1) g[1] = w2 * x2 * (1.0 - P.r0 / r);
    }
  } else {
    const double x1 = x[0], x2 = x[MAXDOF > 1 ? 1 : 0];
    double answer = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double dx = x1 - P.wx[k], dy = x2 - P.wy[k];
      const double u = dx * dx + dy * dy;
      const double ea = pimdk_exp(-P.a0 * u), eb = pimdk_exp(-P.b0 * u);
      answer = answer - 0.5 * ea;
      answer = answer - 0.5 * eb;
      g1 = g1 + P.a0 * dx * ea;
      g1 = g1 + P.b0 * dx * eb;
      g2 = g2 + P.a0 * dy * ea;
      g2 = g2 + P.b0 * dy * eb;
    }
    if (want_v) *v = answer - P.V0;
    if (want_g) {
      g[0] = g1;
      if (MAXDOF > 1) g[1] = g2;
    }
  }
}

}  // namespace pimdk
