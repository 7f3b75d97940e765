// 1D and 2D model surfaces: mcmod_1d.f90:8-58 (V, Vprime) and mcmod_2dtest.f90:11-61.
// One thread per geometry (bead); arithmetic in the reference's order (built with -fmad=false),
// exp from the shared math policy.  The reference's 2D Vprime evaluates 24 exponentials without
// CSE (mcmod_2dtest.f90:52-55); the values are identical, so each is computed once here.
#include "../../include/pimdk_detmath.h"
#include "kernels.h"

namespace pimdk {
namespace {

__global__ void __launch_bounds__(256)
simple_pes_kernel(int kind, SimplePesParams P, GeomLayout L, const double* __restrict__ x, double* __restrict__ v,
                  double* __restrict__ grad, long ngeom, int* __restrict__ flags) {
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < ngeom; g += (long)gridDim.x * blockDim.x) {
    const long base = L.base(g);
    if (kind == PES_1D) {
      double s = 0.0;
      for (int d = 0; d < P.ndof; ++d) {
        const double xi = x[base + d * L.stride_dof];
        const double r = xi / P.x0;
        const double u = r * r - 1.0;
        if (v) s += P.Vheight * (u * u);
        if (grad) {
          const double gv = u * 4.0 * P.Vheight * xi / (P.x0 * P.x0);
          grad[base + d * L.stride_dof] = gv;
          if (gv != gv) atomicOr(flags, PIMDK_FLAG_NAN);
        }
      }
      if (v) v[g] = s;
    } else {
      const double x1 = x[base], x2 = x[base + L.stride_dof];
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
      if (v) v[g] = answer - P.V0;
      if (grad) {
        grad[base] = g1;
        grad[base + L.stride_dof] = g2;
        if (g1 != g1 || g2 != g2) atomicOr(flags, PIMDK_FLAG_NAN);
      }
    }
  }
}

}  // namespace

cudaError_t launch_simple_pes(PesKind kind, const SimplePesParams& P, GeomLayout L, const double* x, double* v,
                              double* grad, long ngeom, int* flags, cudaStream_t st) {
  long blocks = (ngeom + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  if (blocks < 1) blocks = 1;
  simple_pes_kernel<<<(unsigned)blocks, 256, 0, st>>>((int)kind, P, L, x, v, grad, ngeom, flags);
  return cudaGetLastError();
}

}  // namespace pimdk
