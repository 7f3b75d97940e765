// 1D and 2D model surfaces: mcmod_1d.f90:8-58 (V, Vprime) and mcmod_2dtest.f90:11-61.
// One thread per geometry (bead); arithmetic in the reference's order (built with -fmad=false),
// exp from the shared math policy (pes_simple_device.cuh).
#include "kernels.h"
#include "pes_simple_device.cuh"

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
        double xi = x[base + d * L.stride_dof], e, gv;
        SimplePesParams P1 = P;
        P1.ndof = 1;
        simple_pes_eval<1>(kind, P1, &xi, &e, &gv, v != nullptr, grad != nullptr);
        if (v) s += e;
        if (grad) {
          grad[base + d * L.stride_dof] = gv;
          if (gv != gv) atomicOr(flags, PIMDK_FLAG_NAN);
        }
      }
      if (v) v[g] = s;
    } else {
      double xx[2] = {x[base], x[base + L.stride_dof]}, e, gg[2];
      simple_pes_eval<2>(kind, P, xx, &e, gg, v != nullptr, grad != nullptr);
      if (v) v[g] = e;
      if (grad) {
        grad[base] = gg[0];
        grad[base + L.stride_dof] = gg[1];
        if (gg[0] != gg[0] || gg[1] != gg[1]) atomicOr(flags, PIMDK_FLAG_NAN);
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
