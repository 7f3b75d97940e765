// Ring-polymer (open chain) potential and gradient: UM, UMprime, UMforceenergy
// (instantonmod.f90:17-151).  The per-bead PES values/gradients come from the PES kernels; here the
// spring terms are added.  The scalar UM is accumulated by ONE thread in exactly the reference's
// order (bead energy, then that bead's springs, ..., then the fixed-end springs) so that the value
// handed to L-BFGS-B is reproducible to the last bit; the gradient is elementwise.
#include "kernels.h"

namespace pimdk {
namespace {

__global__ void __launch_bounds__(256)
um_grad_kernel(int n, int ndim, int natom, const double* __restrict__ x, const double* __restrict__ a,
               const double* __restrict__ b, const double* __restrict__ mass, double betan, int fixedends,
               const double* __restrict__ gbead, double* __restrict__ grad) {
  const long total = (long)n * ndim * natom;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % n);         // bead (0-based)
    const int dof = (int)(e / n);       // atom*ndim + dim
    const double m = mass[dof / ndim];
    const double bn2 = betan * betan;
    const double xi = x[e];
    double s;
    if (i == 0) {
      s = fixedends ? m * (2.0 * xi - a[dof] - x[e + 1]) / bn2 : m * (xi - x[e + 1]) / bn2;
    } else if (i == n - 1) {
      s = fixedends ? m * (2.0 * xi - x[e - 1] - b[dof]) / bn2 : m * (xi - x[e - 1]) / bn2;
    } else {
      s = m * (2.0 * xi - x[e - 1] - x[e + 1]) / bn2;
    }
    grad[e] = s + gbead[e];
  }
}

__global__ void um_energy_kernel(int n, int ndim, int natom, const double* __restrict__ x,
                                 const double* __restrict__ a, const double* __restrict__ b,
                                 const double* __restrict__ mass, double betan, int fixedends,
                                 const double* __restrict__ vbead, double* __restrict__ um_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double um = 0.0;
  const double bn2 = betan * betan;
  for (int i = 0; i < n; ++i) {
    um = um + vbead[i];
    if (i < n - 1)
      for (int j = 0; j < ndim; ++j)
        for (int k = 0; k < natom; ++k) {
          const long e = (long)(k * ndim + j) * n + i;
          const double d = x[e + 1] - x[e];
          um = um + (0.5 * mass[k] / bn2) * (d * d);
        }
  }
  if (fixedends)
    for (int j = 0; j < ndim; ++j)
      for (int k = 0; k < natom; ++k) {
        const int dof = k * ndim + j;
        const double d1 = x[(long)dof * n] - a[dof];
        um = um + (0.5 * mass[k] / bn2) * (d1 * d1);
        const double d2 = b[dof] - x[(long)dof * n + n - 1];
        um = um + (0.5 * mass[k] / bn2) * (d2 * d2);
      }
  *um_out = um;
}

}  // namespace

cudaError_t launch_um(int n, int ndim, int natom, const double* x, const double* a, const double* b,
                      const double* mass, double betan, int fixedends, const double* vbead, const double* gbead,
                      double* um_out, double* grad_out, cudaStream_t st) {
  if (grad_out) {
    long total = (long)n * ndim * natom;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    um_grad_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, ndim, natom, x, a, b, mass, betan, fixedends, gbead, grad_out);
  }
  if (um_out) um_energy_kernel<<<1, 32, 0, st>>>(n, ndim, natom, x, a, b, mass, betan, fixedends, vbead, um_out);
  return cudaGetLastError();
}

}  // namespace pimdk
