// Ring-polymer (open chain) potential and gradient: UM, UMprime, UMforceenergy
// (instantonmod.f90:17-151).  The per-bead PES values/gradients come from the PES kernels; here the
// spring terms are added.  The scalar UM is accumulated by ONE thread in exactly the reference's
// order (bead energy, then that bead's springs, ..., then the fixed-end springs) so that the value
// handed to L-BFGS-B is reproducible to the last bit (its terms are formed in parallel first); the
// gradient is elementwise.
#include "kernels.h"

namespace pimdk {
namespace {

__global__ void __launch_bounds__(256)
um_grad_kernel(int n, int ndim, int natom, const double* __restrict__ x, const double* __restrict__ a,
               const double* __restrict__ b, const double* __restrict__ mass, double betan, int fixedends,
               const double* __restrict__ gbead, double* __restrict__ grad) {
  const long total = (long)n * ndim * natom;
  // polymer blockIdx.y of a batch: its own x, gradient and end point b; a and the masses are shared
  x += blockIdx.y * total;
  gbead += blockIdx.y * total;
  grad += blockIdx.y * total;
  if (b) b += (long)blockIdx.y * ndim * natom;
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

// One CTA.  The terms of UM (bead energy, that bead's ndof spring terms) are independent products: the CTA forms
// them for a tile of beads in parallel into shared memory, then ONE thread adds them in the reference's order
// (instantonmod.f90:24-33) — the additions are a serial dependency chain either way (8 cycles per DADD), but this
// way no global-memory latency sits inside it.
constexpr int kUmTile = 128;
__global__ void __launch_bounds__(256)
um_energy_kernel(int n, int ndim, int natom, const double* __restrict__ x, const double* __restrict__ a,
                 const double* __restrict__ b, const double* __restrict__ mass, double betan, int fixedends,
                 const double* __restrict__ vbead, double* __restrict__ um_out) {
  extern __shared__ double terms[];   // [kUmTile][1 + ndof]
  const int ndof = ndim * natom, w = 1 + ndof;
  x += (long)blockIdx.y * n * ndof;   // polymer blockIdx.y of a batch
  vbead += (long)blockIdx.y * n;
  um_out += blockIdx.y;
  if (b) b += (long)blockIdx.y * ndof;
  const double bn2 = betan * betan;
  double um = 0.0;
  for (int i0 = 0; i0 < n; i0 += kUmTile) {
    const int nb = min(kUmTile, n - i0);
    for (int t = threadIdx.x; t < nb * w; t += blockDim.x) {
      const int ii = t / w, c = t - ii * w, i = i0 + ii;
      double v;
      if (c == 0) {
        v = vbead[i];
      } else if (i < n - 1) {
        // reference loop order: j = dim outer, k = atom inner (:27-31); c - 1 = j * natom + k
        const int j = (c - 1) / natom, k = (c - 1) - j * natom;
        const long e = (long)(k * ndim + j) * n + i;
        const double d = x[e + 1] - x[e];
        v = (0.5 * mass[k] / bn2) * (d * d);
      } else {
        v = 0.0;   // the last bead has no spring to a successor; its slots are not added
      }
      terms[t] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // terms[] is already in summation order (bead energy, its springs, next bead ...); the last bead of the
      // polymer has no spring slots to add.  Eight loads are issued ahead of the eight dependent additions.
      const int cnt = nb * w - ((i0 + nb == n) ? ndof : 0);
      int t = 0;
      for (; t + 8 <= cnt; t += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = terms[t + q];
#pragma unroll
        for (int q = 0; q < 8; ++q) um = um + v[q];
      }
      for (; t < cnt; ++t) um = um + terms[t];
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
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

// npoly ring polymers x(n,ndim,natom,npoly) with their own end points b(ndim,natom,npoly); a shared
cudaError_t launch_um(int npoly, int n, int ndim, int natom, const double* x, const double* a, const double* b,
                      const double* mass, double betan, int fixedends, const double* vbead, const double* gbead,
                      double* um_out, double* grad_out, cudaStream_t st) {
  if (npoly < 1 || npoly > 65535) return cudaErrorInvalidValue;
  if (grad_out) {
    long total = (long)n * ndim * natom;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    um_grad_kernel<<<dim3((unsigned)blocks, (unsigned)npoly), 256, 0, st>>>(n, ndim, natom, x, a, b, mass, betan, fixedends, gbead,
                                                                           grad_out);
  }
  if (um_out)
    um_energy_kernel<<<dim3(1, (unsigned)npoly), 256, (size_t)kUmTile * (1 + ndim * natom) * sizeof(double), st>>>(
        n, ndim, natom, x, a, b, mass, betan, fixedends, vbead, um_out);
  return cudaGetLastError();
}

}  // namespace pimdk
