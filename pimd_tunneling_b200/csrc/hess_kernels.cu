// Second derivatives for the instanton fluctuation factor:
//   Vdoubleprime of the three in-scope PES plugins   mcmod_1d.f90:37-57, mcmod_2dtest.f90:63-86,
//                                                    mcmod_waterdimer_ccpol.f90:59-76
//   UMhessian                                        instantonmod.f90:155-217 (LAPACK lower band storage)
// Hessians are hess(ndim,natom,ndim,natom,ngeom), column-major: with dof = atom*ndim + dim,
// hess[g*ndof*ndof + dof2*ndof + dof1] = d grad(dof2) / d x(dof1).  Built with -fmad=false: the central
// differences reproduce the reference's operation order, including the in-place perturbation drift of x.
#include "kernels.h"
#include "philox.cuh"
#include "pes_simple_device.cuh"

namespace pimdk {
namespace {

constexpr int kMaxSimpleDof = 4;

// 1D surface: central difference (eps = 1e-4) of the analytic gradient; x(i,j) is perturbed in place
// (x + eps, - 2 eps, + eps) in the loop order i = dim outer, j = atom inner, and keeps the round-off drift.
// 2D surface: the reference's "analytic" expressions are ASSIGNED inside its loop over the six wells, so only the
// last well contributes (and the mixed terms are as written there); restated literally — this is the matrix the
// reference's detJ diagonalises.
__global__ void __launch_bounds__(128)
simple_hessian_kernel(int kind, SimplePesParams P, int ndim, int natom, GeomLayout L, double* __restrict__ x,
                      double* __restrict__ hess, long ngeom) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngeom) return;
  const long base = L.base(g);
  const int nd = ndim * natom;
  double* H = hess + g * nd * nd;
  if (kind == PES_2DTEST) {
    const double x1 = x[base], x2 = x[base + L.stride_dof];
    double h11 = 0.0, h21 = 0.0, h12 = 0.0, h22 = 0.0;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
      const double u = (x1 - P.wx[k]) * (x1 - P.wx[k]) + (x2 - P.wy[k]) * (x2 - P.wy[k]);
      const double ea = pimdk_exp(-P.a0 * u), eb = pimdk_exp(-P.b0 * u);
      const double dvdu = P.a0 * ea + P.b0 * eb;
      const double d2vdu2 = -(P.a0 * P.a0) * ea - (P.b0 * P.b0) * eb;
      const double dudx = x1 - P.wx[k];
      const double dudy = x2 - P.wy[k];
      h11 = (d2vdu2 * dudx + dvdu) * dudx;
      h21 = (d2vdu2 * dudy + dvdu) * dudx;
      h12 = (d2vdu2 * dudy + dvdu) * dudx;
      h22 = (d2vdu2 * dudy + dvdu) * dudy;
    }
    H[0] = h11;  // hess(1,1,1,1)
    H[1] = h21;  // hess(2,1,1,1)
    H[2] = h12;  // hess(1,1,2,1)
    H[3] = h22;  // hess(2,1,2,1)
    return;
  }
  if (kind == PES_SO2) {   // mcmod_so2.f90:76-82, as written there (x(i)*x(j)*omegaforce**2*r0/r**3: no diagonal (1 - r0/r) term)
    const double x1 = x[base], x2 = x[base + L.stride_dof];
    const double r = sqrt(x1 * x1 + x2 * x2);
    const double w2 = P.omegaforce * P.omegaforce, r3 = r * r * r;
    H[0] = x1 * x1 * w2 * P.r0 / r3;
    H[1] = x1 * x2 * w2 * P.r0 / r3;
    H[2] = x1 * x2 * w2 * P.r0 / r3;
    H[3] = x2 * x2 * w2 * P.r0 / r3;
    return;
  }
  const double eps = 1e-4;
  double xx[kMaxSimpleDof], gp[kMaxSimpleDof], gm[kMaxSimpleDof], e;
  for (int d = 0; d < nd; ++d) xx[d] = x[base + d * L.stride_dof];
  SimplePesParams P1 = P;
  P1.ndof = nd;
  for (int i = 0; i < ndim; ++i)
    for (int j = 0; j < natom; ++j) {
      const int d1 = j * ndim + i;
      xx[d1] = xx[d1] + eps;
      simple_pes_eval<kMaxSimpleDof>(kind, P1, xx, &e, gp, false, true);
      xx[d1] = xx[d1] - 2.0 * eps;
      simple_pes_eval<kMaxSimpleDof>(kind, P1, xx, &e, gm, false, true);
      xx[d1] = xx[d1] + eps;
      for (int d2 = 0; d2 < nd; ++d2) H[d2 * nd + d1] = (gp[d2] - gm[d2]) / (2.0 * eps);
    }
  for (int d = 0; d < nd; ++d) x[base + d * L.stride_dof] = xx[d];
}

__global__ void perturb_kernel(GeomLayout L, double* __restrict__ x, long ngeom, int dof, double delta) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngeom) return;
  const long a = L.base(g) + dof * L.stride_dof;
  x[a] = x[a] + delta;
}

// hess(dof1, :) = (gradplus - gradminus) / (2 eps)   (mcmod_waterdimer_ccpol.f90:71)
__global__ void hess_column_kernel(GeomLayout L, const double* __restrict__ gp, const double* __restrict__ gm, long ngeom,
                                   int nd, int dof1, double eps, double* __restrict__ hess) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ngeom * nd) return;
  const long g = t / nd;
  const int d2 = (int)(t - g * nd);
  const long a = L.base(g) + d2 * L.stride_dof;
  hess[g * nd * nd + d2 * nd + dof1] = (gp[a] - gm[a]) / (2.0 * eps);
}

// UMhessian: one thread per (bead, idof1, idof2 >= idof1); band(ndof+1, totdof) zero-filled beforehand
__global__ void um_band_kernel(int n, int ndim, int natom, const double* __restrict__ hess, const double* __restrict__ mass,
                               double betan, int singlewell, double* __restrict__ band) {
  const int nd = ndim * natom;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n * nd * nd) return;
  const int i = (int)(t / (nd * nd));            // bead, 0-based
  const int r = (int)(t - (long)i * nd * nd);
  const int id1 = r / nd, id2 = r - id1 * nd;     // 0-based dof labels of the bead
  if (id2 < id1) return;
  const long fd1 = (long)nd * i + id1;
  const double* H = hess + (singlewell ? 0 : (long)i * nd * nd);
  // hess(j2,k2,j1,k1): perturbed coordinate (j2,k2) = id2, gradient component (j1,k1) = id1
  const double val = H[id1 * nd + id2] / sqrt(mass[id1 / ndim] * mass[id2 / ndim]);
  const long ld = nd + 1;
  if (id1 == id2) {
    band[0 + ld * fd1] = 2.0 / (betan * betan) + val;
    if (i > 0) band[nd + ld * fd1] = -1.0 / (betan * betan);
  } else {
    band[(id2 - id1) + ld * fd1] = val;
  }
}

// dense lower triangle (column-major, lda = N) from LAPACK lower band storage AB(1+r-c, c) = A(r, c), c <= r <= c+kd
__global__ void band_to_dense_kernel(long N, int kd, const double* __restrict__ band, double* __restrict__ A) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * N) return;
  const long c = t / N, r = t - c * N;
  const long lo = r < c ? r : c, hi = r < c ? c : r;   // symmetric fill
  A[t] = (hi - lo <= kd) ? band[(hi - lo) + (long)(kd + 1) * lo] : 0.0;
}

// init_path's readhess branch (verletmodule.f90:49-88): totdof normals N(0, sqrt(1/beta)) ...
__global__ void readhess_normals_kernel(long N, double stdev, uint64_t seed, uint32_t gid, double* __restrict__ tempx) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) tempx[i] = stdev * normal_at(seed, STREAM_READHESS, 0, gid, (uint64_t)i);
}
// ... and the displacement along the eigenvectors of the ring-polymer Hessian (:75-87), literally: modes 2..totdof
// with a non-negative eigenvalue, added in that order;  x(i2,j2,k2) += sqrt(1/(eta2(m) mass(k2))) tempx(m) Z(idof2, m)
// with the reference's row index idof2 = natom*(j2-1 + ndim*(i2-1)) + k2 (atom fastest, whereas UMhessian numbers the
// degrees of freedom dimension fastest: the two agree for natom = 1 or ndim = 1 only — restated, not repaired).
__global__ void readhess_displace_kernel(int n, int ndim, int natom, const double* __restrict__ eta,
                                         const double* __restrict__ Z, const double* __restrict__ mass,
                                         const double* __restrict__ tempx, double* __restrict__ x) {
  const long N = (long)n * ndim * natom;
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;   // x(n,ndim,natom), bead fastest
  if (e >= N) return;
  const int i2 = (int)(e % n), j2 = (int)((e / n) % ndim), k2 = (int)(e / ((long)n * ndim));
  const long idof2 = (long)natom * (j2 + (long)ndim * i2) + k2;
  double acc = x[e];
  for (long m = 1; m < N; ++m) {
    const double et = eta[m];
    if (et < 0.0) continue;
    acc = acc + sqrt(1.0 / (et * mass[k2])) * tempx[m] * Z[idof2 + N * m];
  }
  x[e] = acc;
}

}  // namespace

cudaError_t launch_readhess_displace(int n, int ndim, int natom, const double* eta, const double* Z, const double* mass,
                                     double stdev, uint64_t seed, uint32_t gid, double* tempx, double* x, cudaStream_t st) {
  const long N = (long)n * ndim * natom;
  readhess_normals_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, stdev, seed, gid, tempx);
  readhess_displace_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(n, ndim, natom, eta, Z, mass, tempx, x);
  return cudaGetLastError();
}

cudaError_t launch_simple_hessian(PesKind kind, const SimplePesParams& P, int ndim, int natom, GeomLayout L, double* x,
                                  double* hess, long ngeom, cudaStream_t st) {
  if (ndim * natom > kMaxSimpleDof) return cudaErrorInvalidValue;
  simple_hessian_kernel<<<(unsigned)((ngeom + 127) / 128), 128, 0, st>>>((int)kind, P, ndim, natom, L, x, hess, ngeom);
  return cudaGetLastError();
}
cudaError_t launch_perturb(GeomLayout L, double* x, long ngeom, int dof, double delta, cudaStream_t st) {
  perturb_kernel<<<(unsigned)((ngeom + 255) / 256), 256, 0, st>>>(L, x, ngeom, dof, delta);
  return cudaGetLastError();
}
cudaError_t launch_hess_column(GeomLayout L, const double* gp, const double* gm, long ngeom, int nd, int dof1, double eps,
                               double* hess, cudaStream_t st) {
  hess_column_kernel<<<(unsigned)((ngeom * nd + 255) / 256), 256, 0, st>>>(L, gp, gm, ngeom, nd, dof1, eps, hess);
  return cudaGetLastError();
}
cudaError_t launch_um_band(int n, int ndim, int natom, const double* hess, const double* mass, double betan, int singlewell,
                           double* band, cudaStream_t st) {
  const long nd = (long)ndim * natom, tot = (long)n * nd * nd;
  cudaError_t e = cudaMemsetAsync(band, 0, sizeof(double) * (nd + 1) * n * nd, st);
  if (e != cudaSuccess) return e;
  um_band_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, ndim, natom, hess, mass, betan, singlewell, band);
  return cudaGetLastError();
}
cudaError_t launch_band_to_dense(long N, int kd, const double* band, double* A, cudaStream_t st) {
  band_to_dense_kernel<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(N, kd, band, A);
  return cudaGetLastError();
}

}  // namespace pimdk
