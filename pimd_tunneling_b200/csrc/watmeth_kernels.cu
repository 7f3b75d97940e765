// Water-methane rigid-body surface (watermethane.f90 wmrb / wmrb_grad behind mcmod_watmeth.f90's V, Vprime, Vdoubleprime):
// one thread per geometry, the 63 site pairs walked in the reference's order (water site outer, methane site inner), the
// parameter tables in shared memory (every read warp-uniform).  Built with -fmad=false: bit-identical to the oracle.
#include "kernels.h"
#include "watmeth.cuh"

namespace pimdk {
namespace {

__device__ __forceinline__ const WatMethTab& stage_tab(const WatMethTab* __restrict__ g, unsigned char* smem) {
  const int4* src = reinterpret_cast<const int4*>(g);
  int4* dst = reinterpret_cast<int4*>(smem);
  for (int i = threadIdx.x; i < (int)(sizeof(WatMethTab) / sizeof(int4)); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return *reinterpret_cast<const WatMethTab*>(smem);
}
static_assert(sizeof(WatMethTab) % 16 == 0, "staged in 16-byte granules");

__device__ __forceinline__ double wm_calcr(const double* xs, int i, int j) {   // calcr (:356-367)
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double d = xs[3 * i + k] - xs[3 * (8 + j) + k];
    s = s + d * d;
  }
  return sqrt(s);
}
__device__ __noinline__ double wm_energy(const WatMethTab& T, const WmGln& G, const double* xs) {   // wmrb (:328-334)
  double ereal = 0.0;
#pragma unroll 1
  for (int i = 0; i < kWmWater; ++i)
#pragma unroll 1
    for (int j = 0; j < kWmMethane; ++j) ereal = ereal + wm_pair_energy(T, G, i * 9 + j, wm_calcr(xs, i, j));
  return ereal;
}
__device__ __noinline__ void wm_gradient(const WatMethTab& T, const WmGln& G, const double* xs, double* g) {   // wmrb_grad (:293-303)
  for (int d = 0; d < 51; ++d) g[d] = 0.0;
#pragma unroll 1
  for (int i = 0; i < kWmWater; ++i)
#pragma unroll 1
    for (int j = 0; j < kWmMethane; ++j) {
      const double r12 = wm_calcr(xs, i, j);
      const double gt = wm_pair_gradient(T, G, i * 9 + j, r12);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double rk = xs[3 * i + k] - xs[3 * (8 + j) + k];
        g[3 * i + k] = g[3 * i + k] + (rk * gt / r12);
        g[3 * (8 + j) + k] = g[3 * (8 + j) + k] - (rk * gt / r12);
      }
    }
}

__global__ void __launch_bounds__(128)
watmeth_kernel(const WatMethTab* __restrict__ tab, GeomLayout L, const double* __restrict__ x, double* __restrict__ v,
               double* __restrict__ grad, long ngeom, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  const WatMethTab& T = stage_tab(tab, smem);
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngeom) return;
  const long base = L.base(g);
  double xs[51];
  for (int d = 0; d < 51; ++d) xs[d] = x[base + (long)d * L.stride_dof];
  const WmGln G{wm_gammln(7.0), wm_gammln(9.0), wm_gammln(11.0)};
  if (v) v[g] = wm_energy(T, G, xs);   // mcmod_watmeth.f90:15-27: V0 is not subtracted
  if (grad) {
    double gg[51];
    wm_gradient(T, G, xs, gg);
    bool bad = false;
    for (int d = 0; d < 51; ++d) {
      grad[base + (long)d * L.stride_dof] = gg[d];
      bad = bad || gg[d] != gg[d];
    }
    if (bad) atomicOr(flags, PIMDK_FLAG_NAN);
  }
}

// Vdoubleprime (mcmod_watmeth.f90:42-62): central difference, eps = 1e-4, of the analytic gradient; x(i,j) perturbed in place
// (+eps, -2 eps, +eps) in the loop order i = dim outer, j = atom inner
__global__ void __launch_bounds__(64)
watmeth_hessian_kernel(const WatMethTab* __restrict__ tab, GeomLayout L, double* __restrict__ x, double* __restrict__ hess, long ngeom) {
  extern __shared__ __align__(16) unsigned char smem[];
  const WatMethTab& T = stage_tab(tab, smem);
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngeom) return;
  const long base = L.base(g);
  const double eps = 1e-4;
  double xs[51], gp[51], gm[51];
  for (int d = 0; d < 51; ++d) xs[d] = x[base + (long)d * L.stride_dof];
  const WmGln G{wm_gammln(7.0), wm_gammln(9.0), wm_gammln(11.0)};
  double* H = hess + g * 51 * 51;
#pragma unroll 1
  for (int i = 0; i < 3; ++i)
#pragma unroll 1
    for (int j = 0; j < kWmSites; ++j) {
      const int d1 = j * 3 + i;
      xs[d1] = xs[d1] + eps;
      wm_gradient(T, G, xs, gp);
      xs[d1] = xs[d1] - 2.0 * eps;
      wm_gradient(T, G, xs, gm);
      xs[d1] = xs[d1] + eps;
      for (int d2 = 0; d2 < 51; ++d2) H[d2 * 51 + d1] = (gp[d2] - gm[d2]) / (2.0 * eps);
    }
  for (int d = 0; d < 51; ++d) x[base + (long)d * L.stride_dof] = xs[d];
}

}  // namespace

cudaError_t launch_watmeth(const WatMethTab* tab, GeomLayout L, const double* x, double* v, double* grad, long ngeom, int* flags,
                           cudaStream_t st) {
  watmeth_kernel<<<(unsigned)((ngeom + 127) / 128), 128, sizeof(WatMethTab), st>>>(tab, L, x, v, grad, ngeom, flags);
  return cudaGetLastError();
}
cudaError_t launch_watmeth_hessian(const WatMethTab* tab, GeomLayout L, double* x, double* hess, long ngeom, cudaStream_t st) {
  watmeth_hessian_kernel<<<(unsigned)((ngeom + 63) / 64), 64, sizeof(WatMethTab), st>>>(tab, L, x, hess, ngeom);
  return cudaGetLastError();
}

}  // namespace pimdk
