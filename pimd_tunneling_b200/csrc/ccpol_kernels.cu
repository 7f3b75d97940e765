// CCpol-8sf batched energy and finite-difference gradient kernels (sm_100a).
// Compiled twice by the build: -fmad=false -DPIMDK_CCPOL_STRICT=1 (bit-faithful to the oracle's
// operation order; default at run time) and -fmad=true -DPIMDK_CCPOL_STRICT=0 ("fast").
//
// Replaces mcmod_waterdimer_ccpol.f90:18-58 (V, Vprime) called once per bead by step_v
// (verletmodule.f90:572-573) and by UM/UMprime (instantonmod.f90:26,79).
#if PIMDK_CCPOL_STRICT
#define PIMDK_CCPOL_NS ccpol_strict_impl
#else
#define PIMDK_CCPOL_NS ccpol_fast_impl
#endif
#include "ccpol_device.cuh"
#include "kernels.h"

namespace pimdk {

#if PIMDK_CCPOL_STRICT
#define KNAME(x) x##_strict
#else
#define KNAME(x) x##_fast
#endif

namespace {

constexpr int kBlock = PIMDK_CCPOL_BLOCK;

__device__ __forceinline__ void stage_tables(const CcpolDev* __restrict__ g, CcpolDev* s) {
  const int4* src = reinterpret_cast<const int4*>(g);
  int4* dst = reinterpret_cast<int4*>(s);
  for (int i = threadIdx.x; i < (int)(sizeof(CcpolDev) / sizeof(int4)); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// energies: one thread per geometry
__global__ void __launch_bounds__(kBlock, 1)
KNAME(ccpol_energy_kernel)(const CcpolDev* __restrict__ tab, GeomLayout L, const double* __restrict__ x,
                           double* __restrict__ v, long ngeom, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  CcpolDev* T = reinterpret_cast<CcpolDev*>(smem);
  double* scr_base = reinterpret_cast<double*>(smem + sizeof(CcpolDev));
  stage_tables(tab, T);
  Scratch scr{scr_base + threadIdx.x};
  for (long g = (long)blockIdx.x * kBlock + threadIdx.x; g < ngeom; g += (long)gridDim.x * kBlock) {
    double xb[18];
    const long base = L.base(g);
#pragma unroll
    for (int d = 0; d < 18; ++d) xb[d] = x[base + d * L.stride_dof];
    int fl = 0;
    v[g] = ccpol_V(*T, scr, xb, &fl);
    if (fl) atomicOr(flags, PIMDK_FLAG_NOCONV);
  }
}

// Vprime: 36 threads per geometry (component c in the reference's loop order i=dim outer, j=atom inner;
// even thread = +eps, odd thread = -eps), 7 whole geometries (252 threads) per CTA pass so that a
// geometry never straddles CTAs.  The in-place perturbation drift of the reference
// (x+eps, -2eps, +eps; mcmod_waterdimer_ccpol.f90:48-52) is reproduced and optionally written back.
constexpr int kGeomPerPass = kBlock / 36;

__global__ void __launch_bounds__(kBlock, 1)
KNAME(ccpol_grad_kernel)(const CcpolDev* __restrict__ tab, GeomLayout L, double* __restrict__ x,
                         double* __restrict__ grad, long ngeom, int write_drift, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  CcpolDev* T = reinterpret_cast<CcpolDev*>(smem);
  double* scr_base = reinterpret_cast<double*>(smem + sizeof(CcpolDev));
  stage_tables(tab, T);
  Scratch scr{scr_base + threadIdx.x};
  const double eps = 1e-4;
  const int lg = threadIdx.x / 36;           // geometry slot within the pass
  const int t = threadIdx.x - lg * 36;
  const int c = t >> 1;                      // component in loop order: c = i*6 + j  (i = dim, j = atom)
  const int minus = t & 1;
  const int ci = c / 6, cj = c - ci * 6;
  const int my_dof = cj * 3 + ci;            // position in x(3,6): atom-major
  const long npass = (ngeom + kGeomPerPass - 1) / kGeomPerPass;
  for (long pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const long g = pass * kGeomPerPass + lg;
    const bool active = lg < kGeomPerPass && g < ngeom;
    double xb[18];
    double vpm = 0.0;
    double mydrift = 0.0;
    const long base = active ? L.base(g) : 0;
    if (active) {
#pragma unroll
      for (int d = 0; d < 18; ++d) {
        double x0 = x[base + d * L.stride_dof];
        const int di = d % 3, dj = d / 3;   // d = atom*3 + dim
        const int cd = di * 6 + dj;         // its place in the loop order
        double xp = x0 + eps;
        double xm = xp - 2.0 * eps;
        double xr = xm + eps;               // value left behind by the reference
        double val = x0;
        if (cd < c) val = xr;
        if (cd == c) { val = minus ? xm : xp; mydrift = xr; }
        xb[d] = val;
      }
    }
    __syncthreads();  // every read of x above precedes every drift write below
    if (active) {
      int fl = 0;
      vpm = ccpol_V(*T, scr, xb, &fl);
      if (fl) atomicOr(flags, PIMDK_FLAG_NOCONV);
    }
    const double other = __shfl_xor_sync(0xffffffffu, vpm, 1);
    if (active && !minus) {
      const double gval = (vpm - other) / (2.0 * eps);
      grad[base + my_dof * L.stride_dof] = gval;
      if (gval != gval) atomicOr(flags, PIMDK_FLAG_NAN);
      if (write_drift) x[base + my_dof * L.stride_dof] = mydrift;
    }
  }
}

}  // namespace

size_t KNAME(ccpol_smem_bytes)() { return sizeof(CcpolDev) + (size_t)kScratchSlots * kBlock * sizeof(double); }

cudaError_t KNAME(launch_ccpol_energy)(const CcpolDev* tab, GeomLayout L, const double* x, double* v, long ngeom,
                                       int* flags, int num_sms, cudaStream_t st) {
  static bool attr = false;
  size_t sm = KNAME(ccpol_smem_bytes)();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(KNAME(ccpol_energy_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  long blocks = (ngeom + kBlock - 1) / kBlock;
  if (blocks > 8L * num_sms) blocks = 8L * num_sms;
  if (blocks < 1) blocks = 1;
  KNAME(ccpol_energy_kernel)<<<(unsigned)blocks, kBlock, sm, st>>>(tab, L, x, v, ngeom, flags);
  return cudaGetLastError();
}

cudaError_t KNAME(launch_ccpol_grad)(const CcpolDev* tab, GeomLayout L, double* x, double* grad, long ngeom,
                                     int write_drift, int* flags, int num_sms, cudaStream_t st) {
  static bool attr = false;
  size_t sm = KNAME(ccpol_smem_bytes)();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(KNAME(ccpol_grad_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  long blocks = (ngeom + kGeomPerPass - 1) / kGeomPerPass;
  // persistent-style grid: one resident CTA per SM, each looping over passes of 7 geometries
  if (blocks > 16L * num_sms) blocks = 16L * num_sms;
  if (blocks < 1) blocks = 1;
  KNAME(ccpol_grad_kernel)<<<(unsigned)blocks, kBlock, sm, st>>>(tab, L, x, grad, ngeom, write_drift, flags);
  return cudaGetLastError();
}

}  // namespace pimdk
