// CCpol-8sf batched energy and finite-difference gradient (sm_100a): a four-stage kernel pipeline.
// Compiled twice by the build: -fmad=false -DPIMDK_CCPOL_STRICT=1 (bit-faithful to the oracle's
// operation order; default at run time) and -fmad=true -DPIMDK_CCPOL_STRICT=0 ("fast").
//
// Replaces mcmod_waterdimer_ccpol.f90:18-58 (V, Vprime) called once per bead by step_v
// (verletmodule.f90:572-573) and by UM/UMprime (instantonmod.f90:26,79).
//
//   stage 0  setup    thread = energy   displaced geometry (Vprime's in-place +eps/-2eps/+eps walk), COM
//                                       alignment, Radau embedding, PJT2 monomers   -> 36 coordinates + emon
//   stage 1  sapt     thread = (energy, flexible|rigid geometry)   SAPT-5s'f site-site sum + dipole induction
//   stage 2a rigid    thread = energy   CCpol-8s: iterated induction (indN_iter), damped electrostatics, dispersion
//   stage 2b sweep    8 lanes = energy  CCpol-8s: U0's 625-pair exponential sweep; each lane owns whole bins of
//                                       aj(144) and walks them in the reference's pair order (registers ->
//                                       shared memory, no local memory); lane 0 then forms Erigid
//   stage 3  combine  thread = (geometry, component)   V+ and V-  -> central difference, drift write-back
// Staging buffers are structure-of-arrays [field][energy] so every stage reads and writes coalesced;
// 320 B per energy against ~1e5 FP64 operations.  Each stage has its own register budget / occupancy.
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

#ifndef PIMDK_SAPT_MINB
#define PIMDK_SAPT_MINB 1
#endif
#ifndef PIMDK_RIGID_MINB
#define PIMDK_RIGID_MINB 5
#endif
#ifndef PIMDK_SWEEP_MINB
#define PIMDK_SWEEP_MINB 5
#endif
constexpr int kSetupBlock = 128;
#ifndef PIMDK_SAPT_BLOCK
#define PIMDK_SAPT_BLOCK 512
#endif
constexpr int kSaptBlock = PIMDK_SAPT_BLOCK;

constexpr int kRigidBlock = 128;
constexpr int kSweepBlock = 128;                          // 16 energies x 8 lanes
constexpr int kSweepEnergies = kSweepBlock / kSweepLanes;
constexpr int kSweepDoubles = 75 + 75 + 148;              // sites of A, sites of B, 4 x (36 bins + 1 dummy), per energy
// staging fields per energy: 18 flexible + 18 rigid coordinates, emon, val, vall, erigid, eind, a0u, then the
// SAPT-5s'f sites: 4 blocks (flexible A, flexible B, rigid A, rigid B) of 24 site coordinates + 3 symmetry coordinates
constexpr int kSiteFields = 27;
constexpr int kFields = 44 + 4 * kSiteFields;
enum { F_FLEX = 0, F_RIGID = 18, F_EMON = 36, F_VAL = 37, F_VALL = 38, F_ERIG = 39, F_EIND = 40, F_A0U = 41, F_FCIND = 42, F_SITES = 44 };
constexpr int kTabBytes = (int)((sizeof(CcpolDev) + 15) / 16 * 16);
constexpr int kRigidTableBytes = (int)PIMDK_RIGID_TABLE_BYTES;
constexpr int kSaptTableBytes = kTabBytes - kRigidTableBytes;
static_assert(kRigidTableBytes % 16 == 0 && kSaptTableBytes % 16 == 0, "table blocks are staged in 16-byte granules");

template <int BYTES>
__device__ __forceinline__ const CcpolDev& stage_tables(const CcpolDev* __restrict__ g, unsigned char* smem) {
  static_assert(BYTES % 16 == 0, "16-byte granules");
  const int4* src = reinterpret_cast<const int4*>(g);
  int4* dst = reinterpret_cast<int4*>(smem);
  for (int i = threadIdx.x; i < BYTES / (int)sizeof(int4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return *reinterpret_cast<const CcpolDev*>(smem);
}

// ---- stage 0 ------------------------------------------------------------------------------------
// grad = 1: energy e = 36*g + 2*c + s is the displaced geometry for component c (loop order i=dim outer,
// j=atom inner: c = i*6 + j), s = 0 for +eps, 1 for -eps; components already visited by the reference's
// loop carry its round-off drift x+eps-2eps+eps (mcmod_waterdimer_ccpol.f90:48-52).
__global__ void __launch_bounds__(kSetupBlock)
KNAME(ccpol_setup_kernel)(int iemonomer, GeomLayout L, const double* __restrict__ x, long geom0, long ne, int grad,
                          double* __restrict__ buf) {
  const long e = (long)blockIdx.x * kSetupBlock + threadIdx.x;
  if (e >= ne) return;
  const double eps = 1e-4;
  const long g = geom0 + (grad ? e / 36 : e);
  const int t = grad ? (int)(e % 36) : 0;
  const int c = t >> 1, minus = t & 1;
  const long base = L.base(g);
  double xb[18];
#pragma unroll
  for (int d = 0; d < 18; ++d) {
    const double x0 = x[base + d * L.stride_dof];
    double val = x0;
    if (grad) {
      const int di = d % 3, dj = d / 3;  // d = atom*3 + dim
      const int cd = di * 6 + dj;        // its place in the loop order
      const double xp = x0 + eps;
      const double xm = xp - 2.0 * eps;
      const double xr = xm + eps;        // value left behind by the reference
      if (cd < c) val = xr;
      if (cd == c) val = minus ? xm : xp;
    }
    xb[d] = val;
  }
  double A[3][3], B[3][3], rg[6][3], emon;
  ccpol_setup(iemonomer, xb, A, B, rg, emon);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      buf[(F_FLEX + i * 3 + j) * ne + e] = A[i][j];
      buf[(F_FLEX + 9 + i * 3 + j) * ne + e] = B[i][j];
    }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) buf[(F_RIGID + i * 3 + j) * ne + e] = rg[i][j];
  buf[F_EMON * ne + e] = emon;
}

// ---- stage 1a -----------------------------------------------------------------------------------
// set_sites (proc_sapt5sf_new_ncd.f:1574-1758) for the four monomer geometries of an energy (flexible A, B and
// embedded-rigid A, B): thread = (energy, geometry block).  8 sites (Angstrom) + symmetry coordinates s1..s3.
struct GlobalSlots {  // set_sites' output sink: slot k of this thread's block lives at p[k * stride]
  double* p;
  long stride;
  __device__ __forceinline__ double& operator[](int k) const { return p[k * stride]; }
};
__global__ void __launch_bounds__(128)
KNAME(ccpol_sites_kernel)(long ne, double* __restrict__ buf) {
  const long i = (long)blockIdx.x * 128 + threadIdx.x;
  if (i >= 4 * ne) return;
  const int blk = (int)(i / ne);             // 0 flexible A, 1 flexible B, 2 rigid A, 3 rigid B
  const long e = i - blk * ne;
  const int f0 = blk * 9;                    // F_FLEX + {0, 9}, F_RIGID + {0, 9}
  const double a0 = 0.529177249;
  double c[3][3], s[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) c[a][k] = fast_div(buf[(f0 + a * 3 + k) * ne + e], a0);
  GlobalSlots out{buf + (long)(F_SITES + blk * kSiteFields) * ne + e, ne};
  set_sites(c, out, 0, s);
  out[24] = s[0];
  out[25] = s[1];
  out[26] = s[2];
}

// ---- stage 1b -----------------------------------------------------------------------------------
// dipind (proc_sapt5sf_new_ncd.f:1363-1533): thread = item (an energy's flexible or embedded-rigid geometry).
// Kept apart from the site-pair kernel: it runs once per item but is 27 KB of code (cbrt, pow, damping), and
// the pair kernel's loop has to stay inside the 32 KB instruction cache.
struct GlobalSites {  // slot k of the item in the staging buffer: sites of A, s of A, sites of B, s of B
  const double* p;
  long stride;
  // k: 0..23 sites of A, 24..47 sites of B, 48..50 s of A, 51..53 s of B
  __device__ __forceinline__ double operator[](int k) const {
    const int f = k < 24 ? k : (k < 48 ? k + 3 : (k < 51 ? k - 24 : k));
    return __ldg(p + (long)f * stride);
  }
};
__global__ void __launch_bounds__(128)
KNAME(ccpol_dipind_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf) {
  extern __shared__ __align__(16) unsigned char smem[];
  {
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(tab) + kRigidTableBytes);
    int4* dst = reinterpret_cast<int4*>(smem);
    for (int i = threadIdx.x; i < kSaptTableBytes / (int)sizeof(int4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  const CcpolDev& T = *reinterpret_cast<const CcpolDev*>(smem - kRigidTableBytes);
  const long j = (long)blockIdx.x * 128 + threadIdx.x;
  if (j >= 2 * ne) return;
  const int which = j >= ne;
  const long e = which ? j - ne : j;
  GlobalSites S{buf + (long)(F_SITES + which * 2 * kSiteFields) * ne + e, ne};
  const double sa[3] = {__ldg(S.p + 24 * ne), __ldg(S.p + 25 * ne), __ldg(S.p + 26 * ne)};
  const double sb[3] = {__ldg(S.p + 51 * ne), __ldg(S.p + 52 * ne), __ldg(S.p + 53 * ne)};
  buf[(F_FCIND + which) * ne + e] = dipind(T, S, sa, sb);
}

// ---- stage 1c -----------------------------------------------------------------------------------
// poten's 8 x 8 site-pair sum (proc_sapt5sf_new_ncd.f:130-213) + dipind: thread = item.  Sites and symmetry
// coordinates come from stage 1a through the staging buffer ([slot][item]: coalesced, L1-resident for the
// CTA's lifetime), so the kernel needs no per-thread shared-memory scratch; parameter tables in shared memory
// (every read warp-uniform -> broadcast).  One 512-thread CTA per SM: the pair loop is ~45 KB of straight-line
// FP64 code, more than the instruction cache, and every thread follows the same path through it; warps that
// start together stay close enough in the code that one warp's instruction fetch serves the others (measured:
// 128-thread CTAs at equal or higher occupancy are 11% slower and stall on instruction fetch).
__global__ void __launch_bounds__(kSaptBlock, PIMDK_SAPT_MINB)
KNAME(ccpol_sapt_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf) {
  extern __shared__ __align__(16) unsigned char smem[];
  // stage only the SAPT-5s'f members (param .. sapt_ntask); T is a view whose leading (rigid) members are not backed
  {
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(tab) + kRigidTableBytes);
    int4* dst = reinterpret_cast<int4*>(smem);
    for (int i = threadIdx.x; i < kSaptTableBytes / (int)sizeof(int4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  const CcpolDev& T = *reinterpret_cast<const CcpolDev*>(smem - kRigidTableBytes);
  const long j = (long)blockIdx.x * kSaptBlock + threadIdx.x;
  if (j >= 2 * ne) return;
  const int which = j >= ne;    // 0: flexible geometry (val), 1: embedded rigid geometry (vall)
  const long e = which ? j - ne : j;
  GlobalSites S{buf + (long)(F_SITES + which * 2 * kSiteFields) * ne + e, ne};
  const double sa[3] = {S[48], S[49], S[50]};
  const double sb[3] = {S[51], S[52], S[53]};
  const double val = sapt_pair_sum(T, S, sa, sb);
  buf[(which ? F_VALL : F_VAL) * ne + e] = val + buf[(F_FCIND + which) * ne + e];
}

// ---- stage 2a -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRigidBlock, PIMDK_RIGID_MINB)
KNAME(ccpol_rigid_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CcpolDev& T = stage_tables<kRigidTableBytes>(tab, smem);  // only the CCpol-8s members are valid here
  const long e = (long)blockIdx.x * kRigidBlock + threadIdx.x;
  if (e >= ne) return;
  double rg[6][3];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) rg[i][k] = buf[(F_RIGID + i * 3 + k) * ne + e];
  Frame fa, fb;
  rigid_frames(rg, fa, fb);
  int fl = 0;
  buf[F_EIND * ne + e] = ind2_iter(T, fa, fb, &fl);
  buf[F_A0U * ne + e] = u0_elst_disp(T, fa, fb);
  if (fl) atomicOr(flags, PIMDK_FLAG_NOCONV);
}

// ---- stage 2b -----------------------------------------------------------------------------------
// U0 (proc_ccpol8s-dimer_xyz_ncd.f:118-233): the reference walks the 25x25 site pairs in (nsA, nsB) order and
// adds e^{-beta R} R^p (p = 0..3) into one of 36x4 bins aj(ind) chosen by the pair of site classes.  The sums
// of different bins are independent, so the 36 bins are dealt to 8 lanes (static longest-processing-time
// schedule built on the host, ~79 pairs per lane); each lane walks its bins' pairs in the reference's order
// with the four sums in registers, i.e. performs exactly the reference's additions per bin, and writes each
// finished bin once to shared memory.  Lane 0 then forms E = Eind + sum_nl c(nl) aj(nl) + a0 in the
// reference's order (:105-110).
__global__ void __launch_bounds__(kSweepBlock, PIMDK_SWEEP_MINB)
KNAME(ccpol_sweep_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CcpolDev& T = stage_tables<kRigidTableBytes>(tab, smem);
  double* es = reinterpret_cast<double*>(smem + kRigidTableBytes) + (threadIdx.x / kSweepLanes) * kSweepDoubles;
  double* sA = es;          // sites of monomer A, 25 x 3
  double* sB = es + 75;     // sites of monomer B
  double* aj = es + 150;    // finished bins
  const int lane = threadIdx.x % kSweepLanes;
  const long e = (long)blockIdx.x * kSweepEnergies + threadIdx.x / kSweepLanes;
  const bool active = e < ne;
  const long ec = active ? e : ne - 1;   // lanes of a partial last group still take part in the warp syncs
  {
    double rg[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) rg[i][k] = buf[(F_RIGID + i * 3 + k) * ne + ec];
    Frame fa, fb;
    rigid_frames(rg, fa, fb);
    // fill_sites (:487-548): the group's 50 sites, dealt round-robin to its lanes
#pragma unroll 1
    for (int s = lane; s < 50; s += kSweepLanes) {
      double r[3];
      if (s < 25) {
        frame_site(T, fa, s, r);
        sA[s * 3 + 0] = r[0]; sA[s * 3 + 1] = r[1]; sA[s * 3 + 2] = r[2];
      } else {
        frame_site(T, fb, s - 25, r);
        sB[(s - 25) * 3 + 0] = r[0]; sB[(s - 25) * 3 + 1] = r[1]; sB[(s - 25) * 3 + 2] = r[2];
      }
    }
  }
  __syncwarp();
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  const uint64_t* sched = T.sweep[lane];
  const int nquads = T.sweep_quads;
  // one quad (four site pairs of one bin) per iteration: the four distance / sqrt / exp chains are
  // independent instruction streams, the sums are then added in the reference's order
#pragma unroll 1
  for (int i = 0; i < nquads; ++i) {
    const uint64_t w = sched[i];
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    const int bin = (hi >> 24) & 0x3f;
    const double beta = T.bin_beta[bin];
    double R[4], e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t d = (uint32_t)(w >> (14 * q));
      u0_pair(&sA[d & 0x7f], &sB[(d >> 7) & 0x7f], beta, R[q], e[q]);
    }
    (void)lo;
    if (hi & (1u << 30)) { acc0 = 0.0; acc1 = 0.0; acc2 = 0.0; acc3 = 0.0; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc0 = acc0 + e[q];
      acc1 = acc1 + e[q] * R[q];
      acc2 = acc2 + e[q] * R[q] * R[q];
      acc3 = acc3 + e[q] * R[q] * R[q] * R[q];
    }
    if (hi & (1u << 31)) {
      aj[bin] = acc0; aj[bin + 37] = acc1; aj[bin + 74] = acc2; aj[bin + 111] = acc3;
    }
  }
  if (lane == kSweepLanes - 1) {  // bins outside the quad schedule (the single O-O pair), pair by pair
    for (int i = 0; i < T.sweep_ntail; ++i) {
      const uint32_t d = T.sweep_tail[i];
      const int bin = (d >> 14) & 0x3f;
      double R, ee;
      u0_pair(&sA[d & 0x7f], &sB[(d >> 7) & 0x7f], T.bin_beta[bin], R, ee);
      if (d & (1u << 20)) { acc0 = 0.0; acc1 = 0.0; acc2 = 0.0; acc3 = 0.0; }
      acc0 = acc0 + ee;
      acc1 = acc1 + ee * R;
      acc2 = acc2 + ee * R * R;
      acc3 = acc3 + ee * R * R * R;
      if (d & (1u << 21)) {
        aj[bin] = acc0; aj[bin + 37] = acc1; aj[bin + 74] = acc2; aj[bin + 111] = acc3;
      }
    }
  }
  __syncwarp();
  if (active && lane == 0) {
    double E = buf[F_EIND * ne + e];
#pragma unroll 4
    for (int nl = 0; nl < 144; ++nl) E = E + T.cc[nl] * aj[nl + nl / 36];  // bins stored 37 apart (dummy bin 36)
    E = E + buf[F_A0U * ne + e];
    buf[F_ERIG * ne + e] = E * 627.510;
  }
}

// ---- stage 3 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
KNAME(ccpol_combine_kernel)(int iemonomer, double V0, GeomLayout L, double* __restrict__ x, long geom0, long ne, int grad,
                            const double* __restrict__ buf, double* __restrict__ v, double* __restrict__ gradout,
                            int write_drift, int* __restrict__ flags) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const double eps = 1e-4;
  if (!grad) {
    if (i >= ne) return;
    v[geom0 + i] = ccpol_combine(iemonomer, V0, buf[F_ERIG * ne + i], buf[F_VAL * ne + i], buf[F_VALL * ne + i],
                                 buf[F_EMON * ne + i]);
    return;
  }
  if (i >= ne / 2) return;  // one thread per (geometry, component)
  const long ep = 2 * i, em = ep + 1;
  const double vp = ccpol_combine(iemonomer, V0, buf[F_ERIG * ne + ep], buf[F_VAL * ne + ep], buf[F_VALL * ne + ep],
                                  buf[F_EMON * ne + ep]);
  const double vm = ccpol_combine(iemonomer, V0, buf[F_ERIG * ne + em], buf[F_VAL * ne + em], buf[F_VALL * ne + em],
                                  buf[F_EMON * ne + em]);
  const long g = geom0 + i / 18;
  const int c = (int)(i % 18);
  const int ci = c / 6, cj = c - ci * 6;
  const long addr = L.base(g) + (long)(cj * 3 + ci) * L.stride_dof;
  const double gval = (vp - vm) / (2.0 * eps);
  gradout[addr] = gval;
  if (gval != gval) atomicOr(flags, PIMDK_FLAG_NAN);
  if (write_drift) {
    const double x0 = x[addr];
    const double xp = x0 + eps;
    const double xm = xp - 2.0 * eps;
    x[addr] = xm + eps;
  }
}

size_t sapt_smem() { return kSaptTableBytes; }
size_t dipind_smem() { return kSaptTableBytes; }
size_t rigid_smem() { return kRigidTableBytes; }
size_t sweep_smem() { return kRigidTableBytes + (size_t)kSweepEnergies * kSweepDoubles * sizeof(double); }

}  // namespace

size_t KNAME(ccpol_work_bytes)(long ngeom, int grad) {
  const long cap = grad ? 32768 : 1048576;  // geometries per pass (grad: 1.18 M energies, 377 MB)
  const long n = ngeom < cap ? ngeom : cap;
  return (size_t)(n < 1 ? 1 : n) * kFields * 8 * (grad ? 36 : 1);
}

cudaError_t KNAME(launch_ccpol)(const CcpolDev* tab, int iemonomer, double V0, GeomLayout L, double* x, double* v,
                                double* grad, long ngeom, int write_drift, int* flags, double* work, size_t work_bytes,
                                cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(KNAME(ccpol_sapt_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sapt_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(KNAME(ccpol_rigid_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rigid_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(KNAME(ccpol_sweep_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem());
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int g = grad != nullptr;
  const long chunk = (long)(work_bytes / ((size_t)kFields * 8 * (g ? 36 : 1)));
  if (chunk < 1) return cudaErrorInvalidValue;
  for (long g0 = 0; g0 < ngeom; g0 += chunk) {
    const long ng = (ngeom - g0 < chunk) ? ngeom - g0 : chunk;
    const long ne = ng * (g ? 36 : 1);
    KNAME(ccpol_setup_kernel)<<<(unsigned)((ne + kSetupBlock - 1) / kSetupBlock), kSetupBlock, 0, st>>>(iemonomer, L, x, g0, ne, g, work);
    KNAME(ccpol_sites_kernel)<<<(unsigned)((4 * ne + 127) / 128), 128, 0, st>>>(ne, work);
    KNAME(ccpol_dipind_kernel)<<<(unsigned)((2 * ne + 127) / 128), 128, dipind_smem(), st>>>(tab, ne, work);
    KNAME(ccpol_sapt_kernel)<<<(unsigned)((2 * ne + kSaptBlock - 1) / kSaptBlock), kSaptBlock, sapt_smem(), st>>>(tab, ne, work);
    KNAME(ccpol_rigid_kernel)<<<(unsigned)((ne + kRigidBlock - 1) / kRigidBlock), kRigidBlock, rigid_smem(), st>>>(tab, ne, work, flags);
    KNAME(ccpol_sweep_kernel)<<<(unsigned)((ne + kSweepEnergies - 1) / kSweepEnergies), kSweepBlock, sweep_smem(), st>>>(tab, ne, work);
    const long nt = g ? ne / 2 : ne;
    KNAME(ccpol_combine_kernel)<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(iemonomer, V0, L, x, g0, ne, g, work, v, grad, write_drift, flags);
  }
  return cudaGetLastError();
}

}  // namespace pimdk
