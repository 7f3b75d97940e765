// CCpol-8sf batched energy and finite-difference gradient (sm_100a): a pipeline of seven kernels per pass.
// Compiled twice by the build: -fmad=false -DPIMDK_CCPOL_STRICT=1 (bit-faithful to the oracle's
// operation order; default at run time) and -fmad=true -DPIMDK_CCPOL_STRICT=0 ("fast").
//
// Replaces mcmod_waterdimer_ccpol.f90:18-58 (V, Vprime) called once per bead by step_v
// (verletmodule.f90:572-573) and by UM/UMprime (instantonmod.f90:26,79).
//
//   stage 0  setup    thread = energy   displaced geometry (Vprime's in-place +eps/-2eps/+eps walk), COM
//                                       alignment, Radau or Eckart embedding, PJT2 monomers -> 36 coordinates + emon
//   stage 1a sites    thread = (energy, monomer geometry)   set_sites + the per-monomer part of dipind
//   stage 1b dipind   thread = item (energy x {flexible, rigid})   the pair part of dipind
//   stage 1c sapt     thread = item     SAPT-5s'f 8 x 8 site-pair sum
//   stage 2a rigid    thread = energy   CCpol-8s: frames, iterated induction (indN_iter), damped electrostatics, dispersion
//   stage 2b sweep    CTA = 32 energies x 10 warps (lane = energy, warp = bin)   CCpol-8s: U0's 625-pair exponential
//                                       sweep, each bin walked in the reference's pair order; warp 0 then forms Erigid
//   stage 3  combine  thread = (geometry, component)   V+ and V-  -> central difference, drift write-back
// Staging buffers are structure-of-arrays [field][energy] so every stage reads and writes coalesced (1.5 KB per
// energy against ~6e4 FP64 operations).  Each stage has its own register budget / occupancy.  The pair-sum, rigid and
// sweep kernels take their parameter tables as kernel parameters (constant bank): every table read in them is warp-uniform.
#if PIMDK_CCPOL_STRICT
#define PIMDK_CCPOL_NS ccpol_strict_impl
#else
#define PIMDK_CCPOL_NS ccpol_fast_impl
#endif
// the sweep kernel reads the exp table from shared memory (pimdk_exp_nonpos_sh); the other kernels through the read-only path
#define PIMDK_EXP2_SHARED 1
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
#ifndef PIMDK_SETUP_MINB
#define PIMDK_SETUP_MINB 1
#endif
#ifndef PIMDK_DIPIND_MINB
#define PIMDK_DIPIND_MINB 1
#endif
#ifndef PIMDK_RIGID_MINB
#define PIMDK_RIGID_MINB 5
#endif
#ifndef PIMDK_SWEEP_MINB
#define PIMDK_SWEEP_MINB 3
#endif
constexpr int kSetupBlock = 128;
#ifndef PIMDK_SAPT_BLOCK
#define PIMDK_SAPT_BLOCK 512
#endif
constexpr int kSaptBlock = PIMDK_SAPT_BLOCK;

constexpr int kRigidBlock = 128;
// three CTAs of ten warps per SM (64 registers, no spills; possible since the tables left shared memory: 3 x 76.3 KB):
// 1.976 ms per pass against 2.044 ms for two CTAs of twelve warps at 80 registers (352 x 3: 2.071, 384 x 3: 2.033, 288 x 3 and
// 256 x 3: 2.05, 512 x 2 at 64 registers: 2.13; profiles/r2_sapt_split_and_param_tables.md)
#ifndef PIMDK_SWEEP_BLOCK
#define PIMDK_SWEEP_BLOCK 320
#endif
constexpr int kSweepBlock = PIMDK_SWEEP_BLOCK;            // warps that share 32 energies in the U0 sweep
// staging fields per energy: 18 flexible + 18 rigid coordinates, emon, val, vall, erigid, eind, a0u, then the
// SAPT-5s'f sites: 4 blocks (flexible A, flexible B, rigid A, rigid B) of 24 site coordinates + 3 symmetry coordinates
constexpr int kSiteFields = 27;
constexpr int kFrameFields = 24;                 // body frames of the two rigid monomers (ex, ey, ez, com) x 2
// stage 1a also forms dipind's per-monomer dipole sum and polarisability (4 values per monomer geometry, appended to the
// staging buffer), so that stage 1b reads 14 values per item instead of 54
constexpr int kDipFields = 16;
constexpr int kFields = 44 + 4 * kSiteFields + kFrameFields + kDipFields;
enum { F_FLEX = 0, F_RIGID = 18, F_EMON = 36, F_VAL = 37, F_VALL = 38, F_ERIG = 39, F_EIND = 40, F_A0U = 41, F_FCIND = 42, F_SITES = 44, F_FRAME = 44 + 4 * kSiteFields,
       F_DIP = 44 + 4 * kSiteFields + kFrameFields };
constexpr int kTabBytes = (int)((sizeof(CcpolDev) + 15) / 16 * 16);
constexpr int kRigidTableBytes = (int)PIMDK_RIGID_TABLE_BYTES;
constexpr int kSaptTableBytes = kTabBytes - kRigidTableBytes;
static_assert(kRigidTableBytes % 16 == 0 && kSaptTableBytes % 16 == 0, "table blocks are staged in 16-byte granules");

// ---- stage 0 ------------------------------------------------------------------------------------
// grad = 1: energy e = 36*g + 2*c + s is the displaced geometry for component c (loop order i=dim outer,
// j=atom inner: c = i*6 + j), s = 0 for +eps, 1 for -eps; components already visited by the reference's
// loop carry its round-off drift x+eps-2eps+eps (mcmod_waterdimer_ccpol.f90:48-52).
__global__ void __launch_bounds__(kSetupBlock, PIMDK_SETUP_MINB)
KNAME(ccpol_setup_kernel)(int iemonomer, int iembed, GeomLayout L, const double* __restrict__ x, long geom0, long ne, int grad,
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
  ccpol_setup(iemonomer, iembed, xb, A, B, rg, emon);
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
struct TeeSlots {  // set_sites' sink that also keeps the values in registers for dipind's per-monomer part
  double* p;
  long stride;
  double* loc;
  struct Ref {
    double* g;
    double* l;
    __device__ __forceinline__ void operator=(double v) const { *g = v; *l = v; }
  };
  __device__ __forceinline__ Ref operator[](int k) const { return Ref{p + k * stride, loc + k}; }
};
__global__ void __launch_bounds__(128)
KNAME(ccpol_sites_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf) {
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
  double loc[24];
  TeeSlots out{buf + (long)(F_SITES + blk * kSiteFields) * ne + e, ne, loc};
  set_sites(c, out, 0, s);
  buf[(long)(F_SITES + blk * kSiteFields + 24) * ne + e] = s[0];
  buf[(long)(F_SITES + blk * kSiteFields + 25) * ne + e] = s[1];
  buf[(long)(F_SITES + blk * kSiteFields + 26) * ne + e] = s[2];
  double dm[3], polis;
  dipind_monomer(*tab, [&](int k) { return loc[k]; }, blk & 1, s, dm, polis);
  double* dip = buf + (long)(F_DIP + blk * 4) * ne + e;
  dip[0] = dm[0];
  dip[ne] = dm[1];
  dip[2 * ne] = dm[2];
  dip[3 * ne] = polis;
}

// ---- stage 1b -----------------------------------------------------------------------------------
// the pair part of dipind (proc_sapt5sf_new_ncd.f:1363-1533): thread = item (an energy's flexible or embedded-rigid
// geometry); the per-monomer part ran in stage 1a.  Kept apart from the site-pair kernel: it runs once per item and is
// 10 KB of code (cbrt, pow, damping) that the pair kernel's loop should not have to fetch around.
struct GlobalSites {  // slot k of the item in the staging buffer: sites of A, s of A, sites of B, s of B
  const double* p;
  long stride;
  // k: 0..23 sites of A, 24..47 sites of B, 48..50 s of A, 51..53 s of B
  __device__ __forceinline__ double operator[](int k) const {
    const int f = k < 24 ? k : (k < 48 ? k + 3 : (k < 51 ? k - 24 : k));
    return __ldg(p + (long)f * stride);
  }
};
__global__ void __launch_bounds__(128, PIMDK_DIPIND_MINB)
KNAME(ccpol_dipind_kernel)(const CcpolDev* __restrict__ tab, long ne, double* __restrict__ buf) {
  const long j = (long)blockIdx.x * 128 + threadIdx.x;
  if (j >= 2 * ne) return;
  const int which = j >= ne;
  const long e = which ? j - ne : j;
  const double* sitesA = buf + (long)(F_SITES + which * 2 * kSiteFields) * ne + e;   // O is site 0 of a block
  const double* sitesB = sitesA + (long)kSiteFields * ne;
  const double* dipA = buf + (long)(F_DIP + which * 8) * ne + e;
  const double* dipB = dipA + 4 * ne;
  double Oa[3], Ob[3], dma[3], dmb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Oa[k] = __ldg(sitesA + k * ne);
    Ob[k] = __ldg(sitesB + k * ne);
    dma[k] = __ldg(dipA + k * ne);
    dmb[k] = __ldg(dipB + k * ne);
  }
  buf[(F_FCIND + which) * ne + e] = dipind_pair(__ldg(&tab->parab[10 - 1]), Oa, Ob, dma, dmb, __ldg(dipA + 3 * ne), __ldg(dipB + 3 * ne));
}

// ---- stage 1c -----------------------------------------------------------------------------------
// poten's 8 x 8 site-pair sum (proc_sapt5sf_new_ncd.f:130-213) + dipind: thread = item.  Sites and symmetry
// coordinates come from stage 1a through the staging buffer ([slot][item]: coalesced, L1-resident for the
// CTA's lifetime); parameter tables in the constant bank (every read is warp-uniform: LDCU into a uniform register or a
// c[0][..] operand).  One 512-thread CTA per SM: the pair loop is ~45 KB of straight-line
// FP64 code, more than the instruction cache, and every thread follows the same path through it; warps that
// start together stay close enough in the code that one warp's instruction fetch serves the others (measured:
// 128-thread CTAs at equal or higher occupancy are 11% slower and stall on instruction fetch).
// PIMDK_SAPT_PARAM_TABLES: the SAPT-5s'f tables as a kernel parameter (strict build: 2.459 -> 2.315 ms per pass) or staged per
// CTA into shared memory as in round 1 (fast build: with contraction on, ptxas schedules the parameter form worse, 2.55 against
// 2.23 ms, so that build keeps the shared-memory form)
#ifndef PIMDK_SAPT_PARAM_TABLES
#define PIMDK_SAPT_PARAM_TABLES PIMDK_CCPOL_STRICT
#endif
#if PIMDK_SAPT_PARAM_TABLES
#define PIMDK_SAPT_TABARG const __grid_constant__ SaptParams T
#define PIMDK_SAPT_TABVAL g_sapt
constexpr int kSaptSmemTab = 0;
#else
#define PIMDK_SAPT_TABARG const CcpolDev* __restrict__ tab
#define PIMDK_SAPT_TABVAL tab
constexpr int kSaptSmemTab = kSaptTableBytes;
#endif
template <bool OLD>
__global__ void __launch_bounds__(kSaptBlock, PIMDK_SAPT_MINB)
KNAME(ccpol_sapt_kernel)(PIMDK_SAPT_TABARG, long ne, double* __restrict__ buf) {
  extern __shared__ __align__(16) unsigned char smem[];
#if !PIMDK_SAPT_PARAM_TABLES
  // stage only the SAPT-5s'f members (param .. pairflags); T is a view whose leading (rigid) members are not backed
  {
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(tab) + kRigidTableBytes);
    int4* dst = reinterpret_cast<int4*>(smem);
    for (int i = threadIdx.x; i < kSaptTableBytes / (int)sizeof(int4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  const CcpolDev& T = *reinterpret_cast<const CcpolDev*>(smem - kRigidTableBytes);
#endif
  const long j = (long)blockIdx.x * kSaptBlock + threadIdx.x;
  if (j >= 2 * ne) return;
  const int which = j >= ne;    // 0: flexible geometry (val), 1: embedded rigid geometry (vall)
  const long e = which ? j - ne : j;
  GlobalSites S{buf + (long)(F_SITES + which * 2 * kSiteFields) * ne + e, ne};
  // the 8 sites of B are read once per site of A: keep them in shared memory, slot-major (conflict-free)
  Scratch<kSaptBlock> sitesB{reinterpret_cast<double*>(smem + kSaptSmemTab) + threadIdx.x};
#pragma unroll
  for (int k = 0; k < 24; ++k) sitesB[k] = S[24 + k];
  const double sa[3] = {S[48], S[49], S[50]};
  const double sb[3] = {S[51], S[52], S[53]};
  Scratch<kSaptBlock> qb{reinterpret_cast<double*>(smem + kSaptSmemTab) + 24 * kSaptBlock + threadIdx.x};
  const double val = sapt_pair_sum<OLD>(T, S, sitesB, qb, sa, sb);
  buf[(which ? F_VALL : F_VAL) * ne + e] = val + buf[(F_FCIND + which) * ne + e];
}

// ---- stage 2a -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRigidBlock, PIMDK_RIGID_MINB)
KNAME(ccpol_rigid_kernel)(const __grid_constant__ RigidParams T, long ne, double* __restrict__ buf, int* __restrict__ flags) {
  const long e = (long)blockIdx.x * kRigidBlock + threadIdx.x;
  if (e >= ne) return;
  double rg[6][3];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) rg[i][k] = buf[(F_RIGID + i * 3 + k) * ne + e];
  Frame fa, fb;
  rigid_frames(rg, fa, fb);
#pragma unroll
  for (int k = 0; k < 3; ++k) {   // the sweep stage (2b) builds the 50 sites from these frames
    buf[(F_FRAME + 0 + k) * ne + e] = fa.ex[k];
    buf[(F_FRAME + 3 + k) * ne + e] = fa.ey[k];
    buf[(F_FRAME + 6 + k) * ne + e] = fa.ez[k];
    buf[(F_FRAME + 9 + k) * ne + e] = fa.com[k];
    buf[(F_FRAME + 12 + k) * ne + e] = fb.ex[k];
    buf[(F_FRAME + 15 + k) * ne + e] = fb.ey[k];
    buf[(F_FRAME + 18 + k) * ne + e] = fb.ez[k];
    buf[(F_FRAME + 21 + k) * ne + e] = fb.com[k];
  }
  int fl = 0;
  buf[F_EIND * ne + e] = ind2_iter(T, fa, fb, &fl);
  buf[F_A0U * ne + e] = u0_elst_disp(T, fa, fb);
  if (fl) atomicOr(flags, PIMDK_FLAG_NOCONV);
}

// ---- stage 2b -----------------------------------------------------------------------------------
// U0 (proc_ccpol8s-dimer_xyz_ncd.f:118-233): the reference walks the 25x25 site pairs in (nsA, nsB) order and
// adds e^{-beta R} R^p (p = 0..3) into one of 36x4 bins aj(ind) chosen by the pair of site classes.  The sums
// of different bins are independent.  A CTA owns 32 energies; lane = energy, so a warp-level task is uniform in
// control flow and addresses: the warps first build the 50 sites (fill_sites :487-548) from the frames of stage
// 2a into shared memory (slot-major [slot][energy], conflict-free), then pull bins from a shared counter,
// largest first, and walk each bin's site pairs in the reference's order (block (ca,cb), then block (cb,ca))
// with the four sums in registers — exactly the reference's additions per bin; four consecutive pairs are
// evaluated as independent instruction streams before their terms are added in order.  A finished bin leaves
// c(nl)*aj(nl) in shared memory; warp 0 then forms E = Eind + sum_nl c(nl) aj(nl) + a0 in the reference's order
// (:105-110).
constexpr int kSweepWarps = kSweepBlock / 32;
constexpr int kSweepSlots = 150 + 144;      // sites of A (75), sites of B (75), c(nl)*aj(nl)

// four (or fewer) consecutive pairs of a block: pair k -> (A site a0 + k / nb, B site b0 + k % nb)
template <int N>
__device__ __forceinline__ void sweep_chunk(const double* sA, const double* sB, int a0, int b0, int nb, int k0, double beta,
                                            double& acc0, double& acc1, double& acc2, double& acc3) {
  double R[N], e[N];
  const int sh = nb >> 1;   // log2(nb) for nb = 1, 2, 4
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const int k = k0 + q, i = k >> sh, j = k & (nb - 1);   // class sizes are 1, 2 or 4 (checked with the tables)
    const double* ra = sA + (a0 + i) * 96;     // 3 slots x 32 energies per site
    const double* rb = sB + (b0 + j) * 96;
    double r12 = ra[0] - rb[0];
    double d = r12 * r12;                      // the reference's 0 + r12*r12: a square is never -0
    r12 = ra[32] - rb[32];
    d = d + r12 * r12;
    r12 = ra[64] - rb[64];
    d = d + r12 * r12;
    R[q] = fast_sqrt(d);
    e[q] = pimdk_exp_nonpos_sh(-beta * R[q]);   // beta >= 0 (checked when the tables are built), R >= 0
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    acc0 = acc0 + e[q];
    acc1 = acc1 + e[q] * R[q];
    acc2 = acc2 + e[q] * R[q] * R[q];
    acc3 = acc3 + e[q] * R[q] * R[q] * R[q];
  }
}
// one A site against the four B sites of a class (the common block shape): the A site is read once
__device__ __forceinline__ void sweep_row4(const double* ra, const double* sB, int b0, double beta, double& acc0,
                                           double& acc1, double& acc2, double& acc3) {
  const double ax = ra[0], ay = ra[32], az = ra[64];
  double R[4], e[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double* rb = sB + (b0 + q) * 96;
    double r12 = ax - rb[0];
    double d = r12 * r12;
    r12 = ay - rb[32];
    d = d + r12 * r12;
    r12 = az - rb[64];
    d = d + r12 * r12;
    R[q] = fast_sqrt(d);
    e[q] = pimdk_exp_nonpos_sh(-beta * R[q]);   // beta >= 0 (checked when the tables are built), R >= 0
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    acc0 = acc0 + e[q];
    acc1 = acc1 + e[q] * R[q];
    acc2 = acc2 + e[q] * R[q] * R[q];
    acc3 = acc3 + e[q] * R[q] * R[q] * R[q];
  }
}
__device__ __forceinline__ void sweep_block(const double* sA, const double* sB, int a0, int na, int b0, int nb, double beta,
                                            double& acc0, double& acc1, double& acc2, double& acc3) {
  if (nb == 4) {
#pragma unroll 1
    for (int i = 0; i < na; ++i) sweep_row4(sA + (a0 + i) * 96, sB, b0, beta, acc0, acc1, acc2, acc3);
    return;
  }
  const int np = na * nb;
  if (np >= 4) {
#pragma unroll 1
    for (int k0 = 0; k0 < np; k0 += 4) sweep_chunk<4>(sA, sB, a0, b0, nb, k0, beta, acc0, acc1, acc2, acc3);
  } else if (np == 2) {
    sweep_chunk<2>(sA, sB, a0, b0, nb, 0, beta, acc0, acc1, acc2, acc3);
  } else {
    sweep_chunk<1>(sA, sB, a0, b0, nb, 0, beta, acc0, acc1, acc2, acc3);
  }
}

__global__ void __launch_bounds__(kSweepBlock, PIMDK_SWEEP_MINB)
KNAME(ccpol_sweep_kernel)(const __grid_constant__ RigidParams T, long ne, double* __restrict__ buf) {
  pimdk_exp2_stage();
  extern __shared__ __align__(16) unsigned char smem[];
  double* slots = reinterpret_cast<double*>(smem);
  int* queue = reinterpret_cast<int*>(slots + kSweepSlots * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sA = slots + lane;            // site s, coordinate c of A at sA[(s*3 + c) * 32]
  double* sB = slots + 75 * 32 + lane;
  double* cw = slots + 150 * 32 + lane; // c(nl)*aj(nl) at cw[nl * 32]; the frames live here until the sites are built
  const long e = (long)blockIdx.x * 32 + lane;
  const bool active = e < ne;
  const long ec = active ? e : ne - 1;  // idle lanes of the last CTA recompute its last energy
  if (threadIdx.x == 0) *queue = 0;
  // the two scalars of the final sum are fetched now so that their latency is not part of the CTA's serial tail
  const double eind = buf[F_EIND * ne + ec], a0u = buf[F_A0U * ne + ec];
  for (int k = warp; k < kFrameFields; k += kSweepWarps) cw[k * 32] = buf[(long)(F_FRAME + k) * ne + ec];
  __syncthreads();
  // fill_sites (:487-548): the 50 sites of the 32 energies, dealt to the warps
  for (int s = warp; s < 50; s += kSweepWarps) {
    const int m = s >= 25, k = s - 25 * m;
    const double* f = cw + (12 * m) * 32;
    const double s1 = T.sites[k * 3 + 0], s2 = T.sites[k * 3 + 1], s3 = T.sites[k * 3 + 2];
    double* out = (m ? sB : sA) + k * 96;
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // frame_site: (ex s1 + ey s2 + ez s3) + com
      const double t = f[j * 32] * s1 + f[(3 + j) * 32] * s2 + f[(6 + j) * 32] * s3;
      out[j * 32] = t + f[(9 + j) * 32];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(queue, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= 36) break;
    const uint32_t w = T.tbins[t];
    const int a0 = w & 31, na = (w >> 5) & 7, b0 = (w >> 8) & 31, nb = (w >> 13) & 7, bin = (w >> 16) & 63;
    const double beta = T.bin_beta[bin];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    sweep_block(sA, sB, a0, na, b0, nb, beta, acc0, acc1, acc2, acc3);
    if (a0 != b0) sweep_block(sA, sB, b0, nb, a0, na, beta, acc0, acc1, acc2, acc3);
    cw[bin * 32] = T.cc[bin] * acc0;
    cw[(bin + 36) * 32] = T.cc[bin + 36] * acc1;
    cw[(bin + 72) * 32] = T.cc[bin + 72] * acc2;
    cw[(bin + 108) * 32] = T.cc[bin + 108] * acc3;
  }
  __syncthreads();
  if (warp == 0 && active) {
    double E = eind;
#pragma unroll 16
    for (int nl = 0; nl < 144; ++nl) E = E + cw[nl * 32];
    E = E + a0u;
    buf[F_ERIG * ne + e] = E * 627.510;
  }
}

// ---- stage 3 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
KNAME(ccpol_combine_kernel)(int iemonomer, int icc, double V0, GeomLayout L, double* __restrict__ x, long geom0, long ne, int grad,
                            const double* __restrict__ buf, double* __restrict__ v, double* __restrict__ gradout,
                            int write_drift, int* __restrict__ flags) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const double eps = 1e-4;
  if (!grad) {
    if (i >= ne) return;
    v[geom0 + i] = ccpol_combine(iemonomer, icc, V0, icc ? buf[F_ERIG * ne + i] : 0.0, buf[F_VAL * ne + i], buf[F_VALL * ne + i],
                                 buf[F_EMON * ne + i]);
    return;
  }
  if (i >= ne / 2) return;  // one thread per (geometry, component)
  const long ep = 2 * i, em = ep + 1;
  const double vp = ccpol_combine(iemonomer, icc, V0, icc ? buf[F_ERIG * ne + ep] : 0.0, buf[F_VAL * ne + ep], buf[F_VALL * ne + ep],
                                  buf[F_EMON * ne + ep]);
  const double vm = ccpol_combine(iemonomer, icc, V0, icc ? buf[F_ERIG * ne + em] : 0.0, buf[F_VAL * ne + em], buf[F_VALL * ne + em],
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

size_t sapt_smem() { return kSaptSmemTab + (size_t)(24 + 8) * kSaptBlock * sizeof(double); }
size_t sweep_smem() { return (size_t)kSweepSlots * 32 * sizeof(double) + 16; }

}  // namespace

// Passes.  A gradient pass holds at most 32768 geometries (1.18 M energies x 192 staging fields = 1.8 GB).  With
// PIMDK_CCPOL_STREAMS = N > 1 a call of more than one pass deals its passes round-robin to the caller's stream and
// N - 1 further ones, each with its own slice of the staging buffer: the block scheduler then fills the tail of one pass's
// kernels (and the idle slots of the low-occupancy stages) with CTAs of the other pass.  Passes are independent, so
// the results are the same bits either way.
#ifndef PIMDK_CCPOL_STREAMS
#define PIMDK_CCPOL_STREAMS 2
#endif
namespace {
#ifndef PIMDK_GRAD_PASS
#define PIMDK_GRAD_PASS 32768
#endif
constexpr long kGradPass = PIMDK_GRAD_PASS, kEnergyPass = 1048576;
constexpr int kPassStreams = PIMDK_CCPOL_STREAMS;
inline size_t geom_bytes(int grad) { return (size_t)kFields * 8 * (grad ? 36 : 1); }
// geometries per pass and number of staging halves in use for a buffer of work_bytes
inline long pass_geoms(long ngeom, int grad, size_t work_bytes, int* nbuf) {
  const long slots = (long)(work_bytes / geom_bytes(grad));
  *nbuf = 1;
  if (kPassStreams > 1 && grad && ngeom > kGradPass) {
    const long npass = (ngeom + kGradPass - 1) / kGradPass;
    const int want = npass < kPassStreams ? (int)npass : kPassStreams;
    if (slots >= want * kGradPass) {
      *nbuf = want;
      return kGradPass;
    }
  }
  return slots;
}
}  // namespace

// host copies of the two parameter blocks the kernels take by value; refreshed whenever the tables change
namespace {
SaptParams g_sapt;
RigidParams g_rigid;
}  // namespace
void KNAME(ccpol_host_tables)(const CcpolDev* h) { fill_params(*h, &g_sapt, &g_rigid); }

size_t KNAME(ccpol_work_bytes)(long ngeom, int grad) {
  const long cap = grad ? kGradPass : kEnergyPass;
  const long n = ngeom < cap ? ngeom : cap;
  long nbuf = 1;
  if (kPassStreams > 1 && grad && ngeom > cap) {
    nbuf = (ngeom + cap - 1) / cap;
    if (nbuf > kPassStreams) nbuf = kPassStreams;
  }
  return (size_t)(n < 1 ? 1 : n) * geom_bytes(grad) * (size_t)nbuf;
}

// kernel launches of one launch_ccpol call (7 per pass: setup, sites, dipind, sapt, rigid, sweep, combine)
long KNAME(ccpol_launches)(long ngeom, int grad, int icc, size_t work_bytes) {
  int nbuf;
  const long chunk = pass_geoms(ngeom, grad, work_bytes, &nbuf);
  if (chunk < 1 || ngeom < 1) return 0;
  return (icc ? 7 : 5) * ((ngeom + chunk - 1) / chunk);
}

cudaError_t KNAME(launch_ccpol)(const CcpolDev* tab, int iemonomer, int iembed, int icc, int potparts_old, double V0, GeomLayout L, double* x, double* v,
                                double* grad, long ngeom, int write_drift, int* flags, double* work, size_t work_bytes,
                                cudaStream_t st) {
  // function attributes are per device: one bit per device ordinal (a process may drive several GPUs, or re-initialise on another)
  static unsigned long long attr_mask = 0;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  const unsigned long long dev_bit = 1ull << (cur_dev & 63);
  if (!(attr_mask & dev_bit)) {
    cudaError_t e = cudaFuncSetAttribute(KNAME(ccpol_sapt_kernel)<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sapt_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(KNAME(ccpol_sapt_kernel)<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sapt_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(KNAME(ccpol_sweep_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem());
    if (e != cudaSuccess) return e;
    attr_mask |= dev_bit;
  }
  const int g = grad != nullptr;
  int nbuf;
  const long chunk = pass_geoms(ngeom, g, work_bytes, &nbuf);
  if (chunk < 1) return cudaErrorInvalidValue;
  static cudaStream_t st_x[kPassStreams > 1 ? kPassStreams : 2] = {};
  static cudaEvent_t ev_fork = nullptr, ev_join[kPassStreams > 1 ? kPassStreams : 2] = {};
  cudaStream_t const st_a = st;
  double* const work_a = work;
  static int st_dev = -1;                // device the extra streams live on (a process may re-initialise on another)
  if (nbuf > 1) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (ev_fork && dev != st_dev) ev_fork = nullptr;   // handles of the other device are left to its context
    if (!ev_fork) {
      st_dev = dev;
      cudaError_t e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
      for (int k = 1; k < kPassStreams && e == cudaSuccess; ++k) {
        e = cudaStreamCreateWithFlags(&st_x[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_join[k], cudaEventDisableTiming);
      }
      if (e != cudaSuccess) return e;
    }
    cudaEventRecord(ev_fork, st_a);      // the other streams start after everything queued before this call
    for (int k = 1; k < nbuf; ++k) cudaStreamWaitEvent(st_x[k], ev_fork, 0);
  }
  long ipass = 0;
  for (long g0 = 0; g0 < ngeom; g0 += chunk, ++ipass) {
    const long ng = (ngeom - g0 < chunk) ? ngeom - g0 : chunk;
    const long ne = ng * (g ? 36 : 1);
    const int slice = nbuf > 1 ? (int)(ipass % nbuf) : 0;
    st = slice ? st_x[slice] : st_a;
    work = work_a + (size_t)slice * kGradPass * (geom_bytes(1) / sizeof(double));
    KNAME(ccpol_setup_kernel)<<<(unsigned)((ne + kSetupBlock - 1) / kSetupBlock), kSetupBlock, 0, st>>>(iemonomer, iembed, L, x, g0, ne, g, work);
    KNAME(ccpol_sites_kernel)<<<(unsigned)((4 * ne + 127) / 128), 128, 0, st>>>(tab, ne, work);
    KNAME(ccpol_dipind_kernel)<<<(unsigned)((2 * ne + 127) / 128), 128, 0, st>>>(tab, ne, work);
    if (potparts_old) KNAME(ccpol_sapt_kernel)<true><<<(unsigned)((2 * ne + kSaptBlock - 1) / kSaptBlock), kSaptBlock, sapt_smem(), st>>>(PIMDK_SAPT_TABVAL, ne, work);
    else KNAME(ccpol_sapt_kernel)<false><<<(unsigned)((2 * ne + kSaptBlock - 1) / kSaptBlock), kSaptBlock, sapt_smem(), st>>>(PIMDK_SAPT_TABVAL, ne, work);
    if (icc) {   // CCpol-8s rigid model of the embedded monomers; surfaces 5..9 are SAPT-5s'f alone
      KNAME(ccpol_rigid_kernel)<<<(unsigned)((ne + kRigidBlock - 1) / kRigidBlock), kRigidBlock, 0, st>>>(g_rigid, ne, work, flags);
      KNAME(ccpol_sweep_kernel)<<<(unsigned)((ne + 31) / 32), kSweepBlock, sweep_smem(), st>>>(g_rigid, ne, work);
    }
    const long nt = g ? ne / 2 : ne;
    KNAME(ccpol_combine_kernel)<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(iemonomer, icc, V0, L, x, g0, ne, g, work, v, grad, write_drift, flags);
  }
  for (int k = 1; k < nbuf; ++k) {       // join: the caller's stream continues after the other streams' passes
    cudaEventRecord(ev_join[k], st_x[k]);
    cudaStreamWaitEvent(st_a, ev_join[k], 0);
  }
  return cudaGetLastError();
}

}  // namespace pimdk
