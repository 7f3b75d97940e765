// CCpol-8sf analytic energy + gradient (opt-in mode PIMDK_MODE_ANALYTIC), sm_100a: a four-kernel pipeline over a
// structure-of-arrays staging buffer [field][bead].  The mathematics is csrc/ccpol_grad.cuh (host/device; the same text is
// checked on the CPU against the oracle's dual-number gradient by tests/test_oracle.py).  One bead costs about two energies
// here against the 36 of the reference's central difference (mcmod_waterdimer_ccpol.f90:40-58).
//
//   prep   thread = (bead, monomer)  centre of mass, Radau frame (I, J, K = I x J), flexible SAPT-5s'f sites + s1..s3,
//                                    embedded-rigid sites (rigid body carried by the frame), PJT2 energy and gradient
//   sapt   thread = (bead, flexible | rigid geometry)  poten's 8 x 8 site pairs + dipind with hand-written adjoints
//                                    -> energy, d/d(sites of A, sites of B, s of A, s of B)
//   rigid  warp   = bead             CCpol-8s: lane = site; U0's 25 x 25 exponential pairs by a rotating schedule (lane a meets
//                                    site b = (a + t) mod 25 at step t: no two lanes touch one B site at once), damped
//                                    electrostatics (25 pairs, one per lane), dispersion (9 pairs), induction at its fixed
//                                    point; the site adjoints are reduced by shuffles to the 12 numbers of each monomer's frame
//   back   thread = (bead, atom)     three tangents through comcalc / radau_f1 / set_sites, contracted with the adjoints
//                                    -> dV/dx of that atom; atom 0 also writes V
// Compiled with contraction (this mode makes no operation-order promise).
#include "ccpol_grad.cuh"
#include "kernels.h"

namespace pimdk {
namespace {

using namespace agrad;

enum AF {
  AF_SITES = 0,     // 4 blocks x 27: flexible A, flexible B (24 site coordinates + s1..s3), rigid A, rigid B (24 used)
  AF_FRAME = 108,   // 2 x 12: I, J, K, centre of mass in bohr
  AF_PJG = 132,     // 18: d(monomer energy)/dA, kcal/mol/Angstrom
  AF_EMON = 150,    // 2: monomer energies, kcal/mol
  AF_VAL = 152,
  AF_VALL = 153,
  AF_ERIG = 154,    // kcal/mol
  AF_ADJF = 155,    // 54: adjoints of the flexible item (sites A, sites B, s of A, s of B), kcal/mol per Angstrom
  AF_ADJR = 209,    // 48: adjoints of the rigid item's sites
  AF_ADJFR = 257,   // 24: adjoints of the two frames from the CCpol-8s model, Hartree per unit of (I, J, K, COM in bohr)
  AF_N = 281
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- prep ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
agrad_prep_kernel(const CcpolDev* __restrict__ tab, const CcpolGradTab* __restrict__ gt, int iemonomer, GeomLayout L,
                  const double* __restrict__ x, long geom0, long nb, double* __restrict__ buf) {
  const long i = (long)blockIdx.x * 128 + threadIdx.x;
  if (i >= 2 * nb) return;
  const int m = i >= nb;
  const long e = m ? i - nb : i;
  const long base = L.base(geom0 + e);
  double A9[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) A9[d] = x[base + (long)(9 * m + d) * L.stride_dof] * kAngPlugin;
  double com[3], rel[9], I[3], J[3], K[3];
  comcalc_t<double>(A9, A9 + 3, A9 + 6, com);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) rel[a * 3 + j] = A9[a * 3 + j] - com[j];
  radau_f1_t<double>(rel, rel + 3, rel + 6, I, J);
  cross3(I, J, K);
  double* fr = buf + (long)(AF_FRAME + 12 * m) * nb + e;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    fr[(long)j * nb] = I[j];
    fr[(long)(3 + j) * nb] = J[j];
    fr[(long)(6 + j) * nb] = K[j];
    fr[(long)(9 + j) * nb] = com[j] / kA0;
  }
  {
    double c[3][3], sites[24], s[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j) c[a][j] = A9[a * 3 + j] / kA0;
    set_sites_t<double>(c, sites, s);
    double* out = buf + (long)(AF_SITES + 27 * m) * nb + e;
#pragma unroll
    for (int k = 0; k < 24; ++k) out[(long)k * nb] = sites[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) out[(long)(24 + k) * nb] = s[k];
  }
  {
    double* out = buf + (long)(AF_SITES + 27 * (2 + m)) * nb + e;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double a = __ldg(&gt->sapt_abc[k][0]), b = __ldg(&gt->sapt_abc[k][1]), c = __ldg(&gt->sapt_abc[k][2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) out[(long)(k * 3 + j) * nb] = com[j] + a * I[j] + b * J[j] + c * K[j];
    }
  }
  double g9[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, em = 0.0;
  if (iemonomer == 1) em = pjt2_monomer(A9, g9) * kHar2Kcal;
  buf[(long)(AF_EMON + m) * nb + e] = em;
#pragma unroll
  for (int k = 0; k < 9; ++k) buf[(long)(AF_PJG + 9 * m + k) * nb + e] = g9[k] * kHar2Kcal;
}

// ---- sapt ---------------------------------------------------------------------------------------
// sapt_item_adj (ccpol_grad.cuh: the host-checked statement of this stage) laid out for the SM: one 512-thread CTA per SM,
// the B sites and their adjoints in shared memory (slot-major, conflict-free), the A site of the current row and its adjoint
// in registers (the row's adjoint goes straight to the staging buffer), charges and their chain to s1..s3 recomputed per pair
// instead of being held in arrays — no local memory.
constexpr int kSaptThreads = 512;
template <int STRIDE>
struct Slots {
  double* p;   // &base[threadIdx.x]
  __device__ __forceinline__ double& operator[](int k) const { return p[k * STRIDE]; }
};
__global__ void __launch_bounds__(kSaptThreads, 1)
agrad_sapt_kernel(const __grid_constant__ SaptParams T, const CcpolGradTab* __restrict__ gt, long nb, double* __restrict__ buf) {
  // the SAPT-5s'f tables arrive as a kernel parameter: every read is warp-uniform, i.e. a constant-bank operand
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kSaptTab = 0;
  const long j = (long)blockIdx.x * kSaptThreads + threadIdx.x;
  if (j >= 2 * nb) return;
  const int which = j >= nb;      // 0 flexible geometry, 1 embedded-rigid geometry
  const long e = which ? j - nb : j;
  const double* pa = buf + (long)(AF_SITES + 54 * which) * nb + e;
  const double* pb = pa + (long)27 * nb;
  Slots<kSaptThreads> sitesB{reinterpret_cast<double*>(smem + kSaptTab) + threadIdx.x};
  Slots<kSaptThreads> adjB{reinterpret_cast<double*>(smem + kSaptTab) + 24 * kSaptThreads + threadIdx.x};
  double sA[3], sB[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sA[k] = which ? __ldg(&gt->s_rig[k]) : pa[(long)(24 + k) * nb];
    sB[k] = which ? __ldg(&gt->s_rig[k]) : pb[(long)(24 + k) * nb];
  }
  // ---- dipole induction: dipole sums and polarisabilities of both monomers, pair part with adjoints
  double dma[3] = {0.0, 0.0, 0.0}, dmb[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const double* prm = &T.param[site_type(i) * kNParam];
    const double sg = dipind_sign(i);
    const double qa = flex_charge(prm, sA[0], sA[1], sg * sA[2]) / 18.22262373;
    const double qb = flex_charge(prm, sB[0], sB[1], sg * sB[2]) / 18.22262373;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double b = pb[(long)(i * 3 + k) * nb];
      sitesB[i * 3 + k] = b;
      dma[k] += qa * pa[(long)(i * 3 + k) * nb] / kA0;
      dmb[k] += qb * b / kA0;
    }
  }
  double pA, pB;
  {
    const double* prm = &T.param[0];
    pA = prm[9] + prm[10] * sA[0] + prm[11] * sA[1] + prm[12] * sA[2] + prm[13] * sA[0] * sA[1] + prm[14] * sA[1] * sA[2] +
         prm[15] * sA[0] * sA[0] + prm[16] * sA[1] * sA[1] + prm[17] * sA[2] * sA[2];
    pB = prm[9] + prm[10] * sB[0] + prm[11] * sB[1] + prm[12] * sB[2] + prm[13] * sB[0] * sB[1] + prm[14] * sB[1] * sB[2] +
         prm[15] * sB[0] * sB[0] + prm[16] * sB[1] * sB[1] + prm[17] * sB[2] * sB[2];
  }
  double aOa[3], aOb[3], adma[3], admb[3], apA, apB;
  double E;
  {
    const double Oa[3] = {pa[0], pa[nb], pa[2 * nb]};
    const double Ob[3] = {sitesB[0], sitesB[1], sitesB[2]};
    E = dipind_pair_adj(T.parab[10 - 1], Oa, Ob, dma, dmb, pA, pB, aOa, aOb, adma, admb, apA, apB);
  }
  double asA[3], asB[3];    // adjoints of s1..s3 of A and of B
  {
    const double* prm = &T.param[0];
    asA[0] = apA * (prm[10] + prm[13] * sA[1] + 2.0 * prm[15] * sA[0]);
    asA[1] = apA * (prm[11] + prm[13] * sA[0] + prm[14] * sA[2] + 2.0 * prm[16] * sA[1]);
    asA[2] = apA * (prm[12] + prm[14] * sA[1] + 2.0 * prm[17] * sA[2]);
    asB[0] = apB * (prm[10] + prm[13] * sB[1] + 2.0 * prm[15] * sB[0]);
    asB[1] = apB * (prm[11] + prm[13] * sB[0] + prm[14] * sB[2] + 2.0 * prm[16] * sB[1]);
    asB[2] = apB * (prm[12] + prm[14] * sB[1] + 2.0 * prm[17] * sB[2]);
  }
  // B side of the dipole sums' adjoint: sites of B and, through the charges, s of B
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const double* prm = &T.param[site_type(i) * kNParam];
    const double sg = dipind_sign(i);
    const double q = flex_charge(prm, sB[0], sB[1], sg * sB[2]) / 18.22262373;
    double aq = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      adjB[i * 3 + k] = admb[k] * q / kA0 + (i == 0 ? aOb[k] : 0.0);
      aq += admb[k] * sitesB[i * 3 + k] / kA0;
    }
    double g[3] = {0.0, 0.0, 0.0};
    flex_charge_adj(prm, sB[0], sB[1], sg * sB[2], aq / 18.22262373, g);
    asB[0] += g[0];
    asB[1] += g[1];
    asB[2] += sg * g[2];
  }
  // ---- the 8 x 8 site pairs, row by row
  double* outA = buf + (long)(which ? AF_ADJR : AF_ADJF) * nb + e;
#pragma unroll 1
  for (int ia = 0; ia < 8; ++ia) {
    const double* prmA = &T.param[site_type(ia) * kNParam];
    const double ax = pa[(long)(ia * 3) * nb], ay = pa[(long)(ia * 3 + 1) * nb], az = pa[(long)(ia * 3 + 2) * nb];
    // the row's adjoint starts with the dipole-sum part
    const double sgd = dipind_sign(ia);
    const double qd = flex_charge(prmA, sA[0], sA[1], sgd * sA[2]) / 18.22262373;
    double gx = adma[0] * qd / kA0, gy = adma[1] * qd / kA0, gz = adma[2] * qd / kA0;
    if (ia == 0) { gx += aOa[0]; gy += aOa[1]; gz += aOa[2]; }
    {
      double g[3] = {0.0, 0.0, 0.0};
      flex_charge_adj(prmA, sA[0], sA[1], sgd * sA[2], (adma[0] * ax + adma[1] * ay + adma[2] * az) / kA0 / 18.22262373, g);
      asA[0] += g[0];
      asA[1] += g[1];
      asA[2] += sgd * g[2];
    }
    const double sga = (ia == 2) ? -1.0 : 1.0;
    const double qa = flex_charge(prmA, sA[0], sA[1], sga * sA[2]);
    double aqa = 0.0;
#pragma unroll 1
    for (int ib = 0; ib < 8; ++ib) {
      const double d0 = ax - sitesB[ib * 3], d1 = ay - sitesB[ib * 3 + 1], d2 = az - sitesB[ib * 3 + 2];
      const double r = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      const double* prmB = &T.param[site_type(ib) * kNParam];
      const double sgb = (ib == 2) ? -1.0 : 1.0;
      const double qb = flex_charge(prmB, sB[0], sB[1], sgb * sB[2]);
      PairOut o;
      sapt_pair_adj(T, ia, ib, r, sA, sB, qa, qb, o);
      E += o.e;
      const double f = o.dr / r;
      gx += f * d0; gy += f * d1; gz += f * d2;
      adjB[ib * 3] -= f * d0; adjB[ib * 3 + 1] -= f * d1; adjB[ib * 3 + 2] -= f * d2;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        asA[k] += o.dx[k];
        asB[k] += o.dy[k];
      }
      aqa += o.dqa;
      if (o.dqb != 0.0) {
        double g[3] = {0.0, 0.0, 0.0};
        flex_charge_adj(prmB, sB[0], sB[1], sgb * sB[2], o.dqb, g);
        asB[0] += g[0];
        asB[1] += g[1];
        asB[2] += sgb * g[2];
      }
    }
    {
      double g[3] = {0.0, 0.0, 0.0};
      flex_charge_adj(prmA, sA[0], sA[1], sga * sA[2], aqa, g);
      asA[0] += g[0];
      asA[1] += g[1];
      asA[2] += sga * g[2];
    }
    outA[(long)(ia * 3) * nb] = gx;
    outA[(long)(ia * 3 + 1) * nb] = gy;
    outA[(long)(ia * 3 + 2) * nb] = gz;
  }
  buf[(long)(which ? AF_VALL : AF_VAL) * nb + e] = E;
#pragma unroll 1
  for (int k = 0; k < 24; ++k) outA[(long)(24 + k) * nb] = adjB[k];
  if (!which) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      outA[(long)(48 + k) * nb] = asA[k];
      outA[(long)(51 + k) * nb] = asB[k];
    }
  }
}

// ---- rigid --------------------------------------------------------------------------------------
constexpr int kRigWarps = 8;
constexpr int kRigScratch = 75 + 75 + 75 + 75 + 27 + 30 + 30;   // sA, sB, aB, fE, fD, e0 parts, ind g parts (doubles per warp)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kRigWarps * 32, 2)
agrad_rigid_kernel(const CcpolDev* __restrict__ tab, const CcpolGradTab* __restrict__ gtab, long nb, double* __restrict__ buf,
                   int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kTabBytes = (int)PIMDK_RIGID_TABLE_BYTES;
  constexpr int kGtBytes = (int)((sizeof(CcpolGradTab) + 15) / 16 * 16);
  {
    const int4* s1 = reinterpret_cast<const int4*>(tab);
    int4* d1 = reinterpret_cast<int4*>(smem);
    for (int i = threadIdx.x; i < kTabBytes / 16; i += blockDim.x) d1[i] = s1[i];
    const int4* s2 = reinterpret_cast<const int4*>(gtab);
    int4* d2 = reinterpret_cast<int4*>(smem + kTabBytes);
    for (int i = threadIdx.x; i < kGtBytes / 16; i += blockDim.x) d2[i] = s2[i];
    __syncthreads();
  }
  const CcpolDev& T = *reinterpret_cast<const CcpolDev*>(smem);   // only the CCpol-8s members are valid
  const CcpolGradTab& G = *reinterpret_cast<const CcpolGradTab*>(smem + kTabBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = reinterpret_cast<double*>(smem + kTabBytes + kGtBytes) + warp * kRigScratch;
  double* sA = ws;            // [25][3]
  double* sB = ws + 75;
  double* aB = ws + 150;      // B-site adjoints of the sweep
  double* fE = ws + 225;      // [25][3] electrostatic pair forces
  double* fD = ws + 300;      // [9][3] dispersion pair forces
  double* e0p = ws + 327;     // [10][3] field contributions
  double* gp = ws + 357;      // [10][3] induction adjoint contributions
  const bool site = lane < 25;
  const int l = site ? lane : 0;
  const double ca = G.cc_abc[l][0], cb = G.cc_abc[l][1], cc = G.cc_abc[l][2];
  const double sig = 0.367911875040999981, plen = 1.1216873242;
  const double w0 = 1.0 - sig / plen, w12 = 0.5 * sig / plen;
  for (long e = (long)blockIdx.x * kRigWarps + warp; e < nb; e += (long)gridDim.x * kRigWarps) {
    // sites of both monomers from their frames
    double ra[3], rb[3];
    {
      const double* fa = buf + (long)AF_FRAME * nb + e;
      const double* fb = fa + (long)12 * nb;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        ra[j] = fa[(long)(9 + j) * nb] + ca * fa[(long)j * nb] + cb * fa[(long)(3 + j) * nb] + cc * fa[(long)(6 + j) * nb];
        rb[j] = fb[(long)(9 + j) * nb] + ca * fb[(long)j * nb] + cb * fb[(long)(3 + j) * nb] + cc * fb[(long)(6 + j) * nb];
      }
    }
    if (site) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        sA[l * 3 + j] = ra[j];
        sB[l * 3 + j] = rb[j];
        aB[l * 3 + j] = 0.0;
      }
    }
    __syncwarp();
    double E = 0.0, adA[3] = {0.0, 0.0, 0.0}, adB[3] = {0.0, 0.0, 0.0};
    // U0's exponential sweep: lane a meets B site (a + t) mod 25 at step t.  Five steps are evaluated together (independent
    // instruction streams: the stage is bound by dependent-issue latency), then their B-side updates are applied one step
    // at a time, a warp barrier between them, so that no two lanes ever update one B site at once.
#pragma unroll 1
    for (int t0 = 0; t0 < 25; t0 += 5) {
      double fx[5], fy[5], fz[5];
      int bb[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        int b = l + t0 + q;
        if (b >= 25) b -= 25;
        bb[q] = b;
        const double d0 = ra[0] - sB[b * 3], d1 = ra[1] - sB[b * 3 + 1], d2 = ra[2] - sB[b * 3 + 2];
        const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        const double ri = rsqrt(r2);
        const double R = r2 * ri;
        double pe, de;
        sweep_pair(G.bin5[G.pair_bin[b * 25 + l]], R, pe, de);
        const double f = site ? de * ri : 0.0;
        E += site ? pe : 0.0;
        fx[q] = f * d0; fy[q] = f * d1; fz[q] = f * d2;
        adA[0] += fx[q]; adA[1] += fy[q]; adA[2] += fz[q];
      }
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        if (site) {
          aB[bb[q] * 3] -= fx[q]; aB[bb[q] * 3 + 1] -= fy[q]; aB[bb[q] * 3 + 2] -= fz[q];
        }
        __syncwarp();
      }
    }
    // damped electrostatics: pair (a, b) = (lane / 5, lane % 5); dispersion: (lane / 3, lane % 3)
    if (site) {
      const int a = l / 5, b = l - 5 * a;
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if ((int)T.ind_charge[a] * (int)T.ind_charge[b] != 0) {
        const double d0 = sA[a * 3] - sB[b * 3], d1 = sA[a * 3 + 1] - sB[b * 3 + 1], d2 = sA[a * 3 + 2] - sB[b * 3 + 2];
        const double R = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        double pe, de;
        elst_pair(T.params[T.ind_d1[b * 5 + a] - 1], T.params[T.ind_charge[a] - 1] * T.params[T.ind_charge[b] - 1], R, pe, de);
        E += pe;
        const double f = de / R;
        f0 = f * d0; f1 = f * d1; f2 = f * d2;
      }
      fE[l * 3] = f0; fE[l * 3 + 1] = f1; fE[l * 3 + 2] = f2;
    }
    if (lane < 9) {
      const int a = lane / 3, b = lane - 3 * a, q = b * 3 + a;
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if (T.ind_d6[q] != 0) {
        const double d0 = sA[a * 3] - sB[b * 3], d1 = sA[a * 3 + 1] - sB[b * 3 + 1], d2 = sA[a * 3 + 2] - sB[b * 3 + 2];
        const double R = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double dm[3] = {T.params[T.ind_d6[q] - 1], T.params[T.ind_d8[q] - 1], T.params[T.ind_d10[q] - 1]};
        const double c3[3] = {T.params[T.ind_c6[q] - 1], T.params[T.ind_c8[q] - 1], T.params[T.ind_c10[q] - 1]};
        double pe, de;
        disp_pair(dm, c3, R, pe, de);
        E += pe;
        const double f = de / R;
        f0 = f * d0; f1 = f * d1; f2 = f * d2;
      }
      fD[lane * 3] = f0; fD[lane * 3 + 1] = f1; fD[lane * 3 + 2] = f2;
    }
    // induction: polarisable centres, permanent fields (10 contributions, one per lane)
    double Rp[2][3], E0[2][3], mu[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Rp[0][j] = sA[j] + sig * (0.5 * (sA[3 + j] + sA[6 + j]) - sA[j]) / plen;
      Rp[1][j] = sB[j] + sig * (0.5 * (sB[3 + j] + sB[6 + j]) - sB[j]) / plen;
    }
    if (lane < 10) {
      const int i = lane / 5, q = lane - 5 * i;
      const double* s = i ? sA : sB;             // the OTHER monomer's charged site q
      const double d0 = (i ? Rp[1][0] : Rp[0][0]) - s[q * 3], d1 = (i ? Rp[1][1] : Rp[0][1]) - s[q * 3 + 1],
                   d2 = (i ? Rp[1][2] : Rp[0][2]) - s[q * 3 + 2];
      const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
      const double w = T.chrg[q] / (r2 * sqrt(r2));
      e0p[lane * 3] = w * d0; e0p[lane * 3 + 1] = w * d1; e0p[lane * 3 + 2] = w * d2;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      E0[0][j] = e0p[j] + e0p[3 + j] + e0p[6 + j] + e0p[9 + j] + e0p[12 + j];
      E0[1][j] = e0p[15 + j] + e0p[18 + j] + e0p[21 + j] + e0p[24 + j] + e0p[27 + j];
    }
    int noconv = 0;
    const double Eind = ind_solve(Rp, E0, mu, &noconv);    // every lane runs the same iteration
    if (noconv && lane == 0) atomicOr(flags, PIMDK_FLAG_NOCONV);
    double aE0[2][3], aV[3];
    ind_adj(Rp, mu, aE0, aV);
    if (lane < 10) {
      const int i = lane / 5, q = lane - 5 * i;
      const double* s = i ? sA : sB;
      const double d0 = (i ? Rp[1][0] : Rp[0][0]) - s[q * 3], d1 = (i ? Rp[1][1] : Rp[0][1]) - s[q * 3 + 1],
                   d2 = (i ? Rp[1][2] : Rp[0][2]) - s[q * 3 + 2];
      const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
      const double r3i = 1.0 / (r2 * sqrt(r2)), r5i = r3i / r2;
      const double a0 = i ? aE0[1][0] : aE0[0][0], a1 = i ? aE0[1][1] : aE0[0][1], a2 = i ? aE0[1][2] : aE0[0][2];
      const double ad = a0 * d0 + a1 * d1 + a2 * d2;
      const double cq = T.chrg[q];
      gp[lane * 3] = cq * (a0 * r3i - 3.0 * d0 * ad * r5i);
      gp[lane * 3 + 1] = cq * (a1 * r3i - 3.0 * d1 * ad * r5i);
      gp[lane * 3 + 2] = cq * (a2 * r3i - 3.0 * d2 * ad * r5i);
    }
    __syncwarp();
    if (site) {
      // gather: sweep (B side), electrostatics, dispersion, induction
#pragma unroll
      for (int j = 0; j < 3; ++j) adB[j] += aB[l * 3 + j];
      if (l < 5) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          adA[j] += fE[(l * 5) * 3 + j] + fE[(l * 5 + 1) * 3 + j] + fE[(l * 5 + 2) * 3 + j] + fE[(l * 5 + 3) * 3 + j] + fE[(l * 5 + 4) * 3 + j];
          adB[j] -= fE[l * 3 + j] + fE[(5 + l) * 3 + j] + fE[(10 + l) * 3 + j] + fE[(15 + l) * 3 + j] + fE[(20 + l) * 3 + j];
          adB[j] -= gp[l * 3 + j];          // field at A's centre from B's site l
          adA[j] -= gp[(5 + l) * 3 + j];    // field at B's centre from A's site l
        }
      }
      if (l < 3) {
        const double w = l == 0 ? w0 : w12;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          adA[j] += fD[(l * 3) * 3 + j] + fD[(l * 3 + 1) * 3 + j] + fD[(l * 3 + 2) * 3 + j];
          adB[j] -= fD[l * 3 + j] + fD[(3 + l) * 3 + j] + fD[(6 + l) * 3 + j];
          const double aRpA = aV[j] + gp[j] + gp[3 + j] + gp[6 + j] + gp[9 + j] + gp[12 + j];
          const double aRpB = -aV[j] + gp[15 + j] + gp[18 + j] + gp[21 + j] + gp[24 + j] + gp[27 + j];
          adA[j] += w * aRpA;
          adB[j] += w * aRpB;
        }
      }
    }
    // reduce the site adjoints to the frames (site = COM + a I + b J + c K): the 2 x 25 x 3 site adjoints go through shared
    // memory and 24 lanes form one of the 24 frame adjoints each (25 multiply-adds) instead of 24 shuffle reductions
    const double Esum = warp_sum(E);
    __syncwarp();
    if (site) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        sA[l * 3 + j] = adA[j];     // the sites themselves are no longer needed
        sB[l * 3 + j] = adB[j];
      }
    }
    __syncwarp();
    if (lane < 24) {
      const int m = lane / 12, r = lane - 12 * m, w = r / 3, j = r - 3 * w;   // w: 0 I, 1 J, 2 K, 3 COM
      const double* ad = m ? sB : sA;
      double acc = 0.0;
#pragma unroll 5
      for (int k = 0; k < 25; ++k) acc += (w < 3 ? G.cc_abc[k][w] : 1.0) * ad[k * 3 + j];
      buf[(long)(AF_ADJFR + lane) * nb + e] = acc;
    }
    if (lane == 0) buf[(long)AF_ERIG * nb + e] = (Esum + Eind) * kHar2Kcal;
    __syncwarp();
  }
}

// ---- back ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
agrad_back_kernel(const CcpolGradTab* __restrict__ gt, int iemonomer, int icc, double V0, GeomLayout L, const double* __restrict__ x,
                  long geom0, long nb, const double* __restrict__ buf, double* __restrict__ v, double* __restrict__ gradout,
                  int* __restrict__ flags) {
  typedef Dn<3> D;
  const long i = (long)blockIdx.x * 128 + threadIdx.x;
  if (i >= 6 * nb) return;
  const int atom6 = (int)(i / nb);
  const long e = i - (long)atom6 * nb;
  const int m = atom6 / 3, atom = atom6 - 3 * m;
  const long base = L.base(geom0 + e);
  D X[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) X[d] = D(x[base + (long)(9 * m + d) * L.stride_dof] * kAngPlugin);
#pragma unroll
  for (int t = 0; t < 3; ++t) X[atom * 3 + t].d[t] = 1.0;
  D com[3], rel[9], I[3], J[3], K[3];
  comcalc_t<D>(X, X + 3, X + 6, com);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) rel[a * 3 + j] = X[a * 3 + j] - com[j];
  radau_f1_t<D>(rel, rel + 3, rel + 6, I, J);
  K[0] = I[1] * J[2] - I[2] * J[1];
  K[1] = I[2] * J[0] - I[0] * J[2];
  K[2] = I[0] * J[1] - I[1] * J[0];
  // adjoints of the rigid body (I, J, K, COM in Angstrom): embedded-rigid SAPT-5s'f sites (Etot = Erigid + val - vall) and
  // the CCpol-8s frame adjoints (Hartree, COM in bohr)
  double aI[3] = {0.0, 0.0, 0.0}, aJ[3] = {0.0, 0.0, 0.0}, aK[3] = {0.0, 0.0, 0.0}, aC[3] = {0.0, 0.0, 0.0};
  if (icc) {
    const double* ar = buf + (long)(AF_ADJR + 24 * m) * nb + e;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double a = __ldg(&gt->sapt_abc[k][0]), b = __ldg(&gt->sapt_abc[k][1]), c = __ldg(&gt->sapt_abc[k][2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double w = -ar[(long)(k * 3 + j) * nb];
        aI[j] += a * w; aJ[j] += b * w; aK[j] += c * w; aC[j] += w;
      }
    }
    const double* af = buf + (long)(AF_ADJFR + 12 * m) * nb + e;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      aI[j] += kHar2Kcal * af[(long)j * nb];
      aJ[j] += kHar2Kcal * af[(long)(3 + j) * nb];
      aK[j] += kHar2Kcal * af[(long)(6 + j) * nb];
      aC[j] += kHar2Kcal / kA0 * af[(long)(9 + j) * nb];
    }
  }
  double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int t = 0; t < 3; ++t) g[t] += aI[k] * I[k].d[t] + aJ[k] * J[k].d[t] + aK[k] * K[k].d[t] + aC[k] * com[k].d[t];
  {
    D c[3][3], sites[24], s[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j) c[a][j] = X[a * 3 + j] / kA0;
    set_sites_t<D>(c, sites, s);
    const double* as = buf + (long)(AF_ADJF + 24 * m) * nb + e;
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      const double w = as[(long)k * nb];
#pragma unroll
      for (int t = 0; t < 3; ++t) g[t] += w * sites[k].d[t];
    }
    const double* ass = buf + (long)(AF_ADJF + 48 + 3 * m) * nb + e;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double w = ass[(long)k * nb];
#pragma unroll
      for (int t = 0; t < 3; ++t) g[t] += w * s[k].d[t];
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const double gt_ = (g[t] + buf[(long)(AF_PJG + 9 * m + atom * 3 + t) * nb + e]) * (kAngPlugin / kHar2Kcal);
    if (gradout) gradout[base + (long)(atom6 * 3 + t) * L.stride_dof] = gt_;
    if (gt_ != gt_) atomicOr(flags, PIMDK_FLAG_NAN);
  }
  if (v && atom6 == 0) {
    const double val = buf[(long)AF_VAL * nb + e];
    double Etot = icc ? buf[(long)AF_ERIG * nb + e] + (val - buf[(long)AF_VALL * nb + e]) : val;
    if (iemonomer == 1) Etot += buf[(long)AF_EMON * nb + e] + buf[(long)(AF_EMON + 1) * nb + e];
    v[geom0 + e] = Etot / kHar2Kcal - V0;
  }
}

size_t sapt_smem() { return (size_t)48 * kSaptThreads * sizeof(double); }
SaptParams g_sapt;   // host copy of the SAPT-5s'f tables, passed to agrad_sapt_kernel by value
size_t rigid_smem() {
  return (size_t)PIMDK_RIGID_TABLE_BYTES + (sizeof(CcpolGradTab) + 15) / 16 * 16 + (size_t)kRigWarps * kRigScratch * sizeof(double);
}

}  // namespace

size_t ccpol_analytic_bytes_per_geom() { return (size_t)AF_N * sizeof(double); }
long ccpol_analytic_launches(long ngeom, int icc, size_t work_bytes) {
  const long cap = (long)(work_bytes / ccpol_analytic_bytes_per_geom());
  if (cap < 1 || ngeom < 1) return 0;
  return (icc ? 4 : 3) * ((ngeom + cap - 1) / cap);
}

// v and/or grad for ngeom geometries; `work` holds at least one geometry's fields (a pass takes as many as fit)
void ccpol_host_tables_analytic(const CcpolDev* h) {
  RigidParams unused;
  fill_params(*h, &g_sapt, &unused);
}

cudaError_t launch_ccpol_analytic(const CcpolDev* tab, const CcpolGradTab* gt, int iemonomer, int icc, double V0, GeomLayout L,
                                  const double* x, double* v, double* grad, long ngeom, int* flags, double* work, size_t work_bytes,
                                  int num_sms, cudaStream_t st) {
  static unsigned long long attr_mask = 0;   // per-device opt-in of the dynamic shared memory sizes
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_mask & (1ull << (dev & 63)))) {
    cudaError_t e = cudaFuncSetAttribute(agrad_sapt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sapt_smem());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(agrad_rigid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rigid_smem());
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << (dev & 63);
  }
  const long cap = (long)(work_bytes / ccpol_analytic_bytes_per_geom());
  if (cap < 1) return cudaErrorInvalidValue;
  for (long g0 = 0; g0 < ngeom; g0 += cap) {
    const long nb = ngeom - g0 < cap ? ngeom - g0 : cap;
    agrad_prep_kernel<<<(unsigned)((2 * nb + 127) / 128), 128, 0, st>>>(tab, gt, iemonomer, L, x, g0, nb, work);
    agrad_sapt_kernel<<<(unsigned)((2 * nb + kSaptThreads - 1) / kSaptThreads), kSaptThreads, sapt_smem(), st>>>(g_sapt, gt, nb, work);
    if (icc) {
      long blocks = (nb + kRigWarps - 1) / kRigWarps;
      const long capb = (long)num_sms * 16;
      if (blocks > capb) blocks = capb;
      agrad_rigid_kernel<<<(unsigned)blocks, kRigWarps * 32, rigid_smem(), st>>>(tab, gt, nb, work, flags);
    }
    agrad_back_kernel<<<(unsigned)((6 * nb + 127) / 128), 128, 0, st>>>(gt, iemonomer, icc, V0, L, x, g0, nb, work, v, grad, flags);
  }
  return cudaGetLastError();
}

}  // namespace pimdk
