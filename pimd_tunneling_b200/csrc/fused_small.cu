// Fused warp-per-ring-polymer propagation for small systems (1D / 2D model surfaces, n <= 128 beads):
// propagate_pimd_pile / propagate_pimd_nm (verletmodule.f90:190-250, 372-416) for all NMC steps in ONE
// persistent kernel.  One warp owns one ring polymer for its whole trajectory:
//   * lane l holds beads / normal modes k = l, l+32, ... of every dof in registers (P, Q, x);
//   * the bead <-> normal-mode transform (nmtransform_*, :254-286) is a warp-local mat-vec against the
//     transmatrix held in shared memory (32 KB at n = 64): the vector is staged through a warp-private
//     shared-memory line and broadcast back, each lane accumulating its own outputs with fma in ascending
//     j — the same summation order as the tiled GEMM of nm_kernels.cu, so both paths agree bit for bit;
//   * kick, free ring-polymer rotation, PILE O-step / Andersen resampling and the estimator use the same
//     device functions as the streamed kernels (nm_device.cuh, pes_simple_device.cuh).
// The streamed path needs 5-8 kernel launches per step; for C1 (64 beads x 256 trajectories) a step is
// ~1 us of arithmetic, so it was launch-bound by two orders of magnitude.
#include "kernels.h"
#include "nm_device.cuh"
#include "pes_simple_device.cuh"
#include "philox.cuh"

namespace pimdk {
namespace {

constexpr int kWarpsPerBlock = 4;

template <int NDOF, int S>
struct PolymerRegs {
  double P[NDOF][S], Q[NDOF][S], x[NDOF][S];
};

// y[k] = sum_j T[j][k] v[j] for the lane's k = lane + 32 s; v is read from the warp's staging line
template <int S>
__device__ __forceinline__ void warp_matvec(const double* __restrict__ Ts, const double* __restrict__ vec, int n,
                                            int lane, double* y) {
#pragma unroll
  for (int s = 0; s < S; ++s) y[s] = 0.0;
  // the chain is 8 cycles per term; unrolled so that the shared-memory loads of the next terms are in flight behind it
  // (not unrolled, every term waited for its own loads: ~40 cycles each, two thirds of a C1 step)
#pragma unroll 8
  for (int j = 0; j < n; ++j) {
    const double v = vec[j];
    const double* row = Ts + (long)j * n + lane;
#pragma unroll
    for (int s = 0; s < S; ++s)
      if (lane + 32 * s < n) y[s] = fma(v, row[32 * s], y[s]);
  }
}

template <int NDOF, int S>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
fused_small_kernel(NmTables nm, int pes_kind, SimplePesParams pp, int thermostat, long ntraj, double* __restrict__ xg,
                   double* __restrict__ pg, const double* __restrict__ a, const double* __restrict__ b,
                   const double* __restrict__ dbdl, double dt, long NMC, long imin, double lambda, uint64_t seed,
                   const int64_t* __restrict__ gid, double* __restrict__ dHdr, int* __restrict__ flags, long step0,
                   int keep_sum, double* __restrict__ dHsum, int* __restrict__ clk_count, int* __restrict__ clk_kick,
                   int carry, double dhdrlimit, int npath, const double* __restrict__ lampath, const double* __restrict__ rpath,
                   const double* __restrict__ rspl, const double* __restrict__ xis) {
  extern __shared__ __align__(16) double sm[];
  const int n = nm.n;
  double* Ts = sm;                                   // transmatrix, n x n
  double* vecs = Ts + (long)n * n;                   // one staging line of n doubles per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long i = threadIdx.x; i < (long)n * n; i += blockDim.x) Ts[i] = nm.T[i];
  __syncthreads();
  double* vec = vecs + warp * n;
  const long traj = (long)blockIdx.x * kWarpsPerBlock + warp;
  if (traj >= ntraj) return;
  const uint32_t g = gid ? (uint32_t)gid[traj] : (uint32_t)traj;
  const int ndim = nm.ndim;
  double* xt = xg + traj * (long)NDOF * n;
  double* pt = pg + traj * (long)NDOF * n;

  PolymerRegs<NDOF, S> r;
  // to normal-mode space: P = T p, Q = T x - beadvec (nmtransform_forward)
#pragma unroll
  for (int d = 0; d < NDOF; ++d) {
    for (int k = lane; k < n; k += 32) vec[k] = pt[(long)d * n + k];
    __syncwarp();
    warp_matvec<S>(Ts, vec, n, lane, r.P[d]);
    __syncwarp();
    for (int k = lane; k < n; k += 32) vec[k] = xt[(long)d * n + k];
    __syncwarp();
    warp_matvec<S>(Ts, vec, n, lane, r.Q[d]);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int k = lane + 32 * s;
      if (k < n) {
        r.Q[d][s] = r.Q[d][s] - beadvec_at(nm, a, b, traj, d, k);
        r.x[d][s] = xt[(long)d * n + k];
      }
    }
  }

  auto to_beads = [&]() {  // x = T (Q + beadvec)   (nmtransform_backward)
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int k = lane + 32 * s;
        if (k < n) vec[k] = r.Q[d][s] + beadvec_at(nm, a, b, traj, d, k);
      }
      __syncwarp();
      warp_matvec<S>(Ts, vec, n, lane, r.x[d]);
      __syncwarp();
    }
  };
  auto kick_and_rotate = [&](bool langevin, int nrot, uint64_t step) {  // step_v kick + step_nm rotation(s) [+ O-step]
    double gb[S][NDOF];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int k = lane + 32 * s;
      double xb[NDOF], e;
#pragma unroll
      for (int d = 0; d < NDOF; ++d) xb[d] = r.x[d][s];
      if (k < n) simple_pes_eval<NDOF>(pes_kind, pp, xb, &e, gb[s], false, true);  // Vprime(x(k,:,:))
    }
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int k = lane + 32 * s;
        if (k < n) vec[k] = gb[s][d];
      }
      __syncwarp();
      double G[S];
      warp_matvec<S>(Ts, vec, n, lane, G);  // G = T g
      __syncwarp();
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int k = lane + 32 * s;
        if (k >= n) continue;
        const int ak = (d / ndim) * n + k;
        double P = r.P[d][s], Q = r.Q[d][s];
        P = P - G[s] * dt;
        rotate(nm, ak, P, Q);
        if (langevin) {
          const double xi = normal_at(seed, STREAM_LANGEVIN, step, g, (uint64_t)d * n + k);
          P = nm.c1sq[ak] * P + nm.cnoise[ak] * xi;
        }
        if (nrot >= 2) rotate(nm, ak, P, Q);
        if (P != P) atomicOr(flags, PIMDK_FLAG_NAN);
        r.P[d][s] = P;
        r.Q[d][s] = Q;
      }
    }
  };
  // estimator (verletmodule.f90:397-403): the lane that owns the last bead
  const int last_lane = (n - 1) & 31, last_s = (n - 1) >> 5;
  double acc = keep_sum ? dHdr[traj] : 0.0;   // restart = 2 continues the running sum (verletmodule.f90:200,388)
  // estimator; returns true (warp-uniform) when the dHdrlimit guard drops the contribution (verletmodule.f90:404-409)
  auto estimator = [&](bool guard) -> bool {
    double contr = 0.0;
    if (lane == last_lane) {
    for (int j = 0; j < ndim; ++j)
      for (int k = 0; k < NDOF / ndim; ++k) {
        const int d = k * ndim + j;
        double xl = 0.0;
#pragma unroll
        for (int dd = 0; dd < NDOF; ++dd)
#pragma unroll
          for (int s = 0; s < S; ++s)
            if (dd == d && s == last_s) xl = r.x[dd][s];
        contr = contr + nm.mass[k] * (-xl) * dbdl[traj * NDOF + d];
      }
    }
    bool over = false;
    if (guard) {
      const double c = __shfl_sync(0xffffffffu, contr, last_lane);
      over = !(fabs(c) < dhdrlimit || dhdrlimit < 0.0);
    }
    if (lane == last_lane && !over) acc = acc + contr;
    return over;
  };
  // init_path again (verletmodule.f90:39-48, 102-115): beads on the spline, momenta drawn in normal-mode space (stream 0 at
  // the current step), Q = T x - beadvec
  auto reinit = [&](uint64_t step) {
#pragma unroll
    for (int d = 0; d < NDOF; ++d) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int k = lane + 32 * s;
        if (k < n) {
          const double xv = (double)k * xis[traj] / (double)(n - 1);
          r.x[d][s] = splint_at(lampath, rpath + (long)d * npath, rspl + (long)d * npath, npath, xv);
          const double z = normal_at(seed, STREAM_INIT, step, g, (uint64_t)d * n + k);
          r.P[d][s] = (0.0 + nm.stdev * z) * nm.sigp[(d / ndim) * n + k];
          vec[k] = r.x[d][s];
        }
      }
      __syncwarp();
      warp_matvec<S>(Ts, vec, n, lane, r.Q[d]);
      __syncwarp();
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int k = lane + 32 * s;
        if (k < n) r.Q[d][s] = r.Q[d][s] - beadvec_at(nm, a, b, traj, d, k);
      }
    }
  };

  if (thermostat == 2) {  // time_step_pile (:423-435)
    for (long ii = 1; ii <= NMC; ++ii) {
      kick_and_rotate(true, 2, (uint64_t)(ii + step0));
      to_beads();
      if (ii > imin && estimator(dhdrlimit >= 0.0)) reinit((uint64_t)(ii + step0));
    }
  } else {                // propagate_pimd_nm (:190-250) / time_step_nm (:291-302)
    // collision clock: fresh (:199-202) or continued from the previous call of a run cut into segments
    int count = carry ? clk_count[traj] : 0;
    int rkick = carry ? clk_kick[traj] : poisson_norm(seed, (uint64_t)step0, g, lambda);
    for (long ii = 1; ii <= NMC; ++ii) {
      count = count + 1;
      if (count >= rkick) {
        count = 0;
#pragma unroll
        for (int d = 0; d < NDOF; ++d)
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int k = lane + 32 * s;
            if (k < n) {
              const double z = normal_at(seed, STREAM_ANDERSEN, (uint64_t)(ii + step0), g, (uint64_t)d * n + k);
              r.P[d][s] = (0.0 + nm.stdev * z) * nm.sigp[(d / ndim) * n + k];
            }
          }
        rkick = poisson_norm(seed, (uint64_t)(ii + step0), g, lambda);
      }
#pragma unroll
      for (int d = 0; d < NDOF; ++d)
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int k = lane + 32 * s;
          if (k < n) rotate(nm, (d / ndim) * n + k, r.P[d][s], r.Q[d][s]);
        }
      to_beads();
      kick_and_rotate(false, 1, (uint64_t)(ii + step0));
      to_beads();
      if (ii > imin) estimator(false);   // propagate_pimd_nm has no dHdrlimit guard (verletmodule.f90:236-244)
    }
    if (lane == 0 && clk_count) {
      clk_count[traj] = count;
      clk_kick[traj] = rkick;
    }
  }
  // back to bead space: p = T P ; x is current
#pragma unroll
  for (int d = 0; d < NDOF; ++d) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int k = lane + 32 * s;
      if (k < n) vec[k] = r.P[d][s];
    }
    __syncwarp();
    double pb[S];
    warp_matvec<S>(Ts, vec, n, lane, pb);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int k = lane + 32 * s;
      if (k < n) {
        pt[(long)d * n + k] = pb[s];
        xt[(long)d * n + k] = r.x[d][s];
      }
    }
  }
  if (lane == last_lane) {   // running sum (what write_restart stores, :171) and mean (:247,413)
    dHsum[traj] = acc;
    dHdr[traj] = acc / (double)(NMC + step0 - imin);
  }
}

template <int NDOF, int S>
cudaError_t launch_t(const NmTables& nm, int kind, const SimplePesParams& pp, int thermostat, long ntraj, double* x,
                     double* p, const double* a, const double* b, const double* dbdl, double dt, long NMC, long imin,
                     double lambda, uint64_t seed, const int64_t* gid, double* dHdr, int* flags, long step0, int keep_sum,
                     double* dHsum, int* clk_count, int* clk_kick, int carry, double dhdrlimit, int npath, const double* lampath,
                     const double* rpath, const double* rspl, const double* xis, cudaStream_t st) {
  const size_t smem = ((size_t)nm.n * nm.n + (size_t)kWarpsPerBlock * nm.n) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(fused_small_kernel<NDOF, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((ntraj + kWarpsPerBlock - 1) / kWarpsPerBlock);
  fused_small_kernel<NDOF, S><<<blocks, kWarpsPerBlock * 32, smem, st>>>(nm, kind, pp, thermostat, ntraj, x, p, a, b, dbdl,
                                                                        dt, NMC, imin, lambda, seed, gid, dHdr, flags, step0,
                                                                        keep_sum, dHsum, clk_count, clk_kick, carry, dhdrlimit, npath,
                                                                        lampath, rpath, rspl, xis);
  return cudaGetLastError();
}

}  // namespace

bool fused_small_supported(PesKind kind, int n, int ndim, int natom) {
  const int ndof = ndim * natom;
  if (kind != PES_1D && kind != PES_2DTEST && kind != PES_SO2) return false;
  if ((kind == PES_2DTEST || kind == PES_SO2) && ndof != 2) return false;
  return n >= 2 && n <= 128 && (ndof == 1 || ndof == 2);
}

cudaError_t launch_fused_small(const NmTables& nm, PesKind kind, const SimplePesParams& pp, int thermostat, long ntraj,
                               double* x, double* p, const double* a, const double* b, const double* dbdl, double dt,
                               long NMC, long imin, double lambda, uint64_t seed, const int64_t* gid, double* dHdr,
                               int* flags, long step0, int keep_sum, double* dHsum, int* clk_count, int* clk_kick, int carry,
                               double dhdrlimit, int npath, const double* lampath, const double* rpath, const double* rspl,
                               const double* xis, cudaStream_t st) {
  const int S = (nm.n + 31) / 32;
#define PIMDK_FUSED_CASE(ND, SS)                                                                                        \
  if (nm.ndof == ND && S == SS)                                                                                         \
    return launch_t<ND, SS>(nm, (int)kind, pp, thermostat, ntraj, x, p, a, b, dbdl, dt, NMC, imin, lambda, seed, gid, \
                            dHdr, flags, step0, keep_sum, dHsum, clk_count, clk_kick, carry, dhdrlimit, npath, lampath, rpath, rspl,   \
                            xis, st);
  PIMDK_FUSED_CASE(1, 1) PIMDK_FUSED_CASE(1, 2) PIMDK_FUSED_CASE(1, 3) PIMDK_FUSED_CASE(1, 4)
  PIMDK_FUSED_CASE(2, 1) PIMDK_FUSED_CASE(2, 2) PIMDK_FUSED_CASE(2, 3) PIMDK_FUSED_CASE(2, 4)
#undef PIMDK_FUSED_CASE
  return cudaErrorInvalidValue;
}

}  // namespace pimdk
