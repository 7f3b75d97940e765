// Normal-mode machinery of `module verletint` as batched sm_100a kernels:
//   nmtransform_forward/backward (verletmodule.f90:254-286; MKL dsymv per (dim,atom) vector)
//        -> one FP64 tile GEMM  Y[(traj,dof), :] = f(A)[(traj,dof), :] * T  over all ring polymers,
//           with the beadvec shift (init_nm :328-333) fused as prologue/epilogue;
//   step_nm rotation (:515-539), step_v kick (:576), step_langevin O-step (:651-652)
//        -> one elementwise kernel in normal-mode space (Philox noise generated in registers);
//   Andersen resampling (:208-234), init_path momenta (:102-115), estimator (:397-403).
// State layout on the device is the reference's x(n,ndim,natom,ntraj): bead/mode index fastest,
// so a (traj,dof) row is contiguous and the transform is a row-major GEMM against the symmetric T.
#include "kernels.h"
#include "nm_device.cuh"
#include "pes_simple_device.cuh"
#include "philox.cuh"

namespace pimdk {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

template <int MODE>
__global__ void __launch_bounds__(GT)
nm_gemm_kernel(NmTables nm, const double* __restrict__ A, double* __restrict__ Y, long rows,
               const double* __restrict__ a, const double* __restrict__ b) {
  __shared__ __align__(16) double As[BK][BM];
  __shared__ __align__(16) double Bs[BK][BN];
  const int n = nm.n;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long row0 = (long)blockIdx.y * BM;
  const int col0 = blockIdx.x * BN;

  // global->register staging assignments
  const int a_r = tid >> 1, a_k0 = (tid & 1) * 8;   // A tile: row a_r, 8 consecutive k
  const int b_k = tid >> 4, b_c0 = (tid & 15) * 8;  // T tile: row b_k, 8 consecutive columns
  const long a_row = row0 + a_r;
  const bool a_ok = a_row < rows;
  long a_traj = 0;
  int a_dof = 0;
  if (MODE == GEMM_ADD_BEADVEC && a_ok) {
    a_traj = a_row / nm.ndof;
    a_dof = (int)(a_row - a_traj * nm.ndof);
  }
  double ra[8], rb[8];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = k0 + a_k0 + q;
      double v = 0.0;
      if (a_ok && j < n) {
        v = A[a_row * n + j];
        if (MODE == GEMM_ADD_BEADVEC) v = v + beadvec_at(nm, a, b, a_traj, a_dof, j);
      }
      ra[q] = v;
    }
    const int j = k0 + b_k;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = col0 + b_c0 + q;
      rb[q] = (j < n && c < n) ? nm.T[(long)j * n + c] : 0.0;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int q = 0; q < 8; ++q) As[a_k0 + q][a_r] = ra[q];
#pragma unroll
    for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(&Bs[b_k][b_c0 + q]) = make_double2(rb[q], rb[q + 1]);
  };

  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  load_tiles(0);
  for (int k0 = 0; k0 < n; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < n) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double af[8], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 t = *reinterpret_cast<const double2*>(&As[kk][i * 32 + ty * 2]);
        af[2 * i] = t.x;
        af[2 * i + 1] = t.y;
        const double2 u = *reinterpret_cast<const double2*>(&Bs[kk][i * 32 + tx * 2]);
        bf[2 * i] = u.x;
        bf[2 * i + 1] = u.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(af[i], bf[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long r = row0 + (i >> 1) * 32 + ty * 2 + (i & 1);
    if (r >= rows) continue;
    long traj = 0;
    int dof = 0;
    if (MODE == GEMM_SUB_BEADVEC) {
      traj = r / nm.ndof;
      dof = (int)(r - traj * nm.ndof);
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c = col0 + jj * 32 + tx * 2;
      double y0 = acc[i][2 * jj], y1 = acc[i][2 * jj + 1];
      if (MODE == GEMM_SUB_BEADVEC) {
        if (c < n) y0 = y0 - beadvec_at(nm, a, b, traj, dof, c);
        if (c + 1 < n) y1 = y1 - beadvec_at(nm, a, b, traj, dof, c + 1);
      }
      if (c + 1 < n && ((n & 1) == 0)) {
        *reinterpret_cast<double2*>(&Y[r * n + c]) = make_double2(y0, y1);
      } else {
        if (c < n) Y[r * n + c] = y0;
        if (c + 1 < n) Y[r * n + c + 1] = y1;
      }
    }
  }
}

// ---- the same contraction on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64) ------------------------------------
// north_star: "tensor cores only if ncu shows it beating the FMA pipe".  On B200 the dense FP64 tensor rate equals the
// DFMA rate, so the gain can only come from instruction economy: one m8n8k4 retires 256 FMAs per warp instruction
// against 32 for a DFMA, and a thread needs 2 shared-memory loads per 256 FMAs instead of 16 per 64.
// CTA tile 128 x 128 x 16, 8 warps as 4 (rows) x 2 (columns), warp tile 32 x 64 = 4 x 8 MMA tiles.
// Fragment layout (PTX ISA, m8n8k4 .f64): A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
// C[row = lane/4][col = 2*(lane%4) + {0,1}].  Leading dimensions are = 4 (mod 16) doubles so that the 16 fragment
// loads of a half-warp fall into 16 different 8-byte bank pairs.
constexpr int DK = 16, LDA = DK + 4;

// NT = n8 tiles per warp: 8 -> CTA tile 128 x 128 (one CTA per SM, 208 registers); 4 -> CTA tile 128 x 64 (two CTAs per SM)
template <int MODE, int NT>
__global__ void __launch_bounds__(GT, NT == 4 ? 2 : 1)
nm_gemm_dmma_kernel(NmTables nm, const double* __restrict__ A, double* __restrict__ Y, long rows,
                    const double* __restrict__ a, const double* __restrict__ b) {
  constexpr int TN = 16 * NT;            // CTA tile columns
  constexpr int LDB = TN + 4;
  constexpr int BQ = TN / 16;            // T-tile doubles staged per thread (16 threads per k row)
  __shared__ __align__(16) double As[BM * LDA];
  __shared__ __align__(16) double Bs[DK * LDB];
  const int n = nm.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;           // warp tile origin: rows 32*wm, columns 8*NT*wn
  const long row0 = (long)blockIdx.y * BM;
  const int col0 = blockIdx.x * TN;
  const int a_r = tid >> 1, a_k0 = (tid & 1) * 8;
  const int b_k = tid >> 4, b_c0 = (tid & 15) * BQ;
  const long a_row = row0 + a_r;
  const bool a_ok = a_row < rows;
  long a_traj = 0;
  int a_dof = 0;
  if (MODE == GEMM_ADD_BEADVEC && a_ok) {
    a_traj = a_row / nm.ndof;
    a_dof = (int)(a_row - a_traj * nm.ndof);
  }
  double ra[8], rb[BQ];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = k0 + a_k0 + q;
      double v = 0.0;
      if (a_ok && j < n) {
        v = A[a_row * n + j];
        if (MODE == GEMM_ADD_BEADVEC) v = v + beadvec_at(nm, a, b, a_traj, a_dof, j);
      }
      ra[q] = v;
    }
    const int j = k0 + b_k;
#pragma unroll
    for (int q = 0; q < BQ; ++q) {
      const int c = col0 + b_c0 + q;
      rb[q] = (j < n && c < n) ? nm.T[(long)j * n + c] : 0.0;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(&As[a_r * LDA + a_k0 + q]) = make_double2(ra[q], ra[q + 1]);
#pragma unroll
    for (int q = 0; q < BQ; q += 2) *reinterpret_cast<double2*>(&Bs[b_k * LDB + b_c0 + q]) = make_double2(rb[q], rb[q + 1]);
  };
  double acc[4][NT][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* Af = As + (wm * 32 + (lane >> 2)) * LDA + (lane & 3);
  const double* Bf = Bs + (lane & 3) * LDB + wn * 8 * NT + (lane >> 2);
  load_tiles(0);
  for (int k0 = 0; k0 < n; k0 += DK) {
    store_tiles();
    __syncthreads();
    if (k0 + DK < n) load_tiles(k0 + DK);
#pragma unroll
    for (int kk = 0; kk < DK; kk += 4) {
      double af[4], bf[NT];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Af[i * 8 * LDA + kk];
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = Bf[kk * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = row0 + wm * 32 + i * 8 + (lane >> 2);
    if (r >= rows) continue;
    long traj = 0;
    int dof = 0;
    if (MODE == GEMM_SUB_BEADVEC) {
      traj = r / nm.ndof;
      dof = (int)(r - traj * nm.ndof);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c = col0 + wn * 8 * NT + j * 8 + 2 * (lane & 3);
      double y0 = acc[i][j][0], y1 = acc[i][j][1];
      if (MODE == GEMM_SUB_BEADVEC) {
        if (c < n) y0 = y0 - beadvec_at(nm, a, b, traj, dof, c);
        if (c + 1 < n) y1 = y1 - beadvec_at(nm, a, b, traj, dof, c + 1);
      }
      if (c + 1 < n && ((n & 1) == 0)) {
        *reinterpret_cast<double2*>(&Y[r * n + c]) = make_double2(y0, y1);
      } else {
        if (c < n) Y[r * n + c] = y0;
        if (c + 1 < n) Y[r * n + c + 1] = y1;
      }
    }
  }
}


// ---- the production transform: DMMA tiles fed by a 3-stage cp.async pipeline ------------------------------------------------
// Same tile shape, fragment layout and fma/k order as nm_gemm_dmma_kernel above (identical bits), but the k-tiles of A and
// T travel global -> shared memory asynchronously (cp.async.cg, 16 bytes = 2 doubles per request, zero-filled outside the
// matrix), three tiles in flight, ONE barrier per k-tile: the tensor pipe no longer waits for a store-after-compute
// hand-over twice per tile.  The beadvec shift is not formed inside the main loop any more: the forward form subtracts a
// precomputed beadvec array E in the epilogue (MODE = GEMM_SUB_BEADVEC), the backward form is a plain product of the
// array Q + beadvec that the update kernel wrote (the reference's order: add, then transform; verletmodule.f90:274-279).
// Needs n even (16-byte alignment of every row); odd n takes the register-staged kernels above.
constexpr int kStages = 3;
template <int NT>
constexpr size_t pipe_smem_bytes() { return (size_t)kStages * (BM * LDA + DK * (16 * NT + 4)) * sizeof(double); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled, the source is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MODE = GEMM_KICK_ROTATE (forward transform of the gradient fused with the update that consumes it): the epilogue does not
// store G = g T but applies P <- P - dt G and the rotation of (P, Q) to the two modes its accumulator pair holds — exactly
// nm_update2_kernel's arithmetic for ops = OP_KICK | OP_ROT1 (| OP_CLOCK), so the normal-mode gradient never travels and the
// update kernel's launch is gone.  Y is not written.
struct KickRotateArgs {
  double* P;
  double* Q;
  double dt;
  int clock;                 // advance the Andersen collision clocks (one thread per trajectory)
  uint64_t seed, step;
  const int64_t* gid;
  int* flags;
  int* count;
  int* rkick;
  double lambda;
};
// MODE = GEMM_MODEL_PES (back-transform fused with the gradient of a model surface): the epilogue turns the bead positions
// its accumulators hold into the bead gradient and stores that; x itself is not written.  1D surface: every coordinate on its
// own.  Two-coordinate surfaces (2D double well, SO2 ring): rows 2t and 2t + 1 are the two coordinates of trajectory t and sit
// in lanes l and l ^ 4 of the accumulator layout; the even row's thread takes the bead of its first column, the odd row's thread
// the bead of the second, each evaluates simple_pes_eval once and the two exchange the components (three shuffles per pair).
struct ModelPesArgs {
  int kind;
  SimplePesParams P;
  int* flags;
};
template <int MODE, int NT>
__global__ void __launch_bounds__(GT, NT == 4 ? 2 : 1)
nm_gemm_pipe_kernel(NmTables nm, const double* __restrict__ A, double* __restrict__ Y, long rows,
                    const double* __restrict__ E, KickRotateArgs U, ModelPesArgs M) {
  constexpr int TN = 16 * NT, LDB = TN + 4, STAGE = BM * LDA + DK * LDB;
  extern __shared__ __align__(16) double pipe_smem[];
  const int n = nm.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const long row0 = (long)blockIdx.y * BM;
  const int col0 = blockIdx.x * TN;
  const int ktiles = (n + DK - 1) / DK;
  auto issue = [&](int kt) {
    if (kt < ktiles) {
      double* As = pipe_smem + (kt % kStages) * STAGE;
      double* Bs = As + BM * LDA;
      const int k0 = kt * DK;
#pragma unroll
      for (int c = tid; c < BM * (DK / 2); c += GT) {       // A tile: 128 rows x 8 requests
        const int r = c >> 3, ch = c & 7;
        const long gr = row0 + r;
        const int gk = k0 + 2 * ch;
        const bool ok = gr < rows && gk < n;
        cp_async16(As + r * LDA + 2 * ch, ok ? A + gr * n + gk : A, ok);
      }
#pragma unroll
      for (int c = tid; c < DK * (TN / 2); c += GT) {       // T tile: 16 rows x TN/2 requests
        const int r = c / (TN / 2), ch = c - r * (TN / 2);
        const int gk = k0 + r, gc = col0 + 2 * ch;
        const bool ok = gk < n && gc < n;
        cp_async16(Bs + r * LDB + 2 * ch, ok ? nm.T + (long)gk * n + gc : nm.T, ok);
      }
    }
    cp_async_commit();   // an empty group keeps the wait count uniform in the tail
  };
  double acc[4][NT][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<kStages - 2>();   // tile kt has landed (for this thread's requests) ...
    __syncthreads();                // ... and for everybody's; also: everybody is done with the stage issued next
    issue(kt + kStages - 1);
    const double* As = pipe_smem + (kt % kStages) * STAGE;
    const double* Af = As + (wm * 32 + (lane >> 2)) * LDA + (lane & 3);
    const double* Bf = As + BM * LDA + (lane & 3) * LDB + wn * 8 * NT + (lane >> 2);
#pragma unroll
    for (int kk = 0; kk < DK; kk += 4) {
      double af[4], bf[NT];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Af[i * 8 * LDA + kk];
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = Bf[kk * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
  }
  cp_async_wait<0>();
  if constexpr (MODE == GEMM_MODEL_PES) {   // every lane stays in the loops (shuffles); stores are predicated
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long r = row0 + wm * 32 + i * 8 + (lane >> 2);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int c = col0 + wn * 8 * NT + j * 8 + 2 * (lane & 3);
        const bool valid = r < rows && c < n;   // rows and n are even: a pair of rows / columns is inside or outside together
        const double y0 = acc[i][j][0], y1 = acc[i][j][1];
        double g0, g1;
        if (M.kind == PES_1D) {
          SimplePesParams P1 = M.P;
          P1.ndof = 1;
          simple_pes_eval<1>(M.kind, P1, &y0, nullptr, &g0, false, true);
          simple_pes_eval<1>(M.kind, P1, &y1, nullptr, &g1, false, true);
        } else {
          const int dof = (int)(r & 1);
          const double o0 = __shfl_xor_sync(0xffffffffu, y0, 4), o1 = __shfl_xor_sync(0xffffffffu, y1, 4);
          const double xx[2] = {dof ? o1 : y0, dof ? y1 : o0};   // the bead of column c (even row) or c + 1 (odd row)
          double gg[2];
          simple_pes_eval<2>(M.kind, M.P, xx, nullptr, gg, false, true);
          const double other = __shfl_xor_sync(0xffffffffu, dof ? gg[0] : gg[1], 4);   // the partner's row component of my bead
          g0 = dof ? other : gg[0];
          g1 = dof ? gg[1] : other;
        }
        if (valid) {
          if (g0 != g0 || g1 != g1) atomicOr(M.flags, PIMDK_FLAG_NAN);
          *reinterpret_cast<double2*>(&Y[r * n + c]) = make_double2(g0, g1);
        }
      }
    }
  } else {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = row0 + wm * 32 + i * 8 + (lane >> 2);
    if (r >= rows) continue;
    int akb = 0;
    if (MODE == GEMM_KICK_ROTATE) {
      const long traj = r / nm.ndof;
      const int dof = (int)(r - traj * nm.ndof);
      akb = (dof / nm.ndim) * n;
      if (U.clock && dof == 0 && col0 == 0 && wn == 0 && (lane & 3) == 0) {   // the thread that holds mode 0 of the trajectory's first row
        const uint32_t g = U.gid ? (uint32_t)U.gid[traj] : (uint32_t)traj;
        int c = U.count[traj] + 1;
        if (c >= U.rkick[traj]) {
          c = 0;
          U.rkick[traj] = poisson_norm(U.seed, U.step, g, U.lambda);
        }
        U.count[traj] = c;
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c = col0 + wn * 8 * NT + j * 8 + 2 * (lane & 3);
      if (c >= n) continue;          // n is even: c and c + 1 are inside or outside together
      double y0 = acc[i][j][0], y1 = acc[i][j][1];
      if (MODE == GEMM_KICK_ROTATE) {
        const size_t e = (size_t)r * n + c;
        double2 P = *reinterpret_cast<const double2*>(U.P + e);
        double2 Q = *reinterpret_cast<const double2*>(U.Q + e);
        P.x = P.x - y0 * U.dt;
        P.y = P.y - y1 * U.dt;
        rotate(nm, akb + c, P.x, Q.x);
        rotate(nm, akb + c + 1, P.y, Q.y);
        if (P.x != P.x || P.y != P.y) atomicOr(U.flags, PIMDK_FLAG_NAN);
        *reinterpret_cast<double2*>(U.P + e) = P;
        *reinterpret_cast<double2*>(U.Q + e) = Q;
        continue;
      }
      if (MODE == GEMM_SUB_BEADVEC) {
        const double2 e = *reinterpret_cast<const double2*>(&E[r * n + c]);
        y0 = y0 - e.x;
        y1 = y1 - e.y;
      }
      *reinterpret_cast<double2*>(&Y[r * n + c]) = make_double2(y0, y1);
    }
  }
  }
}

// beadvec(k, dof) of every (trajectory, dof) row, once per propagate call (a and b do not change during it): init_nm
// (verletmodule.f90:328-333) for the whole batch
__global__ void __launch_bounds__(256)
beadvec_kernel(NmTables nm, const double* __restrict__ a, const double* __restrict__ b, long rows, double* __restrict__ BV) {
  const long total = rows * nm.n;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / nm.n;
    const int k = (int)(e - r * nm.n);
    const long traj = r / nm.ndof;
    BV[e] = beadvec_at(nm, a, b, traj, (int)(r - traj * nm.ndof), k);
  }
}
__global__ void __launch_bounds__(256)
add_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ z, long total) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) z[e] = x[e] + y[e];
}

__global__ void __launch_bounds__(256)
nm_update_kernel(NmTables nm, double* __restrict__ Pn, double* __restrict__ Qn, const double* __restrict__ G,
                 double dt, long ntraj, int ops, uint64_t seed, uint64_t step, const int64_t* __restrict__ gid,
                 int* __restrict__ flags) {
  const long per_traj = (long)nm.ndof * nm.n;
  const long total = ntraj * per_traj;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long traj = e / per_traj;
    const long idx = e - traj * per_traj;  // dof*n + k
    const int dof = (int)(idx / nm.n);
    const int k = (int)(idx - (long)dof * nm.n);
    const int ak = (dof / nm.ndim) * nm.n + k;
    double P = Pn[e], Q = Qn[e];
    if (ops & OP_KICK) P = P - G[e] * dt;
    if (ops & OP_ROT1) rotate(nm, ak, P, Q);
    if (ops & OP_LANGEVIN) {
      const uint32_t g = gid ? (uint32_t)gid[traj] : (uint32_t)traj;
      const double xi = normal_at(seed, STREAM_LANGEVIN, step, g, (uint64_t)idx);
      P = nm.c1sq[ak] * P + nm.cnoise[ak] * xi;
    }
    if (ops & OP_ROT2) rotate(nm, ak, P, Q);
    if (P != P) atomicOr(flags, PIMDK_FLAG_NAN);
    Pn[e] = P;
    Qn[e] = Q;
  }
}

// The same update for even n, two consecutive modes (k, k+1) of one (trajectory, dof) row per thread: the two modes are
// the two normals of ONE Box-Muller pair (RNG contract: normal #idx lives in pair idx >> 1), so one Philox block, one log
// and one sincos serve both (the per-element kernel above evaluates them twice and drops half); 16-byte loads and
// stores; row -> (trajectory, dof) by one 32-bit division per thread instead of two 64-bit ones per element.  Same
// arithmetic per element, identical bits.  QB (optional) receives Q + beadvec for the back-transform that follows.
__global__ void __launch_bounds__(256)
nm_update2_kernel(NmTables nm, double* __restrict__ Pn, double* __restrict__ Qn, const double* __restrict__ G,
                  double dt, unsigned rows, int ops, uint64_t seed, uint64_t step, const int64_t* __restrict__ gid,
                  int* __restrict__ flags, const double* __restrict__ BV, double* __restrict__ QB, int* __restrict__ count,
                  int* __restrict__ rkick, double lambda) {
  const int n = nm.n;
  const unsigned rpc = (2 * blockDim.x >= (unsigned)n) ? (2 * blockDim.x) / (unsigned)n : 1;   // rows per CTA pass
  const unsigned tl = (2 * threadIdx.x) / (unsigned)n;            // this thread's row within the pass (0 when n >= 512)
  const unsigned k_first = 2 * threadIdx.x - tl * (unsigned)n;
  if (tl >= rpc) return;
  for (unsigned long long row0 = (unsigned long long)blockIdx.x * rpc; row0 < rows; row0 += (unsigned long long)gridDim.x * rpc) {
    const unsigned row = (unsigned)row0 + tl;
    if (row >= rows) continue;
    const unsigned traj = row / (unsigned)nm.ndof;
    const int dof = (int)(row - traj * (unsigned)nm.ndof);
    const int akb = (dof / nm.ndim) * n;
    const uint32_t g = gid ? (uint32_t)gid[traj] : traj;
    const size_t base = (size_t)row * n;
    // Andersen (verletmodule.f90:204-234): OP_ANDERSEN, first kernel of the step, only READS the collision clock — a
    // trajectory whose clock fires gets fresh momenta before the rotation; OP_CLOCK, second kernel of the step (stream
    // order: every read of the first is done), advances the clock, one thread per trajectory
    const bool fire = (ops & OP_ANDERSEN) && (count[traj] + 1 >= rkick[traj]);
    if ((ops & OP_CLOCK) && dof == 0 && k_first == 0) {
      int c = count[traj] + 1;
      if (c >= rkick[traj]) {
        c = 0;
        rkick[traj] = poisson_norm(seed, step, g, lambda);
      }
      count[traj] = c;
    }
    for (unsigned k = k_first; k < (unsigned)n; k += 2 * blockDim.x) {
      const size_t e = base + k;
      double2 P = *reinterpret_cast<const double2*>(Pn + e);
      double2 Q = *reinterpret_cast<const double2*>(Qn + e);
      const int ak = akb + (int)k;
      if (fire) {
        double z0, z1;
        normal_pair_at(seed, STREAM_ANDERSEN, step, g, (uint64_t)(((unsigned)dof * (unsigned)n + k) >> 1), z0, z1);
        P.x = (0.0 + nm.stdev * z0) * nm.sigp[ak];
        P.y = (0.0 + nm.stdev * z1) * nm.sigp[ak + 1];
      }
      if (ops & OP_KICK) {
        const double2 gg = *reinterpret_cast<const double2*>(G + e);
        P.x = P.x - gg.x * dt;
        P.y = P.y - gg.y * dt;
      }
      if (ops & OP_ROT1) {
        rotate(nm, ak, P.x, Q.x);
        rotate(nm, ak + 1, P.y, Q.y);
      }
      if (ops & OP_LANGEVIN) {
        double z0, z1;
        normal_pair_at(seed, STREAM_LANGEVIN, step, g, (uint64_t)(((unsigned)dof * (unsigned)n + k) >> 1), z0, z1);
        P.x = nm.c1sq[ak] * P.x + nm.cnoise[ak] * z0;
        P.y = nm.c1sq[ak + 1] * P.y + nm.cnoise[ak + 1] * z1;
      }
      if (ops & OP_ROT2) {
        rotate(nm, ak, P.x, Q.x);
        rotate(nm, ak + 1, P.y, Q.y);
      }
      if (P.x != P.x || P.y != P.y) atomicOr(flags, PIMDK_FLAG_NAN);
      *reinterpret_cast<double2*>(Pn + e) = P;
      *reinterpret_cast<double2*>(Qn + e) = Q;
      if (QB) {
        const double2 bv = *reinterpret_cast<const double2*>(BV + e);
        *reinterpret_cast<double2*>(QB + e) = make_double2(Q.x + bv.x, Q.y + bv.y);
      }
    }
  }
}

// momenta for even n, one Box-Muller pair per thread (see nm_update2_kernel)
__global__ void __launch_bounds__(256)
sample_momenta2_kernel(NmTables nm, double* __restrict__ Pn, unsigned rows, uint64_t seed, int stream, uint64_t step,
                       const int64_t* __restrict__ gid, const int* __restrict__ count, const int* __restrict__ rkick) {
  const int n = nm.n;
  const unsigned rpc = (2 * blockDim.x >= (unsigned)n) ? (2 * blockDim.x) / (unsigned)n : 1;
  const unsigned tl = (2 * threadIdx.x) / (unsigned)n;
  const unsigned k_first = 2 * threadIdx.x - tl * (unsigned)n;
  if (tl >= rpc) return;
  for (unsigned long long row0 = (unsigned long long)blockIdx.x * rpc; row0 < rows; row0 += (unsigned long long)gridDim.x * rpc) {
    const unsigned row = (unsigned)row0 + tl;
    if (row >= rows) continue;
    const unsigned traj = row / (unsigned)nm.ndof;
    if (count && !(count[traj] + 1 >= rkick[traj])) continue;
    const int dof = (int)(row - traj * (unsigned)nm.ndof);
    const int akb = (dof / nm.ndim) * n;
    const uint32_t g = gid ? (uint32_t)gid[traj] : traj;
    const size_t base = (size_t)row * n;
    for (unsigned k = k_first; k < (unsigned)n; k += 2 * blockDim.x) {
      double z0, z1;
      normal_pair_at(seed, stream, step, g, (uint64_t)(((unsigned)dof * (unsigned)n + k) >> 1), z0, z1);
      *reinterpret_cast<double2*>(Pn + base + k) =
          make_double2((0.0 + nm.stdev * z0) * nm.sigp[akb + (int)k], (0.0 + nm.stdev * z1) * nm.sigp[akb + (int)k + 1]);
    }
  }
}

__global__ void __launch_bounds__(256)
sample_momenta_kernel(NmTables nm, double* __restrict__ Pn, long ntraj, uint64_t seed, int stream, uint64_t step,
                      const int64_t* __restrict__ gid, const int* __restrict__ count, const int* __restrict__ rkick) {
  const long per_traj = (long)nm.ndof * nm.n;
  const long total = ntraj * per_traj;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long traj = e / per_traj;
    if (count && !(count[traj] + 1 >= rkick[traj])) continue;  // Andersen: only trajectories whose clock fired
    const long idx = e - traj * per_traj;
    const int dof = (int)(idx / nm.n);
    const int k = (int)(idx - (long)dof * nm.n);
    const int ak = (dof / nm.ndim) * nm.n + k;
    const uint32_t g = gid ? (uint32_t)gid[traj] : (uint32_t)traj;
    const double z = normal_at(seed, stream, step, g, (uint64_t)idx);
    Pn[e] = (0.0 + nm.stdev * z) * nm.sigp[ak];
  }
}

__global__ void andersen_clock_kernel(long ntraj, uint64_t seed, uint64_t step, double lambda,
                                      const int64_t* __restrict__ gid, int* __restrict__ count,
                                      int* __restrict__ rkick, int init) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntraj) return;
  const uint32_t g = gid ? (uint32_t)gid[t] : (uint32_t)t;
  if (init) {
    count[t] = 0;
    rkick[t] = poisson_norm(seed, step, g, lambda);
    return;
  }
  int c = count[t] + 1;  // count=count+1 ; if (count .ge. rkick) ...  (verletmodule.f90:204,208)
  if (c >= rkick[t]) {
    c = 0;
    rkick[t] = poisson_norm(seed, step, g, lambda);
  }
  count[t] = c;
}

__global__ void estimator_kernel(NmTables nm, const double* __restrict__ x, const double* __restrict__ dbdl,
                                 double* __restrict__ dHdr, long ntraj) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntraj) return;
  const double* xt = x + t * (long)nm.ndof * nm.n;
  double contr = 0.0;
  for (int j = 0; j < nm.ndim; ++j)
    for (int k = 0; k < nm.natom; ++k) {
      const int dof = k * nm.ndim + j;
      contr = contr + nm.mass[k] * (-xt[(long)dof * nm.n + (nm.n - 1)]) * dbdl[t * nm.ndof + dof];
    }
  dHdr[t] = dHdr[t] + contr;
}

// The estimator needs the LAST bead only: x(n,dof) = sum_j T(n,j) (Q_j + beadvec_j), the last column of the
// back-transform.  The Andersen step (NM(dt/2) V(dt) NM(dt/2)) has no other use for the positions at the end of a
// step, so this kernel replaces its third n x n transform per step by one length-n contraction per (trajectory, dof):
// the same fma chain, j ascending from zero, that the GEMM kernels run for that column — identical bits.
// A CTA owns whole trajectories (kEmRows / ndof of them): the rows' modes are staged through shared memory in
// coalesced 32-column tiles, thread = row runs the chain, then one thread per trajectory adds the ndof terms in the
// reference's order (j = dim outer, k = atom inner; verletmodule.f90:236-244).
constexpr int kEmRows = 32, kEmThreads = 256;
__global__ void __launch_bounds__(kEmThreads)
estimator_modes_kernel(NmTables nm, const double* __restrict__ Q, const double* __restrict__ a,
                       const double* __restrict__ b, const double* __restrict__ dbdl, double* __restrict__ dHdr,
                       long ntraj, const double* __restrict__ BV) {
  extern __shared__ double em_smem[];
  const int n = nm.n, ndof = nm.ndof;
  double* tcol = em_smem;                 // T(:, n): n doubles
  double* tile = em_smem + n;             // [kEmRows][33]
  double* xl = tile + kEmRows * 33;       // last-bead positions of the CTA's rows
  const int tpc = kEmRows / ndof;         // trajectories per CTA
  const long traj0 = (long)blockIdx.x * tpc;
  const int nrow = (int)(((ntraj - traj0 < tpc) ? ntraj - traj0 : tpc) * ndof);   // active rows
  const long row0 = traj0 * ndof;
  for (int m = threadIdx.x; m < n; m += kEmThreads) tcol[m] = nm.T[(long)m * n + (n - 1)];
  const int r = threadIdx.x;              // chain phase: thread = row (first warp)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  for (int m0 = 0; m0 < n; m0 += 32) {
    __syncthreads();
    for (int rr = warp; rr < nrow; rr += kEmThreads / 32) {   // Q + beadvec, formed lane-parallel by all warps
      const int m = m0 + lane, tl = rr / ndof;
      // beadvec: the call's precomputed array when there is one (the same expression, evaluated once per call)
      tile[rr * 33 + lane] = m < n ? Q[(row0 + rr) * (long)n + m] +
                                         (BV ? BV[(row0 + rr) * (long)n + m] : beadvec_at(nm, a, b, traj0 + tl, rr - tl * ndof, m))
                                   : 0.0;
    }
    __syncthreads();
    if (r < nrow) {
      const int mm = (n - m0 < 32) ? n - m0 : 32;
      for (int q = 0; q < mm; ++q) acc = fma(tile[r * 33 + q], tcol[m0 + q], acc);
    }
  }
  if (r < kEmRows) xl[r] = acc;
  __syncthreads();
  if (threadIdx.x < nrow / ndof) {
    const long t = traj0 + threadIdx.x;
    const double* xt = xl + threadIdx.x * ndof;
    double contr = 0.0;
    for (int j = 0; j < nm.ndim; ++j)
      for (int k = 0; k < nm.natom; ++k) {
        const int d = k * nm.ndim + j;
        contr = contr + nm.mass[k] * (-xt[d]) * dbdl[t * ndof + d];
      }
    dHdr[t] = dHdr[t] + contr;
  }
}

// The last-bead estimator of step i and the first update of step i + 1 in one kernel (Andersen steps, beadvec array present):
// both walk the same rows' modes — the estimator to contract Q + beadvec with T's last column, the update to resample the
// momenta of the trajectories whose collision clock fires, rotate (P, Q) by dt/2 and leave Q + beadvec for the back-transform.
// The tile loop of estimator_modes_kernel carries the update along: the estimator takes Q as the step left it, the update then
// overwrites it.  Per element the arithmetic of nm_update2_kernel (ops = OP_ANDERSEN | OP_ROT1), per row the chain of
// estimator_modes_kernel: the same bits as the two kernels.
// Warp 0 runs the rows' chains; warps 1..8 stage the 32-column tiles, one row each (two buffers: tile t + 1 is staged and updated while
// the chains walk tile t, one barrier per tile).
constexpr int kEuRows = 8, kEuThreads = 32 * (kEuRows + 1);   // small CTAs: the update needs the whole machine's warps
__global__ void __launch_bounds__(kEuThreads)
estimator_update_kernel(NmTables nm, double* __restrict__ Pn, double* __restrict__ Qn, const double* __restrict__ BV,
                        double* __restrict__ QB, const double* __restrict__ dbdl, double* __restrict__ dHdr, long ntraj, int do_est,
                        int do_upd, uint64_t seed, uint64_t step, const int64_t* __restrict__ gid, int* __restrict__ flags,
                        const int* __restrict__ count, const int* __restrict__ rkick) {
  extern __shared__ double em_smem[];
  const int n = nm.n, ndof = nm.ndof;
  double* tcol = em_smem;                 // T(:, n): n doubles
  double* tile = em_smem + n;             // [2][kEuRows][33]
  double* xl = tile + 2 * kEuRows * 33;   // last-bead positions of the CTA's rows
  const int tpc = kEuRows / ndof;         // trajectories per CTA
  const long traj0 = (long)blockIdx.x * tpc;
  const int nrow = (int)(((ntraj - traj0 < tpc) ? ntraj - traj0 : tpc) * ndof);   // active rows
  const long row0 = traj0 * ndof;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kStageWarps = kEuThreads / 32 - 1;
  const int ntiles = (n + 31) / 32;
  constexpr int kRowsPerWarp = (kEuRows + kStageWarps - 1) / kStageWarps;
  // what a staging warp needs to know of its rows, once: first normal-mode table index, RNG id, collision flag
  int akb_r[kRowsPerWarp], dofn_r[kRowsPerWarp];
  uint32_t g_r[kRowsPerWarp];
  bool fire_r[kRowsPerWarp];
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i) {
    const int rr = warp - 1 + i * kStageWarps;
    akb_r[i] = dofn_r[i] = 0;
    g_r[i] = 0;
    fire_r[i] = false;
    if (warp > 0 && rr < nrow) {
      const int tl = rr / ndof, dof = rr - tl * ndof;
      const long traj = traj0 + tl;
      akb_r[i] = (dof / nm.ndim) * n;
      dofn_r[i] = dof * n;
      g_r[i] = gid ? (uint32_t)gid[traj] : (uint32_t)traj;
      fire_r[i] = do_upd && count[traj] + 1 >= rkick[traj];   // Andersen collision at the start of the coming step (verletmodule.f90:204-234)
    }
  }
  auto stage = [&](int t) {               // warps 1..: rows of tile t -> buffer t & 1, updated in place behind the read
    double* buf = tile + (t & 1) * kEuRows * 33;
    const int m = t * 32 + lane;
    double Qv[kRowsPerWarp], Bv[kRowsPerWarp], Pv[kRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {          // every load of the warp's rows in flight before the first use
      const int rr = warp - 1 + i * kStageWarps;
      const bool ok = rr < nrow && m < n;
      const size_t e = ok ? (size_t)(row0 + rr) * n + m : 0;
      Qv[i] = ok ? Qn[e] : 0.0;
      Bv[i] = ok ? BV[e] : 0.0;
      Pv[i] = (ok && do_upd) ? Pn[e] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int rr = warp - 1 + i * kStageWarps;
      if (rr >= nrow) continue;
      double Q = Qv[i];
      const double bv = Bv[i];
      const double qb = Q + bv;
      if (m < n && do_upd) {
        const size_t e = (size_t)(row0 + rr) * n + m;
        const int ak = akb_r[i] + m;
        double P = Pv[i];
        if (fire_r[i]) {
          const double z = normal_at(seed, STREAM_ANDERSEN, step, g_r[i], (uint64_t)((unsigned)dofn_r[i] + (unsigned)m));
          P = (0.0 + nm.stdev * z) * nm.sigp[ak];
        }
        rotate(nm, ak, P, Q);
        if (P != P) atomicOr(flags, PIMDK_FLAG_NAN);
        Pn[e] = P;
        Qn[e] = Q;
        QB[e] = Q + bv;
      }
      buf[rr * 33 + lane] = qb;
    }
  };
  if (warp == 0) {
    if (do_est)
      for (int m = lane; m < n; m += 32) tcol[m] = nm.T[(long)m * n + (n - 1)];
  } else {
    stage(0);
  }
  double acc = 0.0;
  for (int t = 0; t < ntiles; ++t) {
    __syncthreads();                      // tile t is staged; the chains are done with the other buffer
    if (warp == 0) {
      if (do_est && lane < nrow) {
        const double* buf = tile + (t & 1) * kEuRows * 33;
        const int m0 = t * 32, mm = (n - m0 < 32) ? n - m0 : 32;
        for (int q = 0; q < mm; ++q) acc = fma(buf[lane * 33 + q], tcol[m0 + q], acc);
      }
    } else if (t + 1 < ntiles) {
      stage(t + 1);
    }
  }
  if (!do_est) return;
  if (warp == 0 && lane < kEuRows) xl[lane] = acc;
  __syncthreads();
  if (threadIdx.x < nrow / ndof) {
    const long t = traj0 + threadIdx.x;
    const double* xt = xl + threadIdx.x * ndof;
    double contr = 0.0;
    for (int j = 0; j < nm.ndim; ++j)
      for (int k = 0; k < nm.natom; ++k) {
        const int d = k * nm.ndim + j;
        contr = contr + nm.mass[k] * (-xt[d]) * dbdl[t * ndof + d];
      }
    dHdr[t] = dHdr[t] + contr;
  }
}

// ---- dHdrlimit (verletmodule.f90:404-409, propagate_pimd_pile only) -----------------------------------------------
// estimator with the outlier guard: |contr| < limit is added; otherwise the contribution is dropped and the trajectory is
// marked for re-initialisation (init_path again: beads back on the spline path, fresh momenta)
__global__ void estimator_limit_kernel(NmTables nm, const double* __restrict__ x, const double* __restrict__ dbdl,
                                       double* __restrict__ dHdr, long ntraj, double limit, int* __restrict__ reinit) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntraj) return;
  const double* xt = x + t * (long)nm.ndof * nm.n;
  double contr = 0.0;
  for (int j = 0; j < nm.ndim; ++j)
    for (int k = 0; k < nm.natom; ++k) {
      const int dof = k * nm.ndim + j;
      contr = contr + nm.mass[k] * (-xt[(long)dof * nm.n + (nm.n - 1)]) * dbdl[t * nm.ndof + dof];
    }
  const bool keep = fabs(contr) < limit || limit < 0.0;
  if (keep) dHdr[t] = dHdr[t] + contr;
  reinit[t] = keep ? 0 : 1;
}
// init_path (verletmodule.f90:39-48, 102-115) for the marked trajectories: x on the spline, momenta drawn directly in
// normal-mode space (the state lives there: p = T P), RNG stream 0 at the current step
__global__ void __launch_bounds__(256)
reinit_kernel(NmTables nm, int npath, const double* __restrict__ lampath, const double* __restrict__ path,
              const double* __restrict__ spl, const double* __restrict__ xi, double* __restrict__ x, double* __restrict__ P,
              long rows, uint64_t seed, uint64_t step, const int64_t* __restrict__ gid, const int* __restrict__ reinit) {
  const long row = blockIdx.x;
  if (row >= rows) return;
  const long traj = row / nm.ndof;
  if (!reinit[traj]) return;
  const int dof = (int)(row - traj * nm.ndof);
  const int n = nm.n;
  const uint32_t g = gid ? (uint32_t)gid[traj] : (uint32_t)traj;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double xv = (double)k * xi[traj] / (double)(n - 1);
    x[row * n + k] = splint_at(lampath, path + (long)dof * npath, spl + (long)dof * npath, npath, xv);
    const double z = normal_at(seed, STREAM_INIT, step, g, (uint64_t)dof * n + k);
    P[row * n + k] = (0.0 + nm.stdev * z) * nm.sigp[(dof / nm.ndim) * n + k];
  }
}
// Q = T x - beadvec for the marked trajectories (same fma chain, j ascending, as the transform kernels)
__global__ void __launch_bounds__(256)
reinit_q_kernel(NmTables nm, const double* __restrict__ x, const double* __restrict__ a, const double* __restrict__ b,
                double* __restrict__ Q, long rows, const int* __restrict__ reinit) {
  const long row = blockIdx.x;
  if (row >= rows) return;
  const long traj = row / nm.ndof;
  if (!reinit[traj]) return;
  const int dof = (int)(row - traj * nm.ndof);
  const int n = nm.n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc = fma(x[row * n + j], nm.T[(long)j * n + k], acc);
    Q[row * n + k] = acc - beadvec_at(nm, a, b, traj, dof, k);
  }
}

__global__ void scale_kernel(double* v, double s, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] / s;
}

unsigned grid_for(long total, int block) {
  long b = (total + block - 1) / block;
  const long cap = 148L * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

static int g_gemm_dmma = 1;   // default: tensor-core path (1.5x the FMA-pipe kernel, bit-identical results)
void set_nm_gemm_dmma(int on) { g_gemm_dmma = on; }

cudaError_t launch_nm_gemm(const NmTables& nm, GemmMode mode, const double* A, double* Y, long rows, const double* a,
                           const double* b, cudaStream_t st, const double* BV) {
  if (rows <= 0) return cudaSuccess;
  dim3 grid((nm.n + BN - 1) / BN, (unsigned)((rows + BM - 1) / BM));
  // production path: cp.async-pipelined DMMA tiles (even n; the forward shift needs the precomputed beadvec array)
  if (g_gemm_dmma && (nm.n & 1) == 0 && (mode == GEMM_PLAIN || (mode == GEMM_SUB_BEADVEC && BV))) {
    static unsigned long long attr_mask = 0;   // the >48 KB dynamic shared memory opt-in is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_mask & (1ull << (dev & 63)))) {
      cudaError_t e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_PLAIN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<4>());
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_SUB_BEADVEC, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<4>());
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_PLAIN, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<8>());
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_SUB_BEADVEC, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<8>());
      if (e != cudaSuccess) return e;
      attr_mask |= 1ull << (dev & 63);
    }
    if (g_gemm_dmma == 3) {
      if (mode == GEMM_PLAIN) nm_gemm_pipe_kernel<GEMM_PLAIN, 8><<<grid, GT, pipe_smem_bytes<8>(), st>>>(nm, A, Y, rows, BV, KickRotateArgs{}, ModelPesArgs{});
      else nm_gemm_pipe_kernel<GEMM_SUB_BEADVEC, 8><<<grid, GT, pipe_smem_bytes<8>(), st>>>(nm, A, Y, rows, BV, KickRotateArgs{}, ModelPesArgs{});
    } else {
      dim3 g4((nm.n + 63) / 64, (unsigned)((rows + BM - 1) / BM));
      if (mode == GEMM_PLAIN) nm_gemm_pipe_kernel<GEMM_PLAIN, 4><<<g4, GT, pipe_smem_bytes<4>(), st>>>(nm, A, Y, rows, BV, KickRotateArgs{}, ModelPesArgs{});
      else nm_gemm_pipe_kernel<GEMM_SUB_BEADVEC, 4><<<g4, GT, pipe_smem_bytes<4>(), st>>>(nm, A, Y, rows, BV, KickRotateArgs{}, ModelPesArgs{});
    }
    return cudaGetLastError();
  }
  if (g_gemm_dmma) {
    const int nt = g_gemm_dmma == 3 ? 8 : 4;   // 128 x 64 CTA tiles (two CTAs per SM) measured faster on every shape
    if (nt == 4) {
      dim3 g4((nm.n + 63) / 64, (unsigned)((rows + BM - 1) / BM));
      switch (mode) {
        case GEMM_PLAIN: nm_gemm_dmma_kernel<GEMM_PLAIN, 4><<<g4, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
        case GEMM_SUB_BEADVEC: nm_gemm_dmma_kernel<GEMM_SUB_BEADVEC, 4><<<g4, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
        case GEMM_ADD_BEADVEC: nm_gemm_dmma_kernel<GEMM_ADD_BEADVEC, 4><<<g4, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
      }
    } else {
      switch (mode) {
        case GEMM_PLAIN: nm_gemm_dmma_kernel<GEMM_PLAIN, 8><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
        case GEMM_SUB_BEADVEC: nm_gemm_dmma_kernel<GEMM_SUB_BEADVEC, 8><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
        case GEMM_ADD_BEADVEC: nm_gemm_dmma_kernel<GEMM_ADD_BEADVEC, 8><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
      }
    }
    return cudaGetLastError();
  }
  switch (mode) {
    case GEMM_PLAIN: nm_gemm_kernel<GEMM_PLAIN><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
    case GEMM_SUB_BEADVEC: nm_gemm_kernel<GEMM_SUB_BEADVEC><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
    case GEMM_ADD_BEADVEC: nm_gemm_kernel<GEMM_ADD_BEADVEC><<<grid, GT, 0, st>>>(nm, A, Y, rows, a, b); break;
  }
  return cudaGetLastError();
}

// G = g T fused with P <- P - dt G, rotate (and the Andersen clocks): available where the pipelined tensor-core kernel is
bool nm_gemm_fuses_kick_rotate(const NmTables& nm, long rows) { return g_gemm_dmma == 1 && (nm.n & 1) == 0 && rows > 0; }
cudaError_t launch_nm_gemm_kick_rotate(const NmTables& nm, const double* g, long rows, double* P, double* Q, double dt, int clock,
                                       uint64_t seed, uint64_t step, const int64_t* gid, int* flags, int* count, int* rkick,
                                       double lambda, cudaStream_t st) {
  if (!nm_gemm_fuses_kick_rotate(nm, rows)) return cudaErrorNotSupported;
  static unsigned long long attr_mask = 0;   // the >48 KB dynamic shared memory opt-in is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_mask & (1ull << (dev & 63)))) {
    cudaError_t e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_KICK_ROTATE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<4>());
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << (dev & 63);
  }
  dim3 g4((nm.n + 63) / 64, (unsigned)((rows + BM - 1) / BM));
  nm_gemm_pipe_kernel<GEMM_KICK_ROTATE, 4><<<g4, GT, pipe_smem_bytes<4>(), st>>>(
      nm, g, nullptr, rows, nullptr, KickRotateArgs{P, Q, dt, clock, seed, step, gid, flags, count, rkick, lambda}, ModelPesArgs{});
  return cudaGetLastError();
}

// grad V(A T) of a model surface in the back-transform's epilogue (A = Q + beadvec): where the pipelined tensor-core kernel runs,
// for the 1D surface (any number of coordinates) and for the two-coordinate surfaces
bool nm_gemm_fuses_model_pes(const NmTables& nm, long rows, int kind) {
  return g_gemm_dmma == 1 && (nm.n & 1) == 0 && rows > 0 && (kind == PES_1D || ((kind == PES_2DTEST || kind == PES_SO2) && nm.ndof == 2));
}
cudaError_t launch_nm_gemm_model_pes(const NmTables& nm, const double* A, long rows, int kind, const SimplePesParams& P, double* grad,
                                     int* flags, cudaStream_t st) {
  if (!nm_gemm_fuses_model_pes(nm, rows, kind)) return cudaErrorNotSupported;
  static unsigned long long attr_mask = 0;   // the >48 KB dynamic shared memory opt-in is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_mask & (1ull << (dev & 63)))) {
    cudaError_t e = cudaFuncSetAttribute(nm_gemm_pipe_kernel<GEMM_MODEL_PES, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem_bytes<4>());
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << (dev & 63);
  }
  dim3 g4((nm.n + 63) / 64, (unsigned)((rows + BM - 1) / BM));
  nm_gemm_pipe_kernel<GEMM_MODEL_PES, 4><<<g4, GT, pipe_smem_bytes<4>(), st>>>(nm, A, grad, rows, nullptr, KickRotateArgs{},
                                                                                ModelPesArgs{kind, P, flags});
  return cudaGetLastError();
}

// true when the streamed path keeps a precomputed beadvec array for this shape (see nm_gemm_pipe_kernel)
bool nm_uses_beadvec_array(const NmTables& nm) { return g_gemm_dmma != 0 && (nm.n & 1) == 0; }

cudaError_t launch_beadvec(const NmTables& nm, const double* a, const double* b, long rows, double* BV, cudaStream_t st) {
  beadvec_kernel<<<grid_for(rows * nm.n, 256), 256, 0, st>>>(nm, a, b, rows, BV);
  return cudaGetLastError();
}
cudaError_t launch_add(const double* x, const double* y, double* z, long total, cudaStream_t st) {
  add_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, z, total);
  return cudaGetLastError();
}

cudaError_t launch_nm_update(const NmTables& nm, double* P, double* Q, const double* G, double dt, long ntraj,
                             int do_kick, int nrot, int do_langevin, uint64_t seed, uint64_t step,
                             const int64_t* gid, int* flags, cudaStream_t st, const double* BV, double* QB, int andersen,
                             int* count, int* rkick, double lambda) {
  int ops = 0;
  if (do_kick) ops |= OP_KICK;
  if (nrot >= 1) ops |= OP_ROT1;
  if (do_langevin) ops |= OP_LANGEVIN;
  if (nrot >= 2) ops |= OP_ROT2;
  if (andersen == 1) ops |= OP_ANDERSEN;
  if (andersen == 2) ops |= OP_CLOCK;
  const long rows = ntraj * (long)nm.ndof;
  const long total = rows * nm.n;
  if ((nm.n & 1) == 0 && rows < 0xffffffffL) {
    const long rpc = 512 >= nm.n ? 512 / nm.n : 1;
    nm_update2_kernel<<<grid_for((rows + rpc - 1) / rpc, 1), 256, 0, st>>>(nm, P, Q, G, dt, (unsigned)rows, ops, seed, step, gid, flags,
                                                                        QB ? BV : nullptr, QB, count, rkick, lambda);
    return cudaGetLastError();
  }
  if (andersen) return cudaErrorInvalidValue;   // the fused Andersen forms exist in the paired kernel only (nm_update_fuses_andersen)
  nm_update_kernel<<<grid_for(total, 256), 256, 0, st>>>(nm, P, Q, G, dt, ntraj, ops, seed, step, gid, flags);
  if (QB) add_kernel<<<grid_for(total, 256), 256, 0, st>>>(Q, BV, QB, total);
  return cudaGetLastError();
}
bool nm_update_fuses_andersen(const NmTables& nm, long ntraj) { return (nm.n & 1) == 0 && ntraj * (long)nm.ndof < 0xffffffffL; }

static void sample_momenta_any(const NmTables& nm, double* P, long ntraj, uint64_t seed, int stream, uint64_t step,
                               const int64_t* gid, const int* count, const int* rkick, cudaStream_t st) {
  const long rows = ntraj * (long)nm.ndof;
  if ((nm.n & 1) == 0 && rows < 0xffffffffL) {
    const long rpc = 512 >= nm.n ? 512 / nm.n : 1;
    sample_momenta2_kernel<<<grid_for((rows + rpc - 1) / rpc, 1), 256, 0, st>>>(nm, P, (unsigned)rows, seed, stream, step, gid, count, rkick);
  } else {
    sample_momenta_kernel<<<grid_for(rows * nm.n, 256), 256, 0, st>>>(nm, P, ntraj, seed, stream, step, gid, count, rkick);
  }
}

cudaError_t launch_andersen(const NmTables& nm, double* P, long ntraj, uint64_t seed, uint64_t step, double lambda,
                            const int64_t* gid, int* count, int* rkick, cudaStream_t st) {
  sample_momenta_any(nm, P, ntraj, seed, STREAM_ANDERSEN, step, gid, count, rkick, st);
  andersen_clock_kernel<<<(unsigned)((ntraj + 127) / 128), 128, 0, st>>>(ntraj, seed, step, lambda, gid, count, rkick, 0);
  return cudaGetLastError();
}

cudaError_t launch_andersen_init(long ntraj, uint64_t seed, uint64_t step0, double lambda, const int64_t* gid, int* count,
                                 int* rkick, cudaStream_t st) {
  andersen_clock_kernel<<<(unsigned)((ntraj + 127) / 128), 128, 0, st>>>(ntraj, seed, step0, lambda, gid, count, rkick, 1);
  return cudaGetLastError();
}

cudaError_t launch_sample_momenta(const NmTables& nm, double* P, long ntraj, uint64_t seed, int stream, uint64_t step,
                                  const int64_t* gid, cudaStream_t st) {
  sample_momenta_any(nm, P, ntraj, seed, stream, step, gid, nullptr, nullptr, st);
  return cudaGetLastError();
}

cudaError_t launch_estimator(const NmTables& nm, const double* x, const double* dbdl, double* dHdr, long ntraj,
                             cudaStream_t st) {
  estimator_kernel<<<(unsigned)((ntraj + 127) / 128), 128, 0, st>>>(nm, x, dbdl, dHdr, ntraj);
  return cudaGetLastError();
}

cudaError_t launch_estimator_modes(const NmTables& nm, const double* Q, const double* a, const double* b,
                                   const double* dbdl, double* dHdr, long ntraj, cudaStream_t st, const double* BV) {
  const int tpc = kEmRows / nm.ndof;     // ndof <= 32 (the caller takes the full back-transform beyond that)
  if (tpc < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(nm.n + kEmRows * 33 + kEmRows) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(estimator_modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  estimator_modes_kernel<<<(unsigned)((ntraj + tpc - 1) / tpc), kEmThreads, smem, st>>>(nm, Q, a, b, dbdl, dHdr, ntraj, BV);
  return cudaGetLastError();
}

cudaError_t launch_estimator_update(const NmTables& nm, double* P, double* Q, const double* BV, double* QB, const double* dbdl,
                                    double* dHdr, long ntraj, int do_est, int do_upd, uint64_t seed, uint64_t step, const int64_t* gid,
                                    int* flags, const int* count, const int* rkick, cudaStream_t st) {
  const int tpc = kEuRows / nm.ndof;
  if (tpc < 1 || !BV || !QB) return cudaErrorInvalidValue;
  if (!do_est && !do_upd) return cudaSuccess;
  const size_t smem = (size_t)(nm.n + 2 * kEuRows * 33 + kEuRows) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(estimator_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  estimator_update_kernel<<<(unsigned)((ntraj + tpc - 1) / tpc), kEuThreads, smem, st>>>(nm, P, Q, BV, QB, dbdl, dHdr, ntraj, do_est, do_upd,
                                                                                       seed, step, gid, flags, count, rkick);
  return cudaGetLastError();
}

cudaError_t launch_estimator_limit(const NmTables& nm, const double* x, const double* dbdl, double* dHdr, long ntraj, double limit,
                                   int* reinit, cudaStream_t st) {
  estimator_limit_kernel<<<(unsigned)((ntraj + 127) / 128), 128, 0, st>>>(nm, x, dbdl, dHdr, ntraj, limit, reinit);
  return cudaGetLastError();
}
cudaError_t launch_reinit(const NmTables& nm, int npath, const double* lampath, const double* path, const double* spl,
                          const double* xi, double* x, double* P, double* Q, const double* a, const double* b, long ntraj,
                          uint64_t seed, uint64_t step, const int64_t* gid, const int* reinit, cudaStream_t st) {
  const long rows = ntraj * (long)nm.ndof;
  if (rows > 0x7fffffffL) return cudaErrorInvalidValue;
  reinit_kernel<<<(unsigned)rows, 256, 0, st>>>(nm, npath, lampath, path, spl, xi, x, P, rows, seed, step, gid, reinit);
  reinit_q_kernel<<<(unsigned)rows, 256, 0, st>>>(nm, x, a, b, Q, rows, reinit);
  return cudaGetLastError();
}

cudaError_t launch_scale(double* v, double s, long n, cudaStream_t st) {
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, s, n);
  return cudaGetLastError();
}

}  // namespace pimdk
