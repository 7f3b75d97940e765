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
                       long ntraj) {
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
      tile[rr * 33 + lane] = m < n ? Q[(row0 + rr) * (long)n + m] + beadvec_at(nm, a, b, traj0 + tl, rr - tl * ndof, m) : 0.0;
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
                           const double* b, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  dim3 grid((nm.n + BN - 1) / BN, (unsigned)((rows + BM - 1) / BM));
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

cudaError_t launch_nm_update(const NmTables& nm, double* P, double* Q, const double* G, double dt, long ntraj,
                             int do_kick, int nrot, int do_langevin, uint64_t seed, uint64_t step,
                             const int64_t* gid, int* flags, cudaStream_t st) {
  int ops = 0;
  if (do_kick) ops |= OP_KICK;
  if (nrot >= 1) ops |= OP_ROT1;
  if (do_langevin) ops |= OP_LANGEVIN;
  if (nrot >= 2) ops |= OP_ROT2;
  const long total = ntraj * (long)nm.ndof * nm.n;
  nm_update_kernel<<<grid_for(total, 256), 256, 0, st>>>(nm, P, Q, G, dt, ntraj, ops, seed, step, gid, flags);
  return cudaGetLastError();
}

cudaError_t launch_andersen(const NmTables& nm, double* P, long ntraj, uint64_t seed, uint64_t step, double lambda,
                            const int64_t* gid, int* count, int* rkick, cudaStream_t st) {
  const long total = ntraj * (long)nm.ndof * nm.n;
  sample_momenta_kernel<<<grid_for(total, 256), 256, 0, st>>>(nm, P, ntraj, seed, STREAM_ANDERSEN, step, gid, count,
                                                              rkick);
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
  const long total = ntraj * (long)nm.ndof * nm.n;
  sample_momenta_kernel<<<grid_for(total, 256), 256, 0, st>>>(nm, P, ntraj, seed, stream, step, gid, nullptr, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_estimator(const NmTables& nm, const double* x, const double* dbdl, double* dHdr, long ntraj,
                             cudaStream_t st) {
  estimator_kernel<<<(unsigned)((ntraj + 127) / 128), 128, 0, st>>>(nm, x, dbdl, dHdr, ntraj);
  return cudaGetLastError();
}

cudaError_t launch_estimator_modes(const NmTables& nm, const double* Q, const double* a, const double* b,
                                   const double* dbdl, double* dHdr, long ntraj, cudaStream_t st) {
  const int tpc = kEmRows / nm.ndof;     // (ndof <= 18 for every surface here)
  if (tpc < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(nm.n + kEmRows * 33 + kEmRows) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(estimator_modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  estimator_modes_kernel<<<(unsigned)((ntraj + tpc - 1) / tpc), kEmThreads, smem, st>>>(nm, Q, a, b, dbdl, dHdr, ntraj);
  return cudaGetLastError();
}

cudaError_t launch_scale(double* v, double s, long n, cudaStream_t st) {
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, s, n);
  return cudaGetLastError();
}

}  // namespace pimdk
