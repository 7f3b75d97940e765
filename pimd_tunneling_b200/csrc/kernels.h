// Internal launch interface between the C ABI (pimdk_api.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ccpol_tables.h"

namespace pimdk {

enum { PIMDK_FLAG_NAN = 1, PIMDK_FLAG_NOCONV = 2, PIMDK_FLAG_BADGID = 4 };

// Where geometry g (one bead of one ring polymer, or one batch entry) lives in a coordinate array.
//   ABI batch   x(ndim,natom,nbatch):          n_inner=1, stride_outer=ndof, stride_dof=1
//   state       x(n,ndim,natom,ntraj) (device): n_inner=n, stride_outer=ndof*n, stride_inner=1, stride_dof=n
// dof index = atom*ndim + dim in both.
struct GeomLayout {
  long n_inner, stride_outer, stride_inner, stride_dof;
  __host__ __device__ __forceinline__ long base(long g) const {
    const long o = g / n_inner;
    return o * stride_outer + (g - o * n_inner) * stride_inner;
  }
};

enum PesKind { PES_NONE = 0, PES_1D = 1, PES_2DTEST = 2, PES_CCPOL = 3, PES_SO2 = 4, PES_WATMETH = 5, PES_MALON = 6 };

struct SimplePesParams {  // mcmod_1d.f90:9-12, mcmod_2dtest.f90:16-24, mcmod_so2.f90:10-14
  double Vheight, x0;
  double omegaforce, r0;   // so2: harmonic ring V = omegaforce**2/2 (r - r0)**2
  double a0, b0;
  double wx[6], wy[6];
  double V0;
  int ndof;
};

// ---- CCpol (ccpol_kernels.cu, built twice) ----
// v != NULL: energies; grad != NULL: finite-difference gradients (write_drift: leave x where the
// reference's in-place perturbation leaves it).  `work` is a staging buffer of ccpol_work_bytes().
#define PIMDK_DECL_CCPOL(sfx)                                                                                       \
  size_t ccpol_work_bytes_##sfx(long ngeom, int grad);                                                              \
  void ccpol_host_tables_##sfx(const CcpolDev* host_tables);   /* whenever the tables change, before the next launch */ \
  long ccpol_launches_##sfx(long ngeom, int grad, int icc, size_t work_bytes);                                                              \
  cudaError_t launch_ccpol_##sfx(const CcpolDev* tab, int iemonomer, int iembed, int icc, int potparts_old, \
                                 double V0, GeomLayout L, double* x, double* v, \
                                 double* grad, long ngeom, int write_drift, int* flags, double* work,               \
                                 size_t work_bytes, cudaStream_t st);
PIMDK_DECL_CCPOL(strict)
PIMDK_DECL_CCPOL(fast)
#undef PIMDK_DECL_CCPOL

// ---- CCpol analytic gradient (ccpol_grad_kernels.cu; opt-in mode PIMDK_MODE_ANALYTIC) ----
namespace agrad { struct CcpolGradTab; }
size_t ccpol_analytic_bytes_per_geom();
void ccpol_host_tables_analytic(const CcpolDev* host_tables);   // whenever the tables change, before the next launch
long ccpol_analytic_launches(long ngeom, int icc, size_t work_bytes);
cudaError_t launch_ccpol_analytic(const CcpolDev* tab, const agrad::CcpolGradTab* gt, int iemonomer, int icc, double V0, GeomLayout L,
                                  const double* x, double* v, double* grad, long ngeom, int* flags, double* work, size_t work_bytes,
                                  int num_sms, cudaStream_t st);

// ---- water-methane rigid-body surface (watmeth_kernels.cu; watermethane.f90 / mcmod_watmeth.f90) ----
struct WatMethTab;
cudaError_t launch_watmeth(const WatMethTab* tab, GeomLayout L, const double* x, double* v, double* grad, long ngeom, int* flags,
                           cudaStream_t st);
cudaError_t launch_watmeth_hessian(const WatMethTab* tab, GeomLayout L, double* x, double* hess, long ngeom, cudaStream_t st);

// ---- malonaldehyde surface (malon_kernels.cu; pes_malonaldehyde.f90 / mcmod_malon.f90) ----
struct MalonTab;
cudaError_t launch_malon(const MalonTab* tab, GeomLayout L, const double* x, double V0, double* v, double* grad, long ngeom, int* flags,
                         cudaStream_t st);
cudaError_t launch_malon_hessian(const MalonTab* tab, GeomLayout L, const double* x, double* hess, long ngeom, cudaStream_t st);

// ---- 1D / 2D model surfaces (pes_simple.cu) ----
cudaError_t launch_simple_pes(PesKind kind, const SimplePesParams& P, GeomLayout L, const double* x, double* v,
                              double* grad, long ngeom, int* flags, cudaStream_t st);

// ---- normal-mode machinery (nm_kernels.cu) ----
struct NmTables {       // device pointers, built once per pimdk_nm_setup / propagate call
  const double* T;      // transmatrix(n,n), symmetric
  const double* sA;     // sin(k pi/(n+1)),      k=1..n     (beadvec pieces, verletmodule.f90:328-333)
  const double* sB;     // sin(n k pi/(n+1))
  const double* lamb2;  // (lam_k*betan)**2
  const double* cosw;   // [atom][k] cos(time*omegak)      (step_nm :528-531), time = dt/2
  const double* sinw;   // [atom][k] sin(omegak*time)
  const double* omega;  // [atom][k] omegak
  const double* bmass;  // [atom][k] beadmass(atom,k)
  const double* wbm;    // [atom][k] omegak*beadmass
  const double* c1sq;   // [atom][k] c1**2                  (step_langevin :651-652)
  const double* cnoise; // [atom][k] sqrt(beadmass/betan)*c2*sqrt(1+c1**2)
  const double* sigp;   // [atom][k] sqrt(beadmass) (momentum resampling :218, init_path :107)
  const double* mass;   // [atom]
  double norm;          // sqrt(2/(n+1))
  double stdev;         // sqrt(1/betan)
  int n, ndim, natom, ndof;
  int cayley;
  double time;          // dt/2
};

// Y[r][k] = sum_j f(A)[r][j] * T[j][k]  for r in [0,rows): FP64 tile GEMM with fused prologue/epilogue.
enum GemmMode {
  GEMM_PLAIN = 0,       // Y = A T
  GEMM_SUB_BEADVEC = 1, // Y = A T - beadvec            (nmtransform_forward, bead>0)
  GEMM_ADD_BEADVEC = 2, // Y = (A + beadvec) T          (nmtransform_backward, bead>0)
  GEMM_KICK_ROTATE = 3, // G = A T consumed in the epilogue: P <- P - dt G, rotate(P, Q)   (launch_nm_gemm_kick_rotate)
  GEMM_MODEL_PES = 4,   // Y = grad V(A T) of a 1D / two-coordinate model surface           (launch_nm_gemm_model_pes)
};
void set_nm_gemm_dmma(int on);
// BV (optional): beadvec(k, dof) of every row, precomputed by launch_beadvec; with it (and n even) GEMM_PLAIN and
// GEMM_SUB_BEADVEC run the cp.async-pipelined tensor-core kernel, and the backward form is a GEMM_PLAIN of Q + beadvec
// (written by launch_nm_update / launch_add) instead of GEMM_ADD_BEADVEC.
cudaError_t launch_nm_gemm(const NmTables& nm, GemmMode mode, const double* A, double* Y, long rows,
                           const double* a /*(ndof)*/, const double* b /*(ndof,ntraj)*/, cudaStream_t st,
                           const double* BV = nullptr);
bool nm_uses_beadvec_array(const NmTables& nm);
// The forward transform of the gradient fused with the update that consumes it (step_v's kick + one step_nm rotation, and the
// Andersen collision clocks when clock != 0): same arithmetic per element as launch_nm_gemm + launch_nm_update(do_kick = 1,
// nrot = 1, no O-step), hence the same bits; the normal-mode gradient is never stored.
// The back-transform fused with the gradient of a model surface (mcmod_1d, mcmod_2dtest, mcmod_so2): grad = Vprime(A T), bead
// positions never stored; the same simple_pes_eval per bead as launch_simple_pes, hence the same bits.
bool nm_gemm_fuses_model_pes(const NmTables& nm, long rows, int kind);
cudaError_t launch_nm_gemm_model_pes(const NmTables& nm, const double* A, long rows, int kind, const SimplePesParams& P, double* grad,
                                     int* flags, cudaStream_t st);
bool nm_gemm_fuses_kick_rotate(const NmTables& nm, long rows);
cudaError_t launch_nm_gemm_kick_rotate(const NmTables& nm, const double* g, long rows, double* P, double* Q, double dt, int clock,
                                       uint64_t seed, uint64_t step, const int64_t* gid, int* flags, int* count, int* rkick,
                                       double lambda, cudaStream_t st);
cudaError_t launch_beadvec(const NmTables& nm, const double* a, const double* b, long rows, double* BV, cudaStream_t st);
cudaError_t launch_add(const double* x, const double* y, double* z, long total, cudaStream_t st);

// P <- P - dt*G ; then rotate(dt/2) . O-step(Philox) . rotate(dt/2) on (P,Q) in normal-mode space (PILE), or
// rotate only (thermostat handled by the caller for Andersen).
cudaError_t launch_nm_update(const NmTables& nm, double* P, double* Q, const double* G, double dt, long ntraj,
                             int do_kick, int nrot, int do_langevin, uint64_t seed, uint64_t step,
                             const int64_t* gid, int* flags, cudaStream_t st, const double* BV = nullptr,
                             double* QB = nullptr /* receives Q + BV */,
                             int andersen = 0 /* 1: resample fired trajectories first; 2: advance the collision clocks */,
                             int* count = nullptr, int* rkick = nullptr, double lambda = 0.0);
bool nm_update_fuses_andersen(const NmTables& nm, long ntraj);
// Andersen: P <- N(0, sqrt(1/betan))*sqrt(beadmass) for trajectories whose counter fired; updates counters.
cudaError_t launch_andersen(const NmTables& nm, double* P, long ntraj, uint64_t seed, uint64_t step, double lambda,
                            const int64_t* gid, int* count, int* rkick, cudaStream_t st);
cudaError_t launch_andersen_init(long ntraj, uint64_t seed, uint64_t step0, double lambda, const int64_t* gid, int* count, int* rkick,
                                 cudaStream_t st);
// init_path momenta directly in normal-mode space (stream 0)
cudaError_t launch_sample_momenta(const NmTables& nm, double* P, long ntraj, uint64_t seed, int stream, uint64_t step,
                                  const int64_t* gid, cudaStream_t st);
// estimator: dHdr[traj] += sum_{dim,atom} mass*(-x(n,dim,atom))*dbdl(dim,atom,traj)   (verletmodule.f90:397-403)
// Andersen steps with a beadvec array: the last-bead estimator of step i (do_est) and the first update of step i + 1 (do_upd:
// resampling of the trajectories whose clock fires, rotation by dt/2, Q + beadvec into QB; `step` is that step's index) in one
// kernel; the bits of launch_estimator_modes followed by launch_nm_update(do_kick = 0, nrot = 1, andersen = 1).
cudaError_t launch_estimator_update(const NmTables& nm, double* P, double* Q, const double* BV, double* QB, const double* dbdl,
                                    double* dHdr, long ntraj, int do_est, int do_upd, uint64_t seed, uint64_t step, const int64_t* gid,
                                    int* flags, const int* count, const int* rkick, cudaStream_t st);
cudaError_t launch_estimator_modes(const NmTables& nm, const double* Q, const double* a, const double* b,
                                   const double* dbdl, double* dHdr, long ntraj, cudaStream_t st,
                                   const double* BV = nullptr);
cudaError_t launch_estimator(const NmTables& nm, const double* x, const double* dbdl, double* dHdr, long ntraj,
                             cudaStream_t st);
// dHdrlimit (verletmodule.f90:404-409): estimator with the outlier guard, and init_path again for the marked trajectories
cudaError_t launch_estimator_limit(const NmTables& nm, const double* x, const double* dbdl, double* dHdr, long ntraj, double limit,
                                   int* reinit, cudaStream_t st);
cudaError_t launch_reinit(const NmTables& nm, int npath, const double* lampath, const double* path, const double* spl,
                          const double* xi, double* x, double* P, double* Q, const double* a, const double* b, long ntraj,
                          uint64_t seed, uint64_t step, const int64_t* gid, const int* reinit, cudaStream_t st);
cudaError_t launch_scale(double* v, double s, long n, cudaStream_t st);

// ---- fused warp-per-ring-polymer propagation for small systems (fused_small.cu) ----
bool fused_small_supported(PesKind kind, int n, int ndim, int natom);
cudaError_t launch_fused_small(const NmTables& nm, PesKind kind, const SimplePesParams& pp, int thermostat, long ntraj,
                               double* x, double* p, const double* a, const double* b, const double* dbdl, double dt,
                               long NMC, long imin, double lambda, uint64_t seed, const int64_t* gid, double* dHdr,
                               int* flags, long step0, int keep_sum, double* dHsum, int* clk_count, int* clk_kick, int carry,
                               double dhdrlimit, int npath, const double* lampath, const double* path, const double* spl,
                               const double* xi, cudaStream_t st);

// ---- second derivatives (hess_kernels.cu): Vdoubleprime, UMhessian (instantonmod.f90:155-217) ----
cudaError_t launch_simple_hessian(PesKind kind, const SimplePesParams& P, int ndim, int natom, GeomLayout L, double* x,
                                  double* hess, long ngeom, cudaStream_t st);
cudaError_t launch_perturb(GeomLayout L, double* x, long ngeom, int dof, double delta, cudaStream_t st);
cudaError_t launch_hess_column(GeomLayout L, const double* gp, const double* gm, long ngeom, int nd, int dof1, double eps,
                               double* hess, cudaStream_t st);
cudaError_t launch_um_band(int n, int ndim, int natom, const double* hess, const double* mass, double betan, int singlewell,
                           double* band, cudaStream_t st);
cudaError_t launch_band_to_dense(long N, int kd, const double* band, double* A, cudaStream_t st);
cudaError_t launch_readhess_displace(int n, int ndim, int natom, const double* eta, const double* Z, const double* mass,
                                     double stdev, uint64_t seed, uint32_t gid, double* tempx, double* x, cudaStream_t st);

// ---- ring-polymer potential (um_kernels.cu): instantonmod.f90:17-151 ----
cudaError_t launch_um(int npoly, int n, int ndim, int natom, const double* x, const double* a, const double* b,
                      const double* mass, double betan, int fixedends, const double* vbead /*n or NULL*/,
                      const double* gbead /*(n,ndim,natom) or NULL*/, double* um_out /*1, may be NULL*/,
                      double* grad_out /*(n,ndim,natom) or NULL*/, cudaStream_t st);

// FP64 pipe peak probe (fp64_peak.cu): returns achieved DFMA TFLOP/s
cudaError_t fp64_peak_probe(int num_sms, double* tflops, cudaStream_t st);
// bit-equality self-test of the three-instruction division by small integers used in the damping series
cudaError_t div_selftest(unsigned long long* mismatches, cudaStream_t st);
// bit-equality self-test of fast_div / fast_sqrt (ccpol_device.cuh) against the built-ins
cudaError_t math_eval(int kind, long n, const double* hx, double* hy, cudaStream_t st);
cudaError_t fastmath_selftest(unsigned long long* mismatches, cudaStream_t st);

}  // namespace pimdk
