/* pimdk.h — C ABI of the B200-native ring-polymer hot path.
 *
 * This is the drop-in boundary for christophevaillant/pimd-tunneling: every entry point replaces
 * a piece of the reference's Fortran (file:line cited per function) and is what the reference's
 * ISO_C_BINDING shim (fortran/pimdk_mod.f90) binds.  Conventions follow the
 * reference build (`ifort -i8 -r8`, makefile:5): all integers are 64-bit, all reals FP64, all
 * arrays are Fortran column-major exactly as the reference declares them:
 *     x(n, ndim, natom [, ntraj])     bead index fastest, trajectory slowest
 *     a(ndim, natom), b/dbdl(ndim, natom [, ntraj])
 * Host-pointer calls borrow the caller's arrays for the duration of the call only.  The `_dev`
 * variants take device pointers on the library's device (same layouts) and enqueue on the
 * library stream without synchronising.  Every call returns 0 on success or a PIMDK_E* code;
 * pimdk_last_error() gives the text.  The library never calls exit(); the Fortran shim turns a
 * non-zero code into the reference's `write(*,*) msg; stop`.  Not thread-safe (neither is the
 * reference: module globals and COMMON /ddaattaa/).  Like the reference's linked-in mcmod_<PES>.o there is ONE
 * selected PES, one V0 and one set of normal-mode tables per process: pimdk_pes_select / pimdk_pes_set_v0 /
 * pimdk_nm_setup replace the previous selection for every later call.
 *
 * There is no CPU fallback: without a CUDA device every compute call fails with PIMDK_ENODEV.
 */
#ifndef PIMDK_H
#define PIMDK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t pimdk_int;

enum {
  PIMDK_OK = 0,
  PIMDK_EINVAL = 1,   /* bad argument / call order */
  PIMDK_ENODEV = 2,   /* no usable CUDA device */
  PIMDK_ECUDA = 3,    /* CUDA runtime error */
  PIMDK_EDATA = 4,    /* PES data files missing or malformed (reference: stop 010..040) */
  PIMDK_ENAN = 5,     /* NaN trap (verletmodule.f90:533-536,577-580 "NaN in pot propagation") */
  PIMDK_ENOCONV = 6   /* "No convergence in indN_iter" (proc_ccpol8s-dimer_xyz_ncd.f:364-369) */
};

enum { PIMDK_THERMOSTAT_ANDERSEN = 1, PIMDK_THERMOSTAT_PILE = 2 }; /* namelist `thermostat`, pimd_par.f90:68,371-378 */
enum { PIMDK_MODE_STRICT = 0, PIMDK_MODE_FAST = 1, PIMDK_MODE_ANALYTIC = 2 };

/* Library / device lifetime.  device < 0 keeps the current CUDA device.  data_dir is where the
 * CCpol-8sf parameter files live: either the reference's own data_SAPT5spfIR_2006 / data_CCpol8s /
 * data_ccdata (which the reference opens from its CWD: main_CCpol-8sf.f:49,115;
 * proc_ccpol8s-dimer_xyz_ncd.f:41; main_CCpol-8sf.f:872) or the packed *.tbl files shipped with
 * this package.  NULL = "." like the reference. */
int pimdk_init(pimdk_int device, const char* data_dir);
int pimdk_finalize(void);
const char* pimdk_last_error(void);
/* Launch everything on this cudaStream_t (default: the legacy default stream). */
int pimdk_set_stream(void* cuda_stream);
/* PIMDK_MODE_STRICT (default): CCpol arithmetic in the reference's operation order without FMA
 * contraction; PIMDK_MODE_FAST: same kernels with contraction.
 * PIMDK_MODE_ANALYTIC (opt-in, NOT the reference's arithmetic): Vprime of ccpol8sf becomes the analytic gradient of the
 * same energy expression (reverse-mode chain rule through the site-pair sums, about two energies per bead instead of
 * the 36 of the central difference, mcmod_waterdimer_ccpol.f90:40-58).  It agrees with the oracle's dual-number
 * gradient to ~1e-12 and with the finite-difference default to that difference's truncation error (~3e-8 of
 * max|grad|); x is not perturbed, so no drift is left behind.  Energies (V) are still the strict ones.  Offered for
 * the Radau-embedded surfaces with potparts (isurf 3 = the plugin's, and 10). */
int pimdk_set_mode(pimdk_int mode);
/* Small systems (1D/2D surfaces, n <= 128 beads) are propagated by one persistent warp-per-ring-polymer
 * kernel (default on); 0 forces the streamed multi-kernel path.  Both give bit-identical results. */
int pimdk_set_fused(pimdk_int enable);
/* Normal-mode transform engine: 0 = FP64 FMA-pipe tile GEMM; 1 (default) = FP64 tensor-core (DMMA m8n8k4) tile GEMM with
 * 128 x 64 CTA tiles, 2 = the same, 3 = DMMA with 128 x 128 CTA tiles.  All engines accumulate in k order with fused
 * multiply-adds and give identical bits. */
int pimdk_set_gemm(pimdk_int kind);

/* ---- PES plugin: module mcmod_mass -------------------------------------------------------
 * pimdk_pes_select  = V_init  (mcmod_1d.f90:8, mcmod_2dtest.f90:11, mcmod_waterdimer_ccpol.f90:9
 *                     -> init_ccpol(3,1,1,0), main_CCpol-8sf.f:1-173)
 *   name: "1d" | "2dtest" | "so2" | "watmeth" | "malon" | "ccpol8sf".  pes_params (optional): "1d": {Vheight, x0};
 *   "2dtest": {a0, b0, rho0}; "so2" (mcmod_so2.f90:10-48, the harmonic ring V = omegaforce**2/2 (r - r0)**2 in two
 *   dimensions): {omegaforce (default 10000), r0 (default 20)}; "watmeth" (mcmod_watmeth.f90 + watermethane.f90 wmrb /
 *   wmrb_grad: rigid-body water-methane site-site surface, x(3,17) = sites H H Q D D T T O | H H H H C M M M M in bohr, analytic
 *   gradient, V0 not subtracted): no parameters; "malon" (mcmod_malon.f90 + pes_malonaldehyde.f90 `pes`: Morse + 3549 Gaussians
 *   over the 36 distances of x(3,9) = C C O C O H H H H in bohr, analytic gradient and Hessian, tables from
 *   <data_dir>/malonaldehyde.tbl): no parameters; "ccpol8sf": {iemonomer (default 1), isurf (default 3; 1..10 select the surfaces of
 *   init_ccpol, main_CCpol-8sf.f:14-107: SAPT data file, Eckart or Radau embedding, potparts or potparts_old,
 *   with or without the CCpol-8s correction)}.
 * pimdk_pes_set_v0  = assignment to module variable V0 (pimd_par.f90:166, rpi_ser.f90:95)
 * pimdk_pes_eval    = V (function) and Vprime (subroutine) over a batch x(ndim,natom,nbatch);
 *                     v (nbatch) and/or grad (ndim,natom,nbatch) may be NULL.  grad = +dV/dx.
 * pimdk_pes_vprime_inplace = literal Vprime(x,grad) semantics: for ccpol8sf x is perturbed in
 *                     place and left where the reference leaves it (x+eps-2eps+eps,
 *                     mcmod_waterdimer_ccpol.f90:48-52). */
int pimdk_pes_select(const char* name, const double* pes_params, pimdk_int nparams);
int pimdk_pes_info(pimdk_int* ndim, pimdk_int* natom);
int pimdk_pes_set_v0(double v0);
int pimdk_pes_eval(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, const double* x, double* v, double* grad);
int pimdk_pes_vprime_inplace(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, double* x, double* grad);
int pimdk_pes_eval_dev(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, const double* x, double* v, double* grad);

/* ---- ring-polymer potential: instantonmod.f90:17-151 (UM, UMprime, UMforceenergy) ----------
 * x, g: (n,ndim,natom); a,b: (ndim,natom), used when fixedends != 0; f or g may be NULL.
 * This is the f/g evaluation `instanton` hands to setulb on task 'FG' (instantonmod.f90:748-765). */
int pimdk_um_forceenergy(pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* x, const double* a,
                         const double* b, const double* mass, double betan, pimdk_int fixedends, double* f, double* g);
/* The same for npoly independent ring polymers in one call: x, g (n,ndim,natom,npoly); b (ndim,natom,npoly), one end
 * point per polymer; a (ndim,natom) shared; f (npoly).  This is the batch of the solid-angle loop of `program rpi`
 * (rpi_par.f90:209-281: npoints**3 instanton optimisations that differ in the rotated end point only): every polymer's
 * values are bit-identical to the single-polymer call. */
int pimdk_um_forceenergy_batch(pimdk_int npoly, pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* x,
                               const double* a, const double* b, const double* mass, double betan, pimdk_int fixedends,
                               double* f, double* g);

/* ---- second derivatives and the fluctuation factor -------------------------------
 * pimdk_pes_hessian = Vdoubleprime(x, hess) of the selected plugin over a batch (mcmod_1d.f90:37-57: central
 *   difference, eps = 1e-4, of the analytic gradient; mcmod_2dtest.f90:63-86: the reference's closed form, restated
 *   literally — its four elements are assigned inside the loop over the wells, so only the last well contributes;
 *   mcmod_waterdimer_ccpol.f90:59-76: central difference, eps = 1e-5, of the finite-difference Vprime).
 *   x(ndim,natom,nbatch) is in/out: it is perturbed in place and keeps the reference's round-off drift.
 *   hess(ndim,natom,ndim,natom,nbatch): hess(i,j,:,:) = d grad(:,:) / d x(i,j).
 * pimdk_um_hessian  = UMhessian(x, singlewell, answer) (instantonmod.f90:155-217, no inithess):
 *   band(ndof+1, totdof), LAPACK lower band storage of the mass-weighted ring-polymer Hessian exactly as the
 *   reference fills it (spring coupling written for beads 2..n at row ndof+1; 2/betan**2 on every diagonal).
 * pimdk_detj        = detJ(x, etasquared, singlewell[, eigvecs]) (instantonmod.f90:782-827): all eigenvalues of
 *   that matrix in ascending order (the reference calls DSBEVD; here a dense FP64 eigensolver on the device),
 *   eigvecs(totdof,totdof) optional (NULL = jobz 'N'). */
int pimdk_pes_hessian(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, double* x, double* hess);
int pimdk_um_hessian(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                     pimdk_int singlewell, double* band);
int pimdk_detj(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
               pimdk_int singlewell, double* etasquared, double* eigvecs);
/* The `readhess` branch of init_path (verletmodule.f90:49-88) for ONE ring polymer x(n,ndim,natom) that already sits on
 * the path: totdof normals N(0, sqrt(1/beta)) (Philox stream 4, keyed by seed and traj_gid), eigenvectors of the
 * ring-polymer Hessian (detJ(x, etasquared, .false., interphess, eigvecs): the interpolated Hessian is accepted and
 * ignored by the reference, instantonmod.f90:813, so UMhessian calls Vdoubleprime on every bead), and
 *   x(i2,j2,k2) += sum over modes 2..totdof with etasquared >= 0 of sqrt(1/(etasquared mass(k2))) tempx eigvecs(idof2, mode),
 * idof2 = natom*(j2-1 + ndim*(i2-1)) + k2 as written there.  x in/out; etasquared (totdof) optional out. */
int pimdk_readhess_displace(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                            double beta, uint64_t seed, pimdk_int traj_gid, double* etasquared);

/* ---- module verletint -----------------------------------------------------------------------
 * pimdk_nm_setup = alloc_nm + the a,b-independent part of init_nm (verletmodule.f90:306-338):
 *   lam, beadmass, transmatrix.  beadvec (which depends on a and each trajectory's b) is formed
 *   inside the kernels from the same expression. */
int pimdk_nm_setup(pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* mass, double betan, double tau);
/* Read back the module arrays (any pointer may be NULL): transmatrix(n,n), lam(n), beadmass(natom,n). */
int pimdk_nm_get(double* transmatrix, double* lam, double* beadmass);
/* nmtransform_forward / nmtransform_backward (verletmodule.f90:254-286) over a batch of nvec
 * bead vectors v(n,nvec).  beadvec == NULL is the reference's bead=0; otherwise beadvec(n,nvec)
 * is subtracted after (forward) / added before (backward) the transform. */
int pimdk_nm_transform(pimdk_int forward, pimdk_int nvec, const double* vin, const double* beadvec, double* vout);

/* init_path (verletmodule.f90:32-119, readhess=.false.): beads on the spline at
 * (k-1)*xi/(n-1), momenta ~ N(0,sqrt(1/betan))*sqrt(beadmass) in normal-mode space, transformed
 * back.  path/splinepath: (npath,ndim,natom); xi: (ntraj); x,p: (n,ndim,natom,ntraj) out. */
int pimdk_init_path(pimdk_int ntraj, pimdk_int npath, const double* lampath, const double* path,
                    const double* splinepath, const double* xi, uint64_t seed, const pimdk_int* traj_gid, double* x,
                    double* p);

/* propagate_pimd_nm (thermostat 1, verletmodule.f90:190-250) / propagate_pimd_pile (thermostat 2,
 * :372-416) for ntraj independent ring polymers = the (lambda x repetition) task loop of
 * pimd_par.f90:321-381 in one call.
 *   x, p      (n,ndim,natom,ntraj) in/out
 *   a         (ndim,natom)           startpoint
 *   b, dbdl   (ndim,natom,ntraj)     endpoints(ii,:,:), gradpoints(ii,:,:)
 *   dHdr      (ntraj) out: mean over steps > imin of sum mass*(-x(n,:,:))*dbdl, i.e. what the
 *             reference returns before the driver divides by betan**2 (pimd_par.f90:379)
 *   Noutput   Andersen: mean collision interval (Poisson); PILE: unused (print cadence)
 *   seed, traj_gid: RNG contract: Philox4x32-10 keyed by seed, counter carries the
 *             global trajectory id so results do not depend on how trajectories are sharded.
 *             traj_gid == NULL means 0..ntraj-1.  The counter carries 32 bits of the id: ids outside
 *             0 .. 2^32-1 are rejected with PIMDK_EINVAL (they would alias RNG streams).
 * Requires pimdk_pes_select and pimdk_nm_setup with matching n, ndim, natom. */
int pimdk_propagate(pimdk_int thermostat, pimdk_int ntraj, double* x, double* p, const double* a, const double* b,
                    const double* dbdl, double dt, double gamma, pimdk_int NMC, pimdk_int imin, pimdk_int Noutput,
                    pimdk_int cayley, uint64_t seed, const pimdk_int* traj_gid, double* dHdr);
int pimdk_propagate_dev(pimdk_int thermostat, pimdk_int ntraj, double* x, double* p, const double* a, const double* b,
                        const double* dbdl, double dt, double gamma, pimdk_int NMC, pimdk_int imin, pimdk_int Noutput,
                        pimdk_int cayley, uint64_t seed, const pimdk_int* traj_gid, double* dHdr);
/* Restart state of module verletint: `restart`, `restartnmc` (verletmodule.f90:10; namelist pimd_par.f90:45,75).
 *   restart = 0/1  dHdr starts from zero (verletmodule.f90:200,388)
 *   restart = 2    the dHdr array handed to pimdk_propagate holds the running sums read from the restart
 *                  files (pimd_par.f90:356-370) and is continued; the mean divides by NMC + restartnmc - imin
 *                  (:247,413).  The RNG step counters continue at restartnmc + 1, so a run of N steps followed
 *                  by a restarted run of M steps draws the random numbers of one run of N + M steps.
 * pimdk_get_dhdr_sums returns the running (un-normalised) sums of the last propagate call: the dHdr that
 * write_restart stores (verletmodule.f90:162-185, called at :206, 246, 394, 412). */
int pimdk_set_restart(pimdk_int restart, pimdk_int restartnmc);
/* Andersen thermostat, a run cut into several propagate calls (restart = 1 writes its files every Noutput steps from
 * inside the loop, verletmodule.f90:205-207, without touching the collision clock `count` / `rkick(1)`, :199-202,208-234):
 * enable = 1 makes the following Andersen calls continue the per-trajectory clocks left by the previous call (same
 * number of trajectories; PIMDK_EINVAL otherwise) instead of starting at count = 0 with a new Poisson interval.
 * enable = 0 (default): every call starts its clocks like a fresh propagate_pimd_nm. */
int pimdk_set_andersen_carry(pimdk_int enable);
/* dHdrlimit of namelist MCDATA (pimd_par.f90:45, 88; verletmodule.f90:404-409, propagate_pimd_pile only): a step whose
 * estimator contribution has |contr| >= limit is not added to dHdr, and the ring polymer is re-initialised on the spot by
 * init_path(xi, ...) — beads back on the spline path at the trajectory's xi, fresh momenta (RNG stream 0 at that step).
 * The guard therefore needs what init_path needs: the spline path and xi(ntraj) of the trajectories of the following
 * propagate calls (same order, same count).  limit < 0 (the reference's default, -1) switches the guard off; the other
 * arguments are then ignored.  Andersen calls ignore the limit, like propagate_pimd_nm. */
int pimdk_set_dhdrlimit(double limit, pimdk_int npath, const double* lampath, const double* path, const double* splinepath,
                        pimdk_int ntraj, const double* xi);
int pimdk_get_dhdr_sums(pimdk_int ntraj, double* sums);
/* index (0-based, into the last call's batch) of the first trajectory that tripped the NaN trap, or -1 */
pimdk_int pimdk_last_nan_trajectory(void);
/* The host-buffer pimdk_propagate cuts large batches into chunks of whole trajectories and overlaps the copy-in of
 * chunk c+1 and the copy-out of chunk c-1 with the propagation of chunk c (two streams, two device buffers; pass
 * page-locked host arrays for the copies to be asynchronous).  Results do not depend on the chunking: they are keyed
 * by the global trajectory id.  ntraj_per_chunk = 0 (default): automatic (>= 128 MB of state per chunk, used from
 * three chunks on); > 0: that many trajectories per chunk. */
int pimdk_set_propagate_chunk(pimdk_int ntraj_per_chunk);

/* Per-lambda statistics of pimd_par.f90:397-409 for the local shard:
 * sums(3,nintegral) = {sum I, sum I**2, count} with I = dHdr/betan**2 and lambda index
 * traj_gid/nrep.  Summing `sums` over ranks (one NCCL all-reduce) and calling pimdk_ti_finish
 * reproduces the root's mean/variance/answer (pimd_par.f90:410-424). */
int pimdk_ti_partial_sums(pimdk_int ntraj, const double* dHdr, const pimdk_int* traj_gid, pimdk_int nrep,
                          pimdk_int nintegral, double betan, double* sums);
int pimdk_ti_finish(pimdk_int nintegral, const double* sums, const double* weights, double betan, double* mean,
                    double* var, double* deltaA, double* sigmaA, double* q_over_q0);
/* ---- multi-GPU: one rank (process) per GPU, independent trajectories sharded by global id (pimd_par.f90:109-110,
 * 281-295), ONE collective per run ------------------------------------------------------------------------------
 * The reference gathers every rank's integrands on the root (MPI_Gather, pimd_par.f90:389) and forms the per-lambda mean
 * and variance there (:397-409).  Here the library owns an NCCL communicator (bound at run time; PIMDK_NCCL_LIB names the
 * library if libnccl.so.2 is not on the loader path) and all-reduces {sum I, sum I**2, count} per lambda point, 3*nintegral
 * doubles, on the library stream.
 *   pimdk_comm_unique_id : rank 0 creates the 128-byte id (ncclGetUniqueId) and hands it to the other ranks by whatever
 *                          means the host program has (MPI_Bcast in the Fortran drivers, a file, a socket)
 *   pimdk_comm_init      : every rank, after pimdk_init on its own device (ncclCommInitRank); nranks = 1 needs no id/NCCL
 *   pimdk_ti_allreduce   : sums(3,nintegral) on the host (from pimdk_ti_partial_sums), summed over ranks in place
 *   pimdk_ti_reduce_dev  : the same for dHdr / traj_gid that are still on the device (dHdr as pimdk_propagate_dev left
 *                          it): per-lambda partial sums formed by a device kernel, all-reduced, copied to sums (host);
 *                          feed pimdk_ti_finish with the result.  Works with one rank too (no NCCL involved). */
#define PIMDK_UNIQUE_ID_BYTES 128
int pimdk_comm_unique_id(void* id);
int pimdk_comm_init(pimdk_int rank, pimdk_int nranks, const void* id);
int pimdk_comm_finalize(void);
int pimdk_comm_info(pimdk_int* rank, pimdk_int* nranks, pimdk_int* nccl_version);
int pimdk_ti_allreduce(pimdk_int nintegral, double* sums);
int pimdk_ti_reduce_dev(pimdk_int ntraj, const double* dHdr, const pimdk_int* traj_gid, pimdk_int nrep, pimdk_int nintegral,
                        double betan, double* sums);
/* gauleg (verletmodule.f90:124-160) */
int pimdk_gauleg(double x1, double x2, pimdk_int nintegral, double* x, double* w);

/* ---- measurement helpers ---------------------------------------------------------------------
 * pimdk_profile(1) makes propagate/pes_eval bracket each kernel family with CUDA events on the
 * library stream; pimdk_profile_get returns accumulated milliseconds and launch counts for
 * family = "pes", "gemm", "update", "estimator".  pimdk_fp64_peak measures the DFMA pipe. */
int pimdk_profile(pimdk_int enable);
int pimdk_profile_get(const char* family, double* ms, pimdk_int* launches);
int pimdk_profile_reset(void);
int pimdk_fp64_peak(double* tflops);
pimdk_int pimdk_launch_count(void);
/* GPU self-test: number of operands (of 9 x 2^28) for which the kernels' three-instruction division by a
 * small integer constant differs from IEEE division in any bit (must be 0). */
int pimdk_selftest_division(pimdk_int* mismatches);
/* GPU self-test: mismatches (of 4 x 2^30: two operand ranges for the division) between the kernels' branch-free IEEE division / square root
 * sequences and the compiler's built-in expansions (must be 0). */
int pimdk_selftest_fastmath(pimdk_int* mismatches);

/* GPU self-test: the shared math policy (pimdk_detmath.h) evaluated on the device for n host-supplied arguments;
 * kind 0 exp, 1 log, 2 sin, 3 cos, 4 acos, 5 tanh, 6 pow(x,-1.5), 7 pow(x,-3), 8 pow(x,0.66666666666666666).
 * tests/ compares the bits with the host form of the same header. */
int pimdk_selftest_math(pimdk_int kind, pimdk_int n, const double* x, double* y);

#ifdef __cplusplus
}
#endif
#endif /* PIMDK_H */
