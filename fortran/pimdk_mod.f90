! pimdk_mod.f90 -- ISO_C_BINDING layer between the (unchanged) Fortran drivers of pimd-tunneling and
! libpimdk.so (include/pimdk.h).  SOURCE ONLY: this image has no Fortran compiler, so this file has not
! been compiled; it is deliberately thin and mechanical.  Build of the reference: ifort -i8 -r8
! (makefile:5), hence integer(c_int64_t) / real(c_double) everywhere.
!
!   module pimdk           raw interfaces, one per C entry point
!   module mcmod_mass      drop-in replacement of mcmod_waterdimer_ccpol.f90 / mcmod_1d.f90 / mcmod_2dtest.f90
!                          (same module name, same procedures; choose the PES with -DPIMDK_PES=...)
!   subroutine pimdk_propagate_tasks   what the task loop of pimd_par.f90:321-381 collapses to (incl. the restart cadence)
!   subroutine pimdk_ti_statistics     what the gather + root statistics of pimd_par.f90:383-424 collapse to (one NCCL all-reduce)
module pimdk
  use iso_c_binding
  implicit none
  interface
     integer(c_int) function pimdk_init(device, data_dir) bind(C, name="pimdk_init")
       import; integer(c_int64_t), value :: device; character(kind=c_char) :: data_dir(*)
     end function
     integer(c_int) function pimdk_finalize() bind(C, name="pimdk_finalize")
       import
     end function
     type(c_ptr) function pimdk_last_error() bind(C, name="pimdk_last_error")
       import
     end function
     integer(c_int) function pimdk_pes_select(name, params, nparams) bind(C, name="pimdk_pes_select")
       import; character(kind=c_char) :: name(*); type(c_ptr), value :: params; integer(c_int64_t), value :: nparams
     end function
     integer(c_int) function pimdk_pes_set_v0(v0) bind(C, name="pimdk_pes_set_v0")
       import; real(c_double), value :: v0
     end function
     integer(c_int) function pimdk_pes_eval(nbatch, ndim, natom, x, v, grad) bind(C, name="pimdk_pes_eval")
       import; integer(c_int64_t), value :: nbatch, ndim, natom
       real(c_double) :: x(*); type(c_ptr), value :: v, grad
     end function
     integer(c_int) function pimdk_pes_vprime_inplace(nbatch, ndim, natom, x, grad) bind(C, name="pimdk_pes_vprime_inplace")
       import; integer(c_int64_t), value :: nbatch, ndim, natom; real(c_double) :: x(*), grad(*)
     end function
     integer(c_int) function pimdk_um_forceenergy(n, ndim, natom, x, a, b, mass, betan, fixedends, f, g) &
          bind(C, name="pimdk_um_forceenergy")
       import; integer(c_int64_t), value :: n, ndim, natom, fixedends; real(c_double), value :: betan
       real(c_double) :: x(*), mass(*); type(c_ptr), value :: a, b, f, g
     end function
     integer(c_int) function pimdk_nm_setup(n, ndim, natom, mass, betan, tau) bind(C, name="pimdk_nm_setup")
       import; integer(c_int64_t), value :: n, ndim, natom; real(c_double) :: mass(*); real(c_double), value :: betan, tau
     end function
     integer(c_int) function pimdk_nm_get(transmatrix, lam, beadmass) bind(C, name="pimdk_nm_get")
       import; type(c_ptr), value :: transmatrix, lam, beadmass
     end function
     integer(c_int) function pimdk_init_path(ntraj, npath, lampath, path, splinepath, xi, seed, gid, x, p) &
          bind(C, name="pimdk_init_path")
       import; integer(c_int64_t), value :: ntraj, npath, seed; type(c_ptr), value :: gid
       real(c_double) :: lampath(*), path(*), splinepath(*), xi(*), x(*), p(*)
     end function
     integer(c_int) function pimdk_propagate(thermostat, ntraj, x, p, a, b, dbdl, dt, gamma, NMC, imin, Noutput, &
          cayley, seed, gid, dHdr) bind(C, name="pimdk_propagate")
       import; integer(c_int64_t), value :: thermostat, ntraj, NMC, imin, Noutput, cayley, seed
       real(c_double), value :: dt, gamma; type(c_ptr), value :: gid
       real(c_double) :: x(*), p(*), a(*), b(*), dbdl(*), dHdr(*)
     end function
     ! Vdoubleprime / UMhessian / detJ (mcmod_*.f90 Vdoubleprime; instantonmod.f90:155-217, 782-827)
     integer(c_int) function pimdk_pes_hessian(nbatch, ndim, natom, x, hess) bind(C, name="pimdk_pes_hessian")
       import; integer(c_int64_t), value :: nbatch, ndim, natom; real(c_double) :: x(*), hess(*)
     end function
     integer(c_int) function pimdk_um_hessian(n, ndim, natom, x, mass, betan, singlewell, band) bind(C, name="pimdk_um_hessian")
       import; integer(c_int64_t), value :: n, ndim, natom, singlewell; real(c_double), value :: betan
       real(c_double) :: x(*), mass(*), band(*)
     end function
     integer(c_int) function pimdk_detj(n, ndim, natom, x, mass, betan, singlewell, etasquared, eigvecs) bind(C, name="pimdk_detj")
       import; integer(c_int64_t), value :: n, ndim, natom, singlewell; real(c_double), value :: betan
       real(c_double) :: x(*), mass(*), etasquared(*); type(c_ptr), value :: eigvecs
     end function
     ! UMforceenergy for npoly independent ring polymers (the solid-angle loop of rpi_par.f90:209-281)
     integer(c_int) function pimdk_um_forceenergy_batch(npoly, n, ndim, natom, x, a, b, mass, betan, fixedends, f, g) &
          bind(C, name="pimdk_um_forceenergy_batch")
       import; integer(c_int64_t), value :: npoly, n, ndim, natom, fixedends; real(c_double), value :: betan
       real(c_double), intent(in) :: x(*), a(*), b(*), mass(*); real(c_double) :: f(*), g(*)
     end function
     ! the readhess branch of init_path (verletmodule.f90:49-88) for one ring polymer
     integer(c_int) function pimdk_readhess_displace(n, ndim, natom, x, mass, betan, beta, seed, traj_gid, etasquared) &
          bind(C, name="pimdk_readhess_displace")
       import; integer(c_int64_t), value :: n, ndim, natom, seed, traj_gid
       real(c_double) :: x(*), etasquared(*); real(c_double), intent(in) :: mass(*); real(c_double), value :: betan, beta
     end function
     ! module variables restart / restartnmc (verletmodule.f90:10) and the running sums write_restart stores (:171)
     integer(c_int) function pimdk_set_restart(restart, restartnmc) bind(C, name="pimdk_set_restart")
       import; integer(c_int64_t), value :: restart, restartnmc
     end function
     integer(c_int) function pimdk_get_dhdr_sums(ntraj, sums) bind(C, name="pimdk_get_dhdr_sums")
       import; integer(c_int64_t), value :: ntraj; real(c_double) :: sums(*)
     end function
     integer(c_int64_t) function pimdk_last_nan_trajectory() bind(C, name="pimdk_last_nan_trajectory")
       import
     end function
     ! trajectories per chunk of the copy-overlapped host-buffer propagate (0 = automatic)
     integer(c_int) function pimdk_set_propagate_chunk(ntraj_per_chunk) bind(C, name="pimdk_set_propagate_chunk")
       import; integer(c_int64_t), value :: ntraj_per_chunk
     end function
     ! Andersen collision clocks continue across the calls of one segmented run (verletmodule.f90:199-234 never resets them)
     integer(c_int) function pimdk_set_andersen_carry(enable) bind(C, name="pimdk_set_andersen_carry")
       import; integer(c_int64_t), value :: enable
     end function
     ! dHdrlimit (pimd_par.f90:45,88; verletmodule.f90:404-409): outlier guard of propagate_pimd_pile with re-initialisation
     integer(c_int) function pimdk_set_dhdrlimit(limit, npath, lampath, path, splinepath, ntraj, xi) bind(C, name="pimdk_set_dhdrlimit")
       import; real(c_double), value :: limit; integer(c_int64_t), value :: npath, ntraj
       real(c_double) :: lampath(*), path(*), splinepath(*), xi(*)
     end function
     ! multi-GPU: the library's NCCL communicator and the one collective of the path (replaces MPI_Gather, pimd_par.f90:389)
     integer(c_int) function pimdk_comm_unique_id(id) bind(C, name="pimdk_comm_unique_id")
       import; character(kind=c_char) :: id(*)
     end function
     integer(c_int) function pimdk_comm_init(rank, nranks, id) bind(C, name="pimdk_comm_init")
       import; integer(c_int64_t), value :: rank, nranks; character(kind=c_char) :: id(*)
     end function
     integer(c_int) function pimdk_comm_finalize() bind(C, name="pimdk_comm_finalize")
       import
     end function
     integer(c_int) function pimdk_ti_partial_sums(ntraj, dHdr, gid, nrep, nintegral, betan, sums) bind(C, name="pimdk_ti_partial_sums")
       import; integer(c_int64_t), value :: ntraj, nrep, nintegral; real(c_double), value :: betan
       real(c_double) :: dHdr(*), sums(*); type(c_ptr), value :: gid
     end function
     integer(c_int) function pimdk_ti_allreduce(nintegral, sums) bind(C, name="pimdk_ti_allreduce")
       import; integer(c_int64_t), value :: nintegral; real(c_double) :: sums(*)
     end function
     integer(c_int) function pimdk_ti_finish(nintegral, sums, weights, betan, mean, var, deltaA, sigmaA, q_over_q0) &
          bind(C, name="pimdk_ti_finish")
       import; integer(c_int64_t), value :: nintegral; real(c_double), value :: betan
       real(c_double) :: sums(*), weights(*), mean(*), var(*), deltaA(*), sigmaA(*), q_over_q0(*)
     end function
     integer(c_int) function pimdk_gauleg(x1, x2, n, x, w) bind(C, name="pimdk_gauleg")
       import; real(c_double), value :: x1, x2; integer(c_int64_t), value :: n; real(c_double) :: x(*), w(*)
     end function
  end interface
contains
  ! the reference's error convention: write a message and stop (verletmodule.f90:533-536,577-580)
  subroutine pimdk_check(rc)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (rc .eq. 0) return
    call c_f_pointer(pimdk_last_error(), msg, (/512/))
    i = 1
    do while (i .lt. 512 .and. msg(i) .ne. c_null_char)
       i = i + 1
    end do
    write(*,*) "pimdk: ", msg(1:i-1)
    if (rc .eq. 5) write(*,*) "NaN in pot propagation, trajectory", pimdk_last_nan_trajectory() + 1
    stop
  end subroutine pimdk_check
end module pimdk

!---------------------------------------------------------------------------------------------------------
! Replacement plugin: same module name and procedures as mcmod_waterdimer_ccpol.f90:1-80, in the CURRENT
! plugin interface (mcmod_waterdimer.f90:1-105): V_init(iproc), V, Vprime, potforce, Vdoubleprime, module variables.
module mcmod_mass
  use iso_c_binding
  use pimdk
  implicit none
  double precision::               V0, eps2=1.0d-5
  integer::                        atom1=1, atom2=2, atom3=3
  integer::                        n, ndim, ndof, natom, xunit, totdof
  logical::                        potforcepresent=.true.
  character, allocatable::         label(:)
  character(len=20)::              basename
  ! which in-repo surface this plugin stands for: "ccpol8sf" (mcmod_waterdimer_ccpol.f90), "2dtest" (mcmod_2dtest.f90),
  ! "1d" (mcmod_1d.f90) or "so2" (mcmod_so2.f90); set before V_init (the reference has one mcmod_<PES>.f90 per surface and picks at link time)
  character(len=16)::              pimdk_pes_name = "ccpol8sf"
contains
  subroutine V_init(iproc)
    integer, intent(in) :: iproc
    ! data files are opened from the CWD like the reference (main_CCpol-8sf.f:49,115)
    call pimdk_check(pimdk_init(-1_c_int64_t, "."//c_null_char))
    call pimdk_check(pimdk_pes_select(trim(pimdk_pes_name)//c_null_char, c_null_ptr, 0_c_int64_t))
    write(*,*) "Potential initializaton complete"
    V0=0.0d0
  end subroutine V_init

  function V(x)
    double precision :: V, x(:,:)
    double precision, target :: xc(ndim,natom), vv(1)
    xc(:,:) = x(:,:)                                   ! callers pass strided slices x(i,:,:)
    call pimdk_check(pimdk_pes_set_v0(V0))
    call pimdk_check(pimdk_pes_eval(1_c_int64_t, int(ndim,c_int64_t), int(natom,c_int64_t), xc, c_loc(vv), c_null_ptr))
    V = vv(1)
  end function V

  subroutine Vprime(x, grad)
    double precision :: x(:,:), grad(:,:)
    double precision :: xc(ndim,natom), gc(ndim,natom)
    xc(:,:) = x(:,:)
    call pimdk_check(pimdk_pes_vprime_inplace(1_c_int64_t, int(ndim,c_int64_t), int(natom,c_int64_t), xc, gc))
    x(:,:) = xc(:,:)                                    ! the reference leaves x+eps-2eps+eps behind (:48-52)
    grad(:,:) = gc(:,:)
  end subroutine Vprime

  subroutine potforce(x, grad, energy)
    double precision :: x(:,:), grad(:,:), energy
    double precision, target :: xc(ndim,natom), gc(ndim,natom), vv(1)
    xc(:,:) = x(:,:)
    call pimdk_check(pimdk_pes_eval(1_c_int64_t, int(ndim,c_int64_t), int(natom,c_int64_t), xc, c_loc(vv), c_loc(gc)))
    grad(:,:) = gc(:,:); energy = vv(1)
  end subroutine potforce

  ! Vdoubleprime (mcmod_waterdimer_ccpol.f90:60-77: central difference of Vprime, eps = 1d-5, x perturbed in place and
  ! left with the drift of 2*ndof nested Vprime calls): the 36 gradient passes run on the device in one call
  subroutine Vdoubleprime(x, hess)
    double precision :: x(:,:), hess(:,:,:,:)
    double precision :: xc(ndim,natom), hc(ndim,natom,ndim,natom)
    xc(:,:) = x(:,:)
    call pimdk_check(pimdk_pes_hessian(1_c_int64_t, int(ndim,c_int64_t), int(natom,c_int64_t), xc, hc))
    x(:,:) = xc(:,:)
    hess(:,:,:,:) = hc(:,:,:,:)
  end subroutine Vdoubleprime
end module mcmod_mass

!---------------------------------------------------------------------------------------------------------
! The task loop of pimd_par.f90:321-381 (init_nm / init_path / propagate_pimd_* per task) as ONE batched call.
! endpoints, gradpoints: (ncalcs, ndim, natom) as in pimd_par.f90:243-257; integrand(ncalcs) out.
! restart / restartnmc are module verletint's (verletmodule.f90:10):
!   restart = 0  one call of NMC steps
!   restart = 1  the reference writes restart_proc<iproc>_<ii>.xyz every Noutput steps from inside its loop
!                (verletmodule.f90:205-207, 394) and at the end (:246, 412): here the run is cut into calls of Noutput
!                steps, each continuing the previous one's running sums, RNG step counters (pimdk_set_restart(2, done))
!                and Andersen collision clocks (pimdk_set_andersen_carry), and the files are written between the calls
!   restart = 2  x, p, dHdr and restartnmc are read from the files first (pimd_par.f90:356-370), then as restart = 1
subroutine pimdk_propagate_tasks(thermostat, ncalcs, first_gid, iproc, startpoint, endpoints, gradpoints, xipoints, &
     lampath, path, splinepath, npath, mass, label, betan, tau, dt, gamma, NMC, imin, Noutput, cayley, seed, &
     restart, dHdrlimit, integrand)
  use iso_c_binding
  use pimdk
  use mcmod_mass, only: n, ndim, natom
  implicit none
  integer :: thermostat, ncalcs, first_gid, iproc, npath, NMC, imin, Noutput, seed, restart, ii, done, left, k, kmin, local
  logical :: cayley
  double precision :: startpoint(ndim,natom), endpoints(ncalcs,ndim,natom), gradpoints(ncalcs,ndim,natom)
  double precision :: xipoints(ncalcs), lampath(npath), path(npath,ndim,natom), splinepath(npath,ndim,natom)
  double precision :: mass(natom), betan, tau, dt, gamma, dHdrlimit, integrand(ncalcs)
  character :: label(natom)
  double precision, allocatable :: x(:,:,:,:), p(:,:,:,:), b(:,:,:), dbdl(:,:,:), dHdr(:), sums(:)
  integer(c_int64_t), allocatable, target :: gid(:)
  allocate(x(n,ndim,natom,ncalcs), p(n,ndim,natom,ncalcs), b(ndim,natom,ncalcs), dbdl(ndim,natom,ncalcs))
  allocate(dHdr(ncalcs), sums(ncalcs), gid(ncalcs))
  do ii = 1, ncalcs
     b(:,:,ii) = endpoints(ii,:,:); dbdl(:,:,ii) = gradpoints(ii,:,:); gid(ii) = first_gid + ii - 1
  end do
  call pimdk_check(pimdk_nm_setup(int(n,c_int64_t), int(ndim,c_int64_t), int(natom,c_int64_t), mass, betan, tau))
  ! outlier guard of propagate_pimd_pile (verletmodule.f90:404-409); dHdrlimit < 0 (default -1) switches it off
  call pimdk_check(pimdk_set_dhdrlimit(dHdrlimit, int(npath,c_int64_t), lampath, path, splinepath, int(ncalcs,c_int64_t), xipoints))
  done = 0
  sums(:) = 0.0d0
  if (restart .lt. 2) then
     call pimdk_check(pimdk_init_path(int(ncalcs,c_int64_t), int(npath,c_int64_t), lampath, path, splinepath, xipoints, &
          int(seed,c_int64_t), c_loc(gid), x, p))
  else
     do ii = 1, ncalcs
        call pimdk_read_restart(iproc, ii, x(:,:,:,ii), p(:,:,:,ii), sums(ii), done)
     end do
  end if
  if (restart .eq. 0) then
     call pimdk_check(pimdk_set_restart(0_c_int64_t, 0_c_int64_t))
     call pimdk_check(pimdk_propagate(int(thermostat,c_int64_t), int(ncalcs,c_int64_t), x, p, startpoint, b, dbdl, dt, gamma, &
          int(NMC,c_int64_t), int(imin,c_int64_t), int(Noutput,c_int64_t), merge(1_c_int64_t,0_c_int64_t,cayley), &
          int(seed,c_int64_t), c_loc(gid), dHdr))
     integrand(:) = dHdr(:)/(betan**2)                  ! pimd_par.f90:379
  else
     left = NMC; local = 0
     do while (left .gt. 0)
        k = min(max(1, Noutput), left)
        kmin = max(0, min(k, imin - local))             ! steps of this segment that lie before imin
        dHdr(:) = sums(:)
        if (restart .eq. 2 .or. local .gt. 0) then
           call pimdk_check(pimdk_set_restart(2_c_int64_t, int(done + local, c_int64_t)))
        else
           call pimdk_check(pimdk_set_restart(1_c_int64_t, 0_c_int64_t))
        end if
        call pimdk_check(pimdk_set_andersen_carry(merge(1_c_int64_t, 0_c_int64_t, thermostat .eq. 1 .and. local .gt. 0)))
        if (kmin .ge. k) then                            ! the whole segment is equilibration: propagate, keep the sums
           call pimdk_check(pimdk_propagate(int(thermostat,c_int64_t), int(ncalcs,c_int64_t), x, p, startpoint, b, dbdl, dt, &
                gamma, int(k,c_int64_t), int(k-1,c_int64_t), int(Noutput,c_int64_t), merge(1_c_int64_t,0_c_int64_t,cayley), &
                int(seed,c_int64_t), c_loc(gid), dHdr))
        else
           call pimdk_check(pimdk_propagate(int(thermostat,c_int64_t), int(ncalcs,c_int64_t), x, p, startpoint, b, dbdl, dt, &
                gamma, int(k,c_int64_t), int(kmin,c_int64_t), int(Noutput,c_int64_t), merge(1_c_int64_t,0_c_int64_t,cayley), &
                int(seed,c_int64_t), c_loc(gid), dHdr))
           call pimdk_check(pimdk_get_dhdr_sums(int(ncalcs,c_int64_t), sums))
        end if
        local = local + k; left = left - k
        do ii = 1, ncalcs                                ! write_restart (verletmodule.f90:162-185)
           call pimdk_write_restart(iproc, ii, x(:,:,:,ii), p(:,:,:,ii), done + local, sums(ii), label)
        end do
     end do
     call pimdk_check(pimdk_set_restart(0_c_int64_t, 0_c_int64_t))
     call pimdk_check(pimdk_set_andersen_carry(0_c_int64_t))
     integrand(:) = sums(:)/dble(NMC + done - imin)/(betan**2)      ! verletmodule.f90:247,413; pimd_par.f90:379
  end if
  deallocate(x, p, b, dbdl, dHdr, sums, gid)
end subroutine pimdk_propagate_tasks

! restart_proc<iproc>_<ii>.xyz in the layout of write_restart (verletmodule.f90:162-185) / the read side pimd_par.f90:356-370
subroutine pimdk_restart_name(iproc, ii, fname)
  implicit none
  integer :: iproc, ii
  character(len=64) :: fname, a, b
  write(a,'(I0)') iproc; write(b,'(I0)') ii
  fname = "restart_proc"//trim(a)//"_"//trim(b)//".xyz"
end subroutine pimdk_restart_name

subroutine pimdk_write_restart(iproc, ii, xprop, pprop, istep, dHdr, label)
  use mcmod_mass, only: n, ndim, natom
  implicit none
  integer :: iproc, ii, istep, i, j, u
  double precision :: xprop(n,ndim,natom), pprop(n,ndim,natom), dHdr
  character :: label(natom)
  character(len=64) :: fname
  call pimdk_restart_name(iproc, ii, fname)
  u = 600 + iproc
  open(u, file=trim(fname))
  do i = 1, n
     write(u,*) natom
     write(u,*) dHdr
     do j = 1, natom
        write(u,*) label(j), xprop(i,:,j)
     end do
  end do
  do i = 1, n
     write(u,*) natom
     write(u,*) istep
     do j = 1, natom
        write(u,*) label(j), pprop(i,:,j)
     end do
  end do
  close(u)
end subroutine pimdk_write_restart

subroutine pimdk_read_restart(iproc, ii, x, pinit, dHdr, restartnmc)
  use mcmod_mass, only: n, ndim, natom
  implicit none
  integer :: iproc, ii, restartnmc, i, j, u, dummyint
  double precision :: x(n,ndim,natom), pinit(n,ndim,natom), dHdr
  character :: dummychar
  character(len=64) :: fname
  call pimdk_restart_name(iproc, ii, fname)
  u = 600 + iproc
  open(u, file=trim(fname))
  do i = 1, n
     read(u,*) dummyint
     read(u,*) dHdr
     do j = 1, natom
        read(u,*) dummychar, x(i,:,j)
     end do
  end do
  do i = 1, n
     read(u,*) dummyint
     read(u,*) restartnmc
     do j = 1, natom
        read(u,*) dummychar, pinit(i,:,j)
     end do
  end do
  close(u)
end subroutine pimdk_read_restart

!---------------------------------------------------------------------------------------------------------
! What pimd_par.f90:383-424 collapses to: no MPI_Gather of every rank's integrands, but ONE all-reduce (NCCL, inside the
! library) of {sum I, sum I**2, count} per lambda point, after which EVERY rank holds the statistics.
! The communicator is created once after V_init: rank 0 calls pimdk_comm_unique_id(id), the id (128 bytes) is broadcast
! with the MPI the driver already has (call MPI_Bcast(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)), then every rank
! calls pimdk_comm_init(iproc, nproc, id).  MPI is then needed for nothing else on this path.
subroutine pimdk_ti_statistics(ncalcs, first_gid, integrand, nrep, nintegral, weights, betan, answer, sigmaA, finalI)
  use iso_c_binding
  use pimdk
  implicit none
  integer :: ncalcs, first_gid, nrep, nintegral, ii
  double precision :: integrand(ncalcs), weights(nintegral), betan, answer, sigmaA, finalI
  double precision :: sums(3,nintegral), mean(nintegral), var(nintegral), dA(1), sA(1), qq(1), dH(ncalcs)
  integer(c_int64_t), allocatable, target :: gid(:)
  allocate(gid(ncalcs))
  do ii = 1, ncalcs
     gid(ii) = first_gid + ii - 1
  end do
  dH(:) = integrand(:)*betan**2                         ! pimdk_ti_partial_sums divides by betan**2 like pimd_par.f90:379
  call pimdk_check(pimdk_ti_partial_sums(int(ncalcs,c_int64_t), dH, c_loc(gid), int(nrep,c_int64_t), int(nintegral,c_int64_t), &
       betan, sums))
  call pimdk_check(pimdk_ti_allreduce(int(nintegral,c_int64_t), sums))
  call pimdk_check(pimdk_ti_finish(int(nintegral,c_int64_t), sums, weights, betan, mean, var, dA, sA, qq))
  answer = dA(1); sigmaA = sA(1)**2; finalI = qq(1)   ! sigmaA as the variance the reference carries (:420-423 take its sqrt)
  deallocate(gid)
end subroutine pimdk_ti_statistics
