"""Host-side mirror of the reference's propagation module `module verletint` (verletmodule.f90),
batched over independent ring polymers: the (lambda x repetition) task loop of pimd_par.f90:321-381
becomes one call.  Array shapes follow the Fortran: x, p are (n, ndim, natom[, ntraj]); a is
(ndim, natom); b, dbdl are (ndim, natom[, ntraj]).  Everything numerical runs in libpimdk.so."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, f64, hptr, lib

ANDERSEN, PILE = 1, 2


class VerletInt:
    def __init__(self, pes, n, mass, beta, tau=1.0, gamma=1.0, dt=1e-3, NMC=0, imin=0, Noutput=100000,
                 cayley=False, seed=0):
        # defaults = pimd_par.f90:60-88
        self.pes = pes
        self.n = int(n)
        self.ndim, self.natom, self.ndof = pes.ndim, pes.natom, pes.ndim * pes.natom
        self.mass = f64(np.asarray(mass, dtype=np.float64).reshape(self.natom))
        self.beta = float(beta)
        self.betan = self.beta / (self.n + 1)  # pimd_par.f90:94
        self.tau, self.gamma, self.dt = float(tau), float(gamma), float(dt)
        self.NMC, self.imin, self.Noutput = int(NMC), int(imin), int(Noutput)
        self.cayley = bool(cayley)
        self.seed = int(seed)
        # module variables restart / restartnmc (verletmodule.f90:10; namelist pimd_par.f90:45, default 0 :75):
        # 0 no restart files, 1 write them, 2 continue from them
        self.restart, self.restartnmc = 0, 0
        self.last_sums = None   # running (un-normalised) dHdr sums of the last propagate call
        # dHdrlimit (module verletint, verletmodule.f90:18; namelist default -1): set_dhdrlimit() arms the outlier guard
        self.dHdrlimit = -1.0
        self._reinit = None
        self._ready = False

    # alloc_nm (verletmodule.f90:350-368): the reference seeds MT19937 from the clock; here the
    # seed is explicit (RNG contract)
    def alloc_nm(self, iproc=0, seed=None):
        if seed is not None:
            self.seed = int(seed)
        return self

    # init_nm (verletmodule.f90:306-338), a,b-independent part; beadvec is formed in-kernel
    def init_nm(self):
        _lib.ensure_init()
        check(lib().pimdk_nm_setup(self.n, self.ndim, self.natom, hptr(self.mass), self.betan, self.tau))
        self._ready = True
        VerletInt._owner = self
        return self

    # the library holds one set of normal-mode tables per process (module verletint's lam, beadmass, transmatrix):
    # the VerletInt used last owns them, like McmodMass owns the PES selection
    _owner = None

    def _need(self):
        self.pes._need()
        if not self._ready or VerletInt._owner is not self:
            self.init_nm()

    @property
    def transmatrix(self):
        self._need()
        T = np.empty((self.n, self.n), order="F")
        check(lib().pimdk_nm_get(hptr(T), None, None))
        return T

    @property
    def lam(self):
        self._need()
        v = np.empty(self.n)
        check(lib().pimdk_nm_get(None, hptr(v), None))
        return v

    @property
    def beadmass(self):
        self._need()
        v = np.empty((self.natom, self.n), order="F")
        check(lib().pimdk_nm_get(None, None, hptr(v)))
        return v

    def beadvec(self, a, b):
        """beadvec(n, ndof) of init_nm (verletmodule.f90:328-333) for one (a, b) pair (host-side helper)."""
        n = self.n
        pi = 3.14159265358979
        a = np.asarray(a, dtype=np.float64).reshape(self.ndim, self.natom, order="F")
        b = np.asarray(b, dtype=np.float64).reshape(self.ndim, self.natom, order="F")
        lam = self.lam
        i = np.arange(1, n + 1, dtype=np.float64)
        out = np.empty((n, self.ndof), order="F")
        for k in range(self.natom):
            for j in range(self.ndim):
                v = a[j, k] * np.sin(i * pi / (n + 1)) + b[j, k] * np.sin(n * i * pi / (n + 1))
                v = v * np.sqrt(2.0 / (n + 1))
                out[:, k * self.ndim + j] = v / (lam * self.betan) ** 2
        return out

    # nmtransform_forward / nmtransform_backward (verletmodule.f90:254-286) over vectors v(n, nvec)
    def nmtransform_forward(self, v, beadvec=None):
        return self._transform(1, v, beadvec)

    def nmtransform_backward(self, q, beadvec=None):
        return self._transform(0, q, beadvec)

    def _transform(self, fwd, v, beadvec):
        self._need()
        v = f64(np.asarray(v, dtype=np.float64).reshape(self.n, -1, order="F"))
        bv = None if beadvec is None else f64(np.asarray(beadvec, dtype=np.float64).reshape(self.n, -1, order="F"))
        out = np.empty_like(v)
        check(lib().pimdk_nm_transform(fwd, v.shape[1], hptr(v), hptr(bv), hptr(out)))
        return out

    # init_path (verletmodule.f90:32-119)
    def init_path(self, xi, lampath, path, splinepath, traj_gid=None, readhess=False):
        """x on the spline path, momenta sampled in normal-mode space.  readhess=True adds the thermal displacement of
        the reference's `readhess` branch (verletmodule.f90:49-88), one ring polymer at a time (an eigen-decomposition
        of order n*ndof each)."""
        if readhess:
            x, p = self.init_path(xi, lampath, path, splinepath, traj_gid=traj_gid)
            gid = np.arange(x.shape[3], dtype=np.int64) if traj_gid is None else np.asarray(traj_gid, dtype=np.int64)
            for t in range(x.shape[3]):
                x[..., t] = self.readhess_displace(x[..., t], int(gid[t]))[0]
            return x, p
        self._need()
        xi = f64(np.atleast_1d(np.asarray(xi, dtype=np.float64)))
        ntraj = xi.size
        npath = len(lampath)
        path = f64(path, (npath, self.ndim, self.natom))
        spl = f64(splinepath, (npath, self.ndim, self.natom))
        lam = f64(np.asarray(lampath, dtype=np.float64))
        gid = None if traj_gid is None else np.ascontiguousarray(traj_gid, dtype=np.int64)
        x = np.empty((self.n, self.ndim, self.natom, ntraj), order="F")
        p = np.empty_like(x)
        check(lib().pimdk_init_path(ntraj, npath, hptr(lam), hptr(path), hptr(spl), hptr(xi), self.seed, hptr(gid),
                                    hptr(x), hptr(p)))
        return x, p

    def readhess_displace(self, x, traj_gid=0):
        """One ring polymer x(n,ndim,natom) displaced along the eigenvectors of its ring-polymer Hessian with
        N(0, 1/(beta eta^2 m)) amplitudes (verletmodule.f90:52-87).  Returns (x, etasquared)."""
        self._need()
        xw = np.array(f64(x, (self.n, self.ndim, self.natom)), order="F")
        eta = np.empty(self.n * self.ndim * self.natom)
        check(lib().pimdk_readhess_displace(self.n, self.ndim, self.natom, hptr(xw), hptr(self.mass), self.betan,
                                            self.beta, self.seed, int(traj_gid), hptr(eta)))
        return xw, eta

    def set_dhdrlimit(self, limit, xi=None, lampath=None, path=None, splinepath=None):
        """dHdrlimit of the namelist (verletmodule.f90:404-409): an over-limit contribution is dropped and the path
        re-initialised by init_path(xi, ...); needs the spline path and the xi of every trajectory of the next calls"""
        self.dHdrlimit = float(limit)
        self._reinit = None if limit < 0 else (f64(np.atleast_1d(np.asarray(xi, dtype=np.float64))), f64(np.asarray(lampath, dtype=np.float64)),
                                               f64(path), f64(splinepath))
        return self

    def _arm_guard(self, ntraj):
        if self.dHdrlimit >= 0.0:
            xi, lam, path, spl = self._reinit
            if xi.size != ntraj:
                raise ValueError("dHdrlimit: xi for %d trajectories, the call has %d" % (xi.size, ntraj))
            check(lib().pimdk_set_dhdrlimit(self.dHdrlimit, lam.size, hptr(lam), hptr(path), hptr(spl), ntraj, hptr(xi)))
        else:
            check(lib().pimdk_set_dhdrlimit(-1.0, 0, None, None, None, 0, None))

    def _propagate(self, thermostat, x, p, a, b, dbdl, traj_gid, dHdr0=None):
        self._need()
        x = np.asarray(x)
        single = x.ndim == 3
        shp4 = (self.n, self.ndim, self.natom, 1 if single else x.shape[3])
        ntraj = shp4[3]
        xw = f64(np.array(x, dtype=np.float64, order="F").reshape(shp4, order="F"))
        pw = f64(np.array(p, dtype=np.float64, order="F").reshape(shp4, order="F"))
        a = f64(np.asarray(a, dtype=np.float64).reshape(self.ndim, self.natom, order="F"))
        b = f64(np.asarray(b, dtype=np.float64).reshape((self.ndim, self.natom, ntraj), order="F"))
        dbdl = f64(np.asarray(dbdl, dtype=np.float64).reshape((self.ndim, self.natom, ntraj), order="F"))
        gid = None if traj_gid is None else np.ascontiguousarray(traj_gid, dtype=np.int64)
        check(lib().pimdk_set_restart(self.restart, self.restartnmc if self.restart == 2 else 0))
        self._arm_guard(ntraj)
        if self.restart == 2:    # dHdr as read from the restart files is continued (verletmodule.f90:200,388)
            if dHdr0 is None:
                raise ValueError("restart = 2 needs the running sums dHdr0 read from the restart files")
            dHdr = f64(np.array(np.atleast_1d(dHdr0), dtype=np.float64).reshape(ntraj))
        else:
            dHdr = np.zeros(ntraj)
        check(lib().pimdk_propagate(thermostat, ntraj, hptr(xw), hptr(pw), hptr(a), hptr(b), hptr(dbdl), self.dt,
                                    self.gamma, self.NMC, self.imin, self.Noutput, 1 if self.cayley else 0, self.seed,
                                    hptr(gid), hptr(dHdr)))
        sums = np.empty(ntraj)
        check(lib().pimdk_get_dhdr_sums(ntraj, hptr(sums)))
        self.last_sums = sums
        if single:
            return xw[..., 0], pw[..., 0], float(dHdr[0])
        return xw, pw, dHdr

    # propagate_pimd_pile (verletmodule.f90:372-416): returns (x, p, dHdr)
    def propagate_pimd_pile(self, x, p, a, b, dbdl, traj_gid=None, dHdr0=None):
        return self._propagate(PILE, x, p, a, b, dbdl, traj_gid, dHdr0)

    # propagate_pimd_nm (verletmodule.f90:190-250)
    def propagate_pimd_nm(self, x, p, a, b, dbdl, traj_gid=None, dHdr0=None):
        return self._propagate(ANDERSEN, x, p, a, b, dbdl, traj_gid, dHdr0)

    # ---- restart files (verletmodule.f90:162-185 write_restart; read side pimd_par.f90:332-370) -------------
    @staticmethod
    def restart_filename(iproc, ii):
        """"restart_proc<iproc>_<ii>.xyz" (pimd_par.f90:336-354; ii = 1-based task index of the rank)"""
        return "restart_proc%d_%d.xyz" % (iproc, ii)

    def write_restart(self, path, xprop, pprop, ii, dHdr):
        """One trajectory's snapshot in the reference's layout: for every bead `natom`, `dHdr` (the running sum),
        then `label x y z` per atom; then for every bead `natom`, `ii` (steps done), `label px py pz`.  The
        reference writes list-directed records; 17 significant digits are written here so that the file
        round-trips every bit.  (The reference hard-codes three coordinates per atom; surfaces with ndim < 3
        write the coordinates they have.)"""
        x = np.asarray(xprop).reshape(self.n, self.ndim, self.natom, order="F")
        p = np.asarray(pprop).reshape(self.n, self.ndim, self.natom, order="F")
        lab = list(self.pes.label)
        with open(path, "w") as f:
            for arr, second in ((x, "%.17e" % float(dHdr)), (p, "%d" % int(ii))):
                for i in range(self.n):
                    f.write(" %d\n %s\n" % (self.natom, second))
                    for j in range(self.natom):
                        f.write(" %s %s\n" % (lab[j], " ".join("%.17e" % arr[i, d, j] for d in range(self.ndim))))

    def read_restart(self, path):
        """-> x(n,ndim,natom), p(n,ndim,natom), dHdr (running sum), restartnmc   (pimd_par.f90:356-370)"""
        x = np.empty((self.n, self.ndim, self.natom), order="F")
        p = np.empty_like(x)
        def num(t):   # Fortran list-directed reals may carry d/D exponents; atom labels are never converted
            return float(t.replace("D", "E").replace("d", "e"))

        with open(path) as f:
            tok = f.read().split()
        pos, second = 0, [None, None]
        for which, arr in enumerate((x, p)):
            for i in range(self.n):
                pos += 1                      # dummyint
                second[which] = tok[pos]      # dHdr / restartnmc
                pos += 1
                for j in range(self.natom):
                    pos += 1                  # dummychar (the label: skipped, whatever letters it holds)
                    for d in range(self.ndim):
                        arr[i, d, j] = num(tok[pos])
                        pos += 1
        return x, p, num(second[0]), int(num(second[1]))

    def propagate_restartable(self, thermostat, x, p, a, b, dbdl, traj_gid=None, iproc=0, directory="."):
        """The reference's restart protocol around one batched propagate call (verletmodule.f90:199-206, 246,
        387-394, 412; pimd_par.f90:328-370).  restart = 1: start fresh, write every trajectory's file every
        Noutput steps and at the end; restart = 2: read the files first and continue.  Trajectory t of the batch
        is the rank's task ii = t + 1.  Returns (x, p, dHdr)."""
        import os

        x = np.array(x, dtype=np.float64, order="F")
        p = np.array(p, dtype=np.float64, order="F")
        ntraj = x.shape[3]
        files = [os.path.join(directory, self.restart_filename(iproc, t + 1)) for t in range(ntraj)]
        sums = np.zeros(ntraj)
        done = 0
        if self.restart == 2:
            for t in range(ntraj):
                x[..., t], p[..., t], sums[t], done_t = self.read_restart(files[t])
                done = done_t
        NMC, imin, restart0 = self.NMC, self.imin, self.restart
        seg = max(1, int(self.Noutput)) if restart0 > 0 else NMC
        left, local = NMC, 0
        try:
            dH = None
            while left > 0:
                k = min(seg, left)
                # one segment = a restarted run of k steps whose first `imin - local` steps are not sampled
                self.NMC, self.imin = k, min(k - 1, max(0, imin - local)) if imin - local < k else k - 1
                self.restart, self.restartnmc = (2, done + local) if (restart0 == 2 or local > 0) else (restart0, 0)
                # the reference writes its files from inside ONE loop and never resets the Andersen collision clock
                # (verletmodule.f90:199-234): segments after the first continue count / rkick of the previous call
                check(lib().pimdk_set_andersen_carry(1 if (thermostat == ANDERSEN and local > 0) else 0))
                if imin - local >= k:   # the whole segment lies before imin: propagate without sampling
                    keep = sums.copy()
                    x, p, _ = self._propagate(thermostat, x, p, a, b, dbdl, traj_gid, sums if self.restart == 2 else None)
                    sums = keep
                else:
                    x, p, dH = self._propagate(thermostat, x, p, a, b, dbdl, traj_gid, sums if self.restart == 2 else None)
                    sums = self.last_sums.copy()
                local += k
                left -= k
                if restart0 > 0:
                    for t in range(ntraj):
                        self.write_restart(files[t], x[..., t], p[..., t], done + local, sums[t])
        finally:
            self.NMC, self.imin, self.restart, self.restartnmc = NMC, imin, restart0, 0
            check(lib().pimdk_set_andersen_carry(0))
        return x, p, sums / float(NMC + done - imin)

    def propagate_dev(self, thermostat, ntraj, x_ptr, p_ptr, a_ptr, b_ptr, dbdl_ptr, gid_ptr, dHdr_ptr, NMC=None):
        """Device-resident form (pointers into HBM, e.g. torch tensors' data_ptr()); enqueues on the
        library stream and returns after the NaN/convergence flag has been read back."""
        self._need()
        check(lib().pimdk_propagate_dev(thermostat, ntraj, x_ptr, p_ptr, a_ptr, b_ptr, dbdl_ptr, self.dt, self.gamma,
                                        self.NMC if NMC is None else NMC, self.imin, self.Noutput,
                                        1 if self.cayley else 0, self.seed, gid_ptr, dHdr_ptr))

    # gauleg (verletmodule.f90:124-160)
    @staticmethod
    def gauleg(x1, x2, nintegral):
        x = np.empty(nintegral)
        w = np.empty(nintegral)
        check(lib().pimdk_gauleg(float(x1), float(x2), nintegral, hptr(x), hptr(w)))
        return x, w
