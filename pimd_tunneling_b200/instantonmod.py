"""Host-side mirror of the ring-polymer algebra on the hot path of `module instantonmod`:
UM, UMprime, UMforceenergy (instantonmod.f90:17-151).  x is (n, ndim, natom) like the Fortran
x(:,:,:); a, b are (ndim, natom).  The L-BFGS-B driver (`instanton`, instantonmod.f90:679-777)
stays on the host and calls UMforceenergy on task 'FG'."""
import ctypes

import numpy as np

from ._lib import check, f64, hptr, lib
from .path import rotate_atoms  # noqa: F401  (instantonmod.f90:346-376; re-exported)


class InstantonMod:
    def __init__(self, pes, mass, beta, n, fixedends=True, rpi=True):
        self.pes = pes
        self.mass = f64(np.asarray(mass, dtype=np.float64).reshape(pes.natom))
        self.n = int(n)
        self.beta = float(beta)
        # rpi_ser.f90:45 uses beta/n, pimd_par.f90:94 uses beta/(n+1)
        self.betan = self.beta / self.n if rpi else self.beta / (self.n + 1)
        self.fixedends = bool(fixedends)

    def _call(self, x, a, b, want_f, want_g):
        p = self.pes
        p._need()
        x = f64(x, (self.n, p.ndim, p.natom))
        a = None if a is None else f64(np.asarray(a, dtype=np.float64).reshape(p.ndim, p.natom))
        b = None if b is None else f64(np.asarray(b, dtype=np.float64).reshape(p.ndim, p.natom))
        f = ctypes.c_double(0.0)
        g = np.empty_like(x) if want_g else None
        check(lib().pimdk_um_forceenergy(self.n, p.ndim, p.natom, hptr(x), hptr(a), hptr(b), hptr(self.mass),
                                         self.betan, 1 if self.fixedends else 0,
                                         ctypes.addressof(f) if want_f else None, hptr(g)))
        return f.value, g

    def UM(self, x, a=None, b=None):
        return self._call(x, a, b, True, False)[0]

    def UMprime(self, x, a=None, b=None):
        return self._call(x, a, b, False, True)[1]

    def UMforceenergy(self, x, a=None, b=None):
        """returns (answer, UM) like `call UMforceenergy(x, answer, UM, a, b)`"""
        f, g = self._call(x, a, b, True, True)
        return g, f

    def UMforceenergy_batch(self, X, a, B):
        """npoly ring polymers at once: X (n,ndim,natom,npoly), end points a (ndim,natom) shared and B
        (ndim,natom,npoly); returns (answer(n,ndim,natom,npoly), UM(npoly)).  Bit-identical, polymer by polymer,
        to UMforceenergy."""
        p = self.pes
        p._need()
        X = np.asarray(X, dtype=np.float64)
        npoly = X.shape[3]
        Xw = f64(np.array(X, order="F").reshape((self.n, p.ndim, p.natom, npoly), order="F"))
        a = f64(np.asarray(a, dtype=np.float64).reshape(p.ndim, p.natom))
        Bw = f64(np.array(B, dtype=np.float64, order="F").reshape((p.ndim, p.natom, npoly), order="F"))
        f = np.empty(npoly)
        g = np.empty_like(Xw)
        check(lib().pimdk_um_forceenergy_batch(npoly, self.n, p.ndim, p.natom, hptr(Xw), hptr(a), hptr(Bw), hptr(self.mass),
                                               self.betan, 1 if self.fixedends else 0, hptr(f), hptr(g)))
        return g, f

    def instanton(self, xtilde, a=None, b=None, m=8, factr=1e6, pgtol=None, maxls=40, maxiter=15000):
        """`call instanton(xtilde[, a, b])` (instantonmod.f90:679-777): L-BFGS-B (m = 8, factr = 1e6, pgtol = eps2, the
        reference's settings; scipy's implementation of lbfgsb.f) on the GPU action and gradient.  Returns the
        optimised xtilde(n,ndim,natom)."""
        from scipy.optimize import fmin_l_bfgs_b

        p = self.pes
        shape = (self.n, p.ndim, p.natom)
        x0 = np.asarray(xtilde, dtype=np.float64).reshape(shape, order="F")

        def fg(v):
            g_, f_ = self.UMforceenergy(v.reshape(shape, order="F"), a, b)
            return f_, g_.reshape(-1, order="F")

        xs, self.last_UM, self.last_info = fmin_l_bfgs_b(fg, x0.reshape(-1, order="F"), m=m, factr=factr,
                                                         pgtol=p.eps2 if pgtol is None else pgtol, maxls=maxls,
                                                         maxiter=maxiter)
        return np.asfortranarray(xs.reshape(shape, order="F"))

    def instanton_batch(self, xtilde0, well1, endpoints, m=8, factr=1e6, pgtol=None, maxls=40, maxiter=15000):
        """`call instanton(xtilderot, well1, endpoints(ii,:,:))` (instantonmod.f90:679-777) for every end point of
        the solid-angle loop (rpi_par.f90:252-255), run side by side.  Each optimisation is an ordinary L-BFGS-B run
        (scipy's implementation of lbfgsb.f, the reference's settings) in its own thread; the f/g requests of the
        threads are coalesced into ONE batched GPU call per round, so every optimisation sees exactly the values — and
        takes exactly the iterates — of a run on its own.  Returns (x(n,ndim,natom,npts), f(npts), info list)."""
        import threading

        from scipy.optimize import fmin_l_bfgs_b

        p = self.pes
        ends = np.asarray(endpoints, dtype=np.float64).reshape(-1, p.ndim, p.natom)
        npts = ends.shape[0]
        shape = (self.n, p.ndim, p.natom)
        x0 = np.asarray(xtilde0, dtype=np.float64).reshape(shape, order="F")
        pgtol = self.pes.eps2 if pgtol is None else pgtol
        cv = threading.Condition()
        pending, results = {}, {}
        active = [npts]
        out = [None] * npts

        def fg_of(i):
            def fg(v):
                with cv:
                    pending[i] = v.reshape(shape, order="F")
                    cv.notify_all()
                    while i not in results:
                        cv.wait()
                    f_, g_ = results.pop(i)
                return f_, g_
            return fg

        def worker(i):
            try:
                out[i] = fmin_l_bfgs_b(fg_of(i), x0.reshape(-1, order="F"), m=m, factr=factr, pgtol=pgtol, maxls=maxls,
                                       maxiter=maxiter)
            finally:
                with cv:
                    active[0] -= 1
                    cv.notify_all()

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(npts)]
        for t in threads:
            t.start()
        failure = None
        while True:
            with cv:
                while active[0] > 0 and len(pending) < active[0]:
                    cv.wait()
                if active[0] == 0:
                    break
                ids = sorted(pending)
                X = np.stack([pending.pop(i) for i in ids], axis=3)
            try:
                g_, f_ = self.UMforceenergy_batch(X, well1, np.moveaxis(ends[ids], 0, 2))
                res = {i: (float(f_[k]), g_[..., k].reshape(-1, order="F").copy()) for k, i in enumerate(ids)}
            except Exception as e:      # hand the error to the waiting optimisers instead of dead-locking them
                failure = e
                res = {i: (float("nan"), np.full(x0.size, np.nan)) for i in ids}
            with cv:
                results.update(res)
                cv.notify_all()
        for t in threads:
            t.join()
        if failure is not None:
            raise failure
        xs = np.stack([o[0].reshape(shape, order="F") for o in out], axis=3)
        return np.asfortranarray(xs), np.array([o[1] for o in out]), [o[2] for o in out]

    def angular_sweep(self, xtilde0, well1, well2, npoints, cutofftheta=2.0 * np.pi, cutoffphi=np.pi, lndetj0=None, N=None,
                      **lbfgs):
        """The solid-angle loop of `program rpi` (rpi_par.f90:209-300) for ndim = 3: end points = well2 rotated by
        Gauss-Legendre angles (rotate_atoms about axes 1, 3, 1), one instanton optimisation per end point (batched),
        fluctuation factor, kink action and I(beta) = tanh(omega N) (1 if omega N > 1) per point.  Returns a dict with
        theta, phi, eta, weight and Ibeta arrays in the order the reference writes angularI.dat."""
        from numpy.polynomial.legendre import leggauss

        p = self.pes
        if p.ndim != 3:
            raise ValueError("the solid-angle loop rotates three-dimensional atoms (rotate_atoms)")

        def gauleg(x1, x2, n):
            t, w = leggauss(n)
            return 0.5 * (x2 - x1) * t + 0.5 * (x2 + x1), 0.5 * (x2 - x1) * w

        th, wth = gauleg(0.0, min(2.0 * np.pi, cutofftheta), npoints)
        ph, wph = gauleg(0.0, min(np.pi, cutoffphi), npoints)
        et, wet = gauleg(0.0, min(2.0 * np.pi, cutofftheta), npoints)
        ends, rows = [], []
        for ii in range(npoints):
            for jj in range(npoints):
                for kk in range(npoints):
                    w = np.array(well2, dtype=np.float64).reshape(3, p.natom).copy()
                    for axis, ang in ((1, et[kk]), (3, ph[jj]), (1, th[ii])):
                        w = rotate_atoms(w, axis, ang)
                    ends.append(w)
                    rows.append((th[ii], ph[jj], et[kk], wth[ii] * wph[jj] * wet[kk]))
        ends = np.array(ends)
        xs, fs, _ = self.instanton_batch(xtilde0, well1, ends, **lbfgs)
        if lndetj0 is None:
            xharm = np.empty((self.n, p.ndim, p.natom), order="F")
            xharm[:] = np.asarray(well1, dtype=np.float64).reshape(1, p.ndim, p.natom)
            e0 = self.detJ(xharm, singlewell=True)
            lndetj0 = float(np.sum(np.log(e0[e0 > 0.0])))
        N = self.n if N is None else N
        ib = np.empty(len(rows))
        for i in range(len(rows)):
            eta2 = self.detJ(xs[..., i], singlewell=False)[1:]
            lndetj = float(np.sum(np.log(eta2[eta2 > 0.0])))
            gam = np.exp(0.5 * (lndetj - lndetj0))
            sk = self.betan * fs[i]
            om = self.betan * np.exp(-sk) * np.sqrt(sk / (2.0 * np.pi)) / gam
            ib[i] = 1.0 if om * N > 1.0 else np.tanh(om * N)
        r = np.array(rows)
        return {"theta": r[:, 0], "phi": r[:, 1], "eta": r[:, 2], "weight": r[:, 3], "Ibeta": ib, "x": xs, "UM": fs}

    # ---- second derivatives ----------------------------------------------------------------
    def UMhessian(self, x, singlewell=False):
        """answer(ndof+1, totdof): the mass-weighted ring-polymer Hessian in LAPACK lower band storage, exactly as
        `call UMhessian(x, singlewell, answer)` fills it (instantonmod.f90:155-217).  x is not modified here (the
        reference's in-place finite-difference drift is applied to a copy)."""
        p = self.pes
        xw = np.array(f64(x, (self.n, p.ndim, p.natom)), order="F")
        band = np.empty((p.ndof + 1, self.n * p.ndof), order="F")
        p._need()
        check(lib().pimdk_um_hessian(self.n, p.ndim, p.natom, hptr(xw), hptr(self.mass), self.betan,
                                     1 if singlewell else 0, hptr(band)))
        return band

    def detJ(self, x, singlewell=False, eigvecs=False):
        """etasquared(totdof) [, eigvecs(totdof,totdof)] like `call detJ(x, etasquared, singlewell[, eigvecs=...])`
        (instantonmod.f90:782-827): eigenvalues of the UMhessian matrix, ascending."""
        p = self.pes
        xw = np.array(f64(x, (self.n, p.ndim, p.natom)), order="F")
        N = self.n * p.ndof
        eta = np.empty(N)
        z = np.empty((N, N), order="F") if eigvecs else None
        p._need()
        check(lib().pimdk_detj(self.n, p.ndim, p.natom, hptr(xw), hptr(self.mass), self.betan, 1 if singlewell else 0,
                               hptr(eta), hptr(z)))
        return (eta, z) if eigvecs else eta

    def rpi_splitting(self, xtilde, well1, well2):
        """The closing section of `program rpi` (rpi_ser.f90:221-236, 350-381): fluctuation factor from the two
        determinants (harmonic well with singlewell=.true.; instanton, first eigenvalue skipped), kink action and
        the ring-polymer-instanton splitting.  Returns a dict with lndetj0, lndetj, phi (gammetilde), s_kink,
        theta (omega), delta (Hartree)."""
        p = self.pes
        xharm = np.empty((self.n, p.ndim, p.natom), order="F")
        xharm[:] = np.asarray(well1, dtype=np.float64).reshape(1, p.ndim, p.natom)
        eta0 = self.detJ(xharm, singlewell=True)
        lndetj0 = float(np.sum(np.log(eta0[eta0 > 0.0])))
        eta = self.detJ(xtilde, singlewell=False)
        tail = eta[1:]
        lndetj = float(np.sum(np.log(tail[tail > 0.0])))
        phi = float(np.exp(0.5 * (lndetj - lndetj0)))
        um = self.UM(xtilde, well1, well2) if self.fixedends else self.UM(xtilde)
        s_kink = self.betan * um
        theta = self.betan * np.exp(-s_kink) * np.sqrt(s_kink / (2.0 * 3.14159265358979)) / phi
        return {"lndetj0": lndetj0, "lndetj": lndetj, "skipped0": int(np.sum(eta0 <= 0.0)), "skipped": int(np.sum(tail <= 0.0)),
                "phi": phi, "s_kink": float(s_kink), "theta": float(theta), "delta": float(2.0 * theta / self.betan)}

