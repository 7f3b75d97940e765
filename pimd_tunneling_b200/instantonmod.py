"""Host-side mirror of the ring-polymer algebra on the hot path of `module instantonmod`:
UM, UMprime, UMforceenergy (instantonmod.f90:17-151).  x is (n, ndim, natom) like the Fortran
x(:,:,:); a, b are (ndim, natom).  The L-BFGS-B driver (`instanton`, instantonmod.f90:679-777)
stays on the host and calls UMforceenergy on task 'FG'."""
import ctypes

import numpy as np

from ._lib import check, f64, hptr, lib


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

    # ---- second derivatives ----------------------------------------------------------------
    def UMhessian(self, x, singlewell=False):
        """answer(ndof+1, totdof): the mass-weighted ring-polymer Hessian in LAPACK lower band storage, exactly as
        `call UMhessian(x, singlewell, answer)` fills it (instantonmod.f90:155-217).  x is not modified here (the
        reference's in-place finite-difference drift is applied to a copy)."""
        p = self.pes
        xw = np.array(f64(x, (self.n, p.ndim, p.natom)), order="F")
        band = np.empty((p.ndof + 1, self.n * p.ndof), order="F")
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
