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
