"""Host-side mirror of the reference's PES plugin interface `module mcmod_mass`
(current form: mcmod_waterdimer.f90:1-105; in-scope plugins mcmod_1d.f90, mcmod_2dtest.f90,
mcmod_waterdimer_ccpol.f90).  Same names and argument meaning: V_init, V, Vprime, potforce; module
variables n, ndim, natom, ndof, totdof, V0, eps2, potforcepresent.  Coordinates are (ndim, natom)
arrays like the Fortran x(:,:); batched variants take (ndim, natom, nbatch).  All arithmetic runs in
the CUDA library; nothing here computes a potential on the CPU."""
import numpy as np

from . import _lib
from ._lib import check, f64, hptr, lib

_SHAPES = {"1d": (1, 1), "2dtest": (2, 1), "so2": (2, 1), "watmeth": (3, 17), "malon": (3, 9), "ccpol8sf": (3, 6)}


class McmodMass:
    """One selected PES (= one linked mcmod_<PES>.o in the reference, makefile:105-290)."""

    eps2 = 1.0e-5             # pgtol handed to L-BFGS-B (instantonmod.f90:746)
    potforcepresent = True    # potforce is provided for every PES here
    atom1, atom2, atom3 = 1, 2, 3

    def __init__(self, name, params=None, n=0, isurf=None, iemonomer=None):
        """params: "1d" {Vheight, x0}; "2dtest" {a0, b0, rho0}; "so2" {omegaforce, r0} (mcmod_so2.f90); "ccpol8sf" {iemonomer, isurf} — the first two arguments
        of init_ccpol(isurf, iemon, iembedang, ixyz) (main_CCpol-8sf.f:1; the plugin calls init_ccpol(3,1,1,0)), also
        settable by keyword."""
        if name not in _SHAPES:
            raise ValueError("unknown PES %r (1d, 2dtest, so2, watmeth, malon, ccpol8sf)" % name)
        self.name = name
        if name == "ccpol8sf" and (isurf is not None or iemonomer is not None):
            p0 = [1.0, 3.0] if params is None else (list(np.asarray(params, dtype=np.float64)) + [3.0])[:2]
            params = [p0[0] if iemonomer is None else float(iemonomer), p0[1] if isurf is None else float(isurf)]
        self.params = None if params is None else np.asarray(params, dtype=np.float64)
        self.ndim, self.natom = _SHAPES[name]
        self.ndof = self.ndim * self.natom
        self.n = n
        self.totdof = n * self.ndof
        self.V0 = 0.0
        self.label = ["O", "H", "H", "O", "H", "H"] if name == "ccpol8sf" else ["X"] * self.natom
        if name == "watmeth":   # watermethane.f90:6-7
            self.label = list("HHQDDTTOHHHHCMMMM")
        if name == "malon":     # pes_malonaldehyde.f90:12-21; mcmod_malon.f90:5 aligns on atoms 1, 2, 4
            self.label = list("CCOCOHHHH")
            self.atom1, self.atom2, self.atom3 = 1, 2, 4
        self.basename = ""
        self._selected = False

    # The library holds ONE selected PES and one V0 per process (like the reference's single linked mcmod_<PES>.o).
    # Several McmodMass objects may be alive: the one used last owns the selection, and an object that finds another
    # owner re-selects its surface and pushes its V0 again before it computes anything.
    _owner = None

    def _select(self):
        p = self.params
        check(lib().pimdk_pes_select(self.name.encode(), hptr(p), 0 if p is None else p.size))
        if self.V0 != 0.0:
            check(lib().pimdk_pes_set_v0(float(self.V0)))
        McmodMass._owner = self

    # subroutine V_init(iproc)
    def V_init(self, iproc=0):
        _lib.ensure_init()
        self.V0 = 0.0
        self._select()
        self._selected = True
        return self

    def set_V0(self, v0):
        """assignment to the module variable V0 (pimd_par.f90:166)"""
        self._need()
        check(lib().pimdk_pes_set_v0(float(v0)))
        self.V0 = float(v0)

    def _need(self):
        if not self._selected:
            raise RuntimeError("V_init has not been called")
        if McmodMass._owner is not self or not _lib._initialised:
            _lib.ensure_init()
            self._select()

    # function V(x)
    def V(self, x):
        return float(self.V_batch(np.asarray(x, dtype=np.float64).reshape(self.ndim, self.natom, 1, order="F"))[0])

    # subroutine Vprime(x, grad)
    def Vprime(self, x, inplace=False):
        """Returns grad(ndim,natom) = +dV/dx.  inplace=True reproduces the reference's in-place
        finite-difference perturbation of x for ccpol8sf (mcmod_waterdimer_ccpol.f90:48-52)."""
        self._need()
        xb = f64(np.asarray(x, dtype=np.float64).reshape(self.ndim, self.natom, 1, order="F"))
        g = np.empty_like(xb)
        if inplace:
            check(lib().pimdk_pes_vprime_inplace(1, self.ndim, self.natom, hptr(xb), hptr(g)))
            np.asarray(x)[...] = xb[:, :, 0]
        else:
            check(lib().pimdk_pes_eval(1, self.ndim, self.natom, hptr(xb), None, hptr(g)))
        return g[:, :, 0]

    # subroutine potforce(x, grad, energy)
    def potforce(self, x):
        xb = np.asarray(x, dtype=np.float64).reshape(self.ndim, self.natom, 1, order="F")
        v, g = self.eval_batch(xb, energy=True, gradient=True)
        return g[:, :, 0], float(v[0])

    # subroutine Vdoubleprime(x, hess)
    def Vdoubleprime(self, x, inplace=False):
        """hess(ndim,natom,ndim,natom) with hess[i,j,:,:] = d grad / d x(i,j) (mcmod_1d.f90:37-57,
        mcmod_2dtest.f90:63-86, mcmod_waterdimer_ccpol.f90:59-76).  inplace=True leaves the reference's
        finite-difference drift in x."""
        xb = np.array(np.asarray(x, dtype=np.float64).reshape(self.ndim, self.natom, 1, order="F"), order="F")
        h = self.Vdoubleprime_batch(xb)
        if inplace:
            np.asarray(x)[...] = xb[:, :, 0]
        return h[..., 0]

    def Vdoubleprime_batch(self, x):
        """x(ndim,natom,nbatch), F-contiguous float64, updated in place; returns hess(ndim,natom,ndim,natom,nbatch)"""
        self._need()
        assert x.dtype == np.float64 and x.flags["F_CONTIGUOUS"] and x.shape[:2] == (self.ndim, self.natom)
        nb = x.shape[2]
        h = np.empty((self.ndim, self.natom, self.ndim, self.natom, nb), order="F")
        check(lib().pimdk_pes_hessian(nb, self.ndim, self.natom, hptr(x), hptr(h)))
        return h

    # batched forms --------------------------------------------------------------------------
    def eval_batch(self, x, energy=True, gradient=True):
        self._need()
        x = f64(x)
        assert x.ndim == 3 and x.shape[:2] == (self.ndim, self.natom), x.shape
        nb = x.shape[2]
        v = np.empty(nb, dtype=np.float64) if energy else None
        g = np.empty_like(x) if gradient else None
        check(lib().pimdk_pes_eval(nb, self.ndim, self.natom, hptr(x), hptr(v), hptr(g)))
        return v, g

    def V_batch(self, x):
        return self.eval_batch(x, energy=True, gradient=False)[0]

    def Vprime_batch(self, x):
        return self.eval_batch(x, energy=False, gradient=True)[1]

    def Vprime_batch_inplace(self, x):
        self._need()
        assert x.dtype == np.float64 and x.flags["F_CONTIGUOUS"] and x.shape[:2] == (self.ndim, self.natom)
        g = np.empty_like(x)
        check(lib().pimdk_pes_vprime_inplace(x.shape[2], self.ndim, self.natom, hptr(x), hptr(g)))
        return g
