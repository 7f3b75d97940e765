"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DATA_DIR = os.path.join(ROOT, "pimd_tunneling_b200", "data")
REF_DATA = "/root/reference/data_ccpol"

SAPT_FOR_SURF = {1: "SAPT5spf_2014", 2: "SAPT5spfIR_2014", 3: "SAPT5spfIR_2006", 4: "SAPT5spfIR_2006",
                 5: "SAPT5spf_2014", 6: "SAPT5spfIR_2014", 7: "SAPT5spfIR_2006", 8: "SAPT5spf_2006",
                 9: "SAPT5spfIR_2006", 10: "SAPT5spfIR_2014"}

# known-answer geometry and energies of test_parameters (main_CCpol-8sf.f:180-189), Angstrom / kcal/mol
GOLDEN_GEOM_ANG = np.array([
    0.6458557220e-01, 0.3399054992e-02, -0.1782922818e-01,
    -0.5505396894e+00, 0.6383283738e-01, -0.6241648475e+00,
    -0.4744802670e+00, -0.1177783097e+00, 0.9071276528e+00,
    -0.5658499752e-01, -0.1827211353e-03, 0.2523332441e+01,
    0.6334615998e+00, 0.2269642816e+00, 0.2055825014e+01,
    0.2645834235e+00, -0.2240643644e+00, 0.3293430563e+01])
GOLDEN_VAL = [-1.45754, -1.36660, -1.34676, -1.34688, -0.76978, -0.67884, -0.65912, -0.75268, -0.65451, -1.36648]
GOLDEN_VALM = [25.04582, 25.13676, 25.15660, 25.15648, 25.73358, 25.82452, 25.84424, 25.75068, 25.84885, 25.13688]

_P = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def build():
    """make the oracle library (no-op when up to date); serialised across processes by a file lock because
    bench.py's CPU legs start one worker per core and each of them lands here"""
    import fcntl

    so = os.path.join(ORACLE_DIR, "liboracle.so")
    with open(os.path.join(ORACLE_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            # -march=native: a library built on another host (it travels with the repo snapshot) is rebuilt when
            # this host's CPU feature flags differ (oracle/.cpu_stamp is a prerequisite in the Makefile)
            import hashlib

            flags = ""
            try:
                with open("/proc/cpuinfo") as f:
                    flags = next((ln for ln in f if ln.startswith("flags")), "")
            except OSError:
                pass
            stamp, h = os.path.join(ORACLE_DIR, ".cpu_stamp"), hashlib.sha1(flags.encode()).hexdigest()
            if not os.path.exists(stamp) or open(stamp).read().strip() != h:
                with open(stamp, "w") as f:
                    f.write(h + "\n")
            subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return so


class Oracle:
    def __init__(self):
        self.L = ctypes.CDLL(build())
        L = self.L
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_V.restype = ctypes.c_double
        L.orc_UM.restype = ctypes.c_double
        L.orc_UMforceenergy.restype = ctypes.c_double
        L.orc_splint.restype = ctypes.c_double
        L.orc_splin_grad.restype = ctypes.c_double
        L.orc_normal.restype = ctypes.c_double
        L.orc_opcount_name.restype = ctypes.c_char_p
        L.orc_ccpol_tables_image.restype = ctypes.c_long
        L.orc_nm_setup.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.orc_step_nm.argtypes = [ctypes.c_double, _P, _P]
        L.orc_step_v.argtypes = [ctypes.c_double, _P, _P]
        L.orc_step_langevin.argtypes = [_P, ctypes.c_ulonglong]
        L.orc_set_rng.argtypes = [ctypes.c_ulonglong, ctypes.c_uint]
        L.orc_propagate.argtypes = [ctypes.c_int, _P, _P, _P, ctypes.c_long, ctypes.c_long, ctypes.c_long, _P]
        L.orc_pes_eval.argtypes = [ctypes.c_long, _P, _P, _P]
        L.orc_init_path.argtypes = [ctypes.c_double, _P, _P, _P, ctypes.c_int, _P, _P]
        L.orc_gauleg.argtypes = [ctypes.c_double, ctypes.c_double, _P, _P, ctypes.c_int]
        L.orc_spline.argtypes = [_P, _P, ctypes.c_int, ctypes.c_double, ctypes.c_double, _P]
        L.orc_splint.argtypes = [_P, _P, _P, ctypes.c_int, ctypes.c_double]
        L.orc_splin_grad.argtypes = [_P, _P, _P, ctypes.c_int, ctypes.c_double]
        L.orc_normal.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_ulonglong]
        L.orc_poisson.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_double]
        L.orc_pots.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, _P]
        L.orc_sample_momenta.argtypes = [_P, ctypes.c_int, ctypes.c_ulonglong]
        L.orc_pes_set_V0.argtypes = [ctypes.c_double]
        L.orc_nm_forward.argtypes = [_P, _P, ctypes.c_int]
        L.orc_nm_backward.argtypes = [_P, _P, ctypes.c_int]
        self.ndim = self.natom = 0

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    # ---- tables / PES ---------------------------------------------------------------------
    def load_ccpol(self, isurf=3, iemon=1, text_dir=None):
        if text_dir:
            self._chk(self.L.orc_ccpol_load_text(text_dir.encode(), isurf, iemon))
        else:
            sapt = os.path.join(DATA_DIR, "sapt_%s.tbl" % SAPT_FOR_SURF[isurf])
            cc = os.path.join(DATA_DIR, "ccpol8s.tbl")
            self._chk(self.L.orc_ccpol_load_packed(sapt.encode(), cc.encode(), isurf, iemon))

    def tables_image(self):
        sz = self.L.orc_ccpol_tables_image(None, 0)
        buf = (ctypes.c_ubyte * sz)()
        self.L.orc_ccpol_tables_image(buf, sz)
        return bytes(buf)

    def ccpol_energy_ang(self, xyz18):
        e = ctypes.c_double()
        x = np.ascontiguousarray(xyz18, dtype=np.float64)
        self._chk(self.L.orc_ccpol_energy_ang(_p(x), ctypes.byref(e)))
        return e.value

    def ccpol_analytic_gradient(self, x):
        """x(3,6) bohr -> (V Hartree without V0, grad(3,6) Hartree/bohr): forward-mode dual numbers through the oracle's
        templates (oracle/dual.hpp) — the analytic gradient the reference does not have"""
        xk = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(3, 6).T.reshape(-1))
        v = ctypes.c_double()
        g = np.zeros(18)
        self._chk(self.L.orc_ccpol_analytic_gradient(_p(xk), ctypes.byref(v), _p(g)))
        return v.value, g.reshape(6, 3).T.copy()

    def ccpol_opcount(self, xyz18):
        nk = self.L.orc_opcount_kinds()
        cnt = np.zeros(nk)
        e = ctypes.c_double()
        x = np.ascontiguousarray(xyz18, dtype=np.float64)
        self._chk(self.L.orc_ccpol_opcount(_p(x), _p(cnt), nk, ctypes.byref(e)))
        return {self.L.orc_opcount_name(i).decode(): int(cnt[i]) for i in range(nk)}, e.value

    def select(self, name, ndim=None, natom=None):
        if name == "ccpol8sf":
            self.load_ccpol()
        if name == "malon":
            self._chk(self.L.orc_malon_load(os.path.join(DATA_DIR, "malonaldehyde.tbl").encode()))
        self._chk(self.L.orc_pes_select(name.encode()))
        shapes = {"1d": (1, 1), "2dtest": (2, 1), "so2": (2, 1), "watmeth": (3, 17), "malon": (3, 9), "ccpol8sf": (3, 6)}
        self.ndim, self.natom = shapes[name]
        if ndim:
            self.ndim, self.natom = ndim, natom
            self.L.orc_pes_set_dims(ndim, natom)
        return self

    def set_V0(self, v0):
        self.L.orc_pes_set_V0(float(v0))

    def set_dhdrlimit(self, limit, xi=0.0, lampath=None, path=None, splinepath=None):
        """dHdrlimit and the path init_path re-initialises from (verletmodule.f90:404-409); call after nm_setup"""
        self.L.orc_set_dhdrlimit.argtypes = [ctypes.c_double, ctypes.c_double, _P, _P, _P, ctypes.c_int]
        if limit < 0:
            self.L.orc_set_dhdrlimit(float(limit), 0.0, None, None, None, 0)
            return
        lam = np.ascontiguousarray(lampath, dtype=np.float64)
        pth = np.asfortranarray(path, dtype=np.float64)
        spl = np.asfortranarray(splinepath, dtype=np.float64)
        self.L.orc_set_dhdrlimit(float(limit), float(xi), _p(lam), _p(pth), _p(spl), lam.size)

    def set_so2(self, omegaforce, r0):
        self.L.orc_pes_set_so2.argtypes = [ctypes.c_double, ctypes.c_double]
        self.L.orc_pes_set_so2(float(omegaforce), float(r0))

    def pes_eval(self, x, energy=True, gradient=True):
        """x(ndim,natom,nbatch) F-order; returns (v, grad, x_after) — x_after carries the FD drift."""
        x = np.array(x, dtype=np.float64, order="F")
        nb = x.shape[2]
        v = np.empty(nb) if energy else None
        g = np.empty_like(x) if gradient else None
        self._chk(self.L.orc_pes_eval(nb, _p(x), _p(v), _p(g)))
        return v, g, x

    # ---- verletint ------------------------------------------------------------------------
    def nm_setup(self, n, mass, betan, tau=1.0, gamma=1.0, dt=1e-3, cayley=False, fixedends=True):
        m = np.ascontiguousarray(mass, dtype=np.float64)
        self.n = n
        self._chk(self.L.orc_nm_setup(n, self.ndim, self.natom, _p(m), betan, tau, gamma, dt, int(cayley), int(fixedends)))

    def init_nm(self, a, b):
        a = np.array(a, dtype=np.float64, order="F")
        b = np.array(b, dtype=np.float64, order="F")
        self._chk(self.L.orc_init_nm(_p(a), _p(b)))

    def get_nm(self):
        n, nd = self.n, self.ndim * self.natom
        T = np.empty((n, n), order="F")
        lam = np.empty(n)
        bm = np.empty((self.natom, n), order="F")
        bv = np.empty((n, nd), order="F")
        self.L.orc_get_nm(_p(T), _p(lam), _p(bm), _p(bv))
        return T, lam, bm, bv

    def set_rng(self, seed, gid):
        self.L.orc_set_rng(seed, gid)

    def propagate(self, thermostat, x, p, dbdl, NMC, imin=0, Noutput=100000):
        x = np.array(x, dtype=np.float64, order="F")
        p = np.array(p, dtype=np.float64, order="F")
        d = np.array(dbdl, dtype=np.float64, order="F")
        out = ctypes.c_double()
        self._chk(self.L.orc_propagate(thermostat, _p(x), _p(p), _p(d), NMC, imin, Noutput, ctypes.byref(out)))
        return x, p, out.value

    def init_path(self, xi, lampath, path, spl):
        n, nd, na = self.n, self.ndim, self.natom
        lam = np.ascontiguousarray(lampath, dtype=np.float64)
        path = np.array(path, dtype=np.float64, order="F")
        spl = np.array(spl, dtype=np.float64, order="F")
        x = np.empty((n, nd, na), order="F")
        p = np.empty((n, nd, na), order="F")
        self._chk(self.L.orc_init_path(xi, _p(lam), _p(path), _p(spl), len(lam), _p(x), _p(p)))
        return x, p

    def step_nm(self, time, x, p):
        x = np.array(x, dtype=np.float64, order="F")
        p = np.array(p, dtype=np.float64, order="F")
        self.L.orc_step_nm(time, _p(x), _p(p))
        return x, p

    def UM(self, x, a, b):
        x = np.array(x, dtype=np.float64, order="F")
        a = np.array(a, dtype=np.float64, order="F")
        b = np.array(b, dtype=np.float64, order="F")
        return self.L.orc_UM(_p(x), _p(a), _p(b))

    def UMprime(self, x, a, b):
        x = np.array(x, dtype=np.float64, order="F")
        a = np.array(a, dtype=np.float64, order="F")
        b = np.array(b, dtype=np.float64, order="F")
        g = np.empty_like(x)
        self.L.orc_UMprime(_p(x), _p(g), _p(a), _p(b))
        return g

    def UMforceenergy(self, x, a, b):
        x = np.array(x, dtype=np.float64, order="F")
        a = np.array(a, dtype=np.float64, order="F")
        b = np.array(b, dtype=np.float64, order="F")
        g = np.empty_like(x)
        f = self.L.orc_UMforceenergy(_p(x), _p(g), _p(a), _p(b))
        return g, f

    def Vdoubleprime(self, x):
        """-> hess(ndim,natom,ndim,natom), x after the in-place perturbation"""
        x = np.array(x, dtype=np.float64, order="F")
        h = np.empty((self.ndim, self.natom, self.ndim, self.natom), order="F")
        self.L.orc_Vdoubleprime(_p(x), _p(h))
        return h, x

    def UMhessian(self, x, singlewell=False):
        """-> answer(ndof+1, totdof) (needs nm_setup for n, mass, betan)"""
        x = np.array(x, dtype=np.float64, order="F")
        n = x.shape[0]
        ndof = self.ndim * self.natom
        band = np.empty((ndof + 1, n * ndof), order="F")
        self.L.orc_UMhessian(_p(x), int(singlewell), _p(band))
        return band

    def gauleg(self, x1, x2, n):
        x = np.empty(n)
        w = np.empty(n)
        self.L.orc_gauleg(x1, x2, _p(x), _p(w), n)
        return x, w

    def spline(self, x, y, yp1=1e31, ypn=1e31):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        y2 = np.empty_like(x)
        self._chk(self.L.orc_spline(_p(x), _p(y), len(x), yp1, ypn, _p(y2)))
        return y2

    def splint(self, xa, ya, y2a, x):
        return self.L.orc_splint(_p(np.ascontiguousarray(xa)), _p(np.ascontiguousarray(ya)),
                                 _p(np.ascontiguousarray(y2a)), len(xa), float(x))

    def splin_grad(self, xa, ya, y2a, x):
        return self.L.orc_splin_grad(_p(np.ascontiguousarray(xa)), _p(np.ascontiguousarray(ya)),
                                     _p(np.ascontiguousarray(y2a)), len(xa), float(x))


# rigid-body site coordinates of the water-methane surface as its header gives them (watermethane.f90:9-28), bohr:
# water H H Q D D T T O, methane H H H H C M M M M (each monomer about its own origin)
WATMETH_WATER = np.array([[0.0, 1.45365, -1.12169], [0.0, -1.45365, -1.12169], [0.0, 0.0, -0.04490], [0.0, 0.70785, 0.34527],
                          [0.0, -0.70785, 0.34527], [0.60787, 0.0, 0.35218], [-0.60787, 0.0, 0.35218], [0.0, 0.0, 0.0]])
WATMETH_METHANE = np.array([[0.0, 0.0, 2.07704], [1.95825, 0.0, -0.69235], [-0.97913, 1.69590, -0.69235], [-0.97913, -1.69590, -0.69235],
                            [0.0, 0.0, 0.0], [0.0, 0.0, 1.03852], [0.97913, 0.0, -0.34617], [-0.48956, 0.84795, -0.34617],
                            [-0.48956, -0.84795, -0.34617]])


def watmeth_geometries(nbatch, seed=0, rmin=5.5, rmax=9.0):
    """x(3,17,nbatch): the two rigid bodies, each randomly rotated, centres rmin..rmax bohr apart along a random direction"""
    rng = np.random.default_rng(seed)

    def rot():
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        a, b, c, d = q
        return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                         [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                         [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])

    x = np.empty((3, 17, nbatch), order="F")
    for k in range(nbatch):
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        x[:, :8, k] = (WATMETH_WATER @ rot().T).T
        x[:, 8:, k] = (WATMETH_METHANE @ rot().T).T + (u * rng.uniform(rmin, rmax))[:, None]
    return x


def thermal_dimer_geometries(nbatch, seed=0, sigma=0.05):
    """synthetic water-dimer geometries (bohr): the golden geometry plus Gaussian displacements"""
    rng = np.random.default_rng(seed)
    base = GOLDEN_GEOM_ANG / 0.529177
    x = base[None, :] + rng.normal(0.0, sigma, size=(nbatch, 18))
    return np.asfortranarray(x.T.reshape(3, 6, nbatch, order="F"))


def random_dimer_geometries(nbatch, seed=0, rmin=4.2, rmax=14.0, sigma=0.12):
    """water dimers (bohr, x(3,6,nbatch) F-order) over the whole range a ring polymer can reach: each monomer of the golden
    geometry moved to its centre of mass, rotated at random and distorted by N(0, sigma) per coordinate, the second one
    placed at a random direction rmin..rmax bohr away (close contacts on the repulsive wall up to the damped long range)"""
    rng = np.random.default_rng(seed)
    base = (GOLDEN_GEOM_ANG / 0.529177).reshape(6, 3)
    m = np.array([15.9949146221, 1.0078250321, 1.0078250321])
    x = np.empty((3, 6, nbatch), order="F")
    for k in range(nbatch):
        for mono in range(2):
            a = base[3 * mono:3 * mono + 3]
            a = a - (m[:, None] * a).sum(0) / m.sum()
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            a = a @ q.T + rng.normal(scale=sigma, size=(3, 3))
            if mono == 1:
                d = rng.normal(size=3)
                a = a + d / np.linalg.norm(d) * rng.uniform(rmin, rmax)
            x[:, 3 * mono:3 * mono + 3, k] = a.T
    return x


# ---- malonaldehyde (pes_malonaldehyde.f90) -----------------------------------------------------------------------
# the minimum-energy structure the reference file lists in its header (pes_malonaldehyde.f90:12-21), Angstrom,
# atom order C C O C O H H H H; the surface takes bohr
MALON_MIN_ANG = np.array([[0.0035239647, 0.0, -1.1379095138], [1.1859410623, 0.0, -0.4671627095], [1.2951373352, 0.0, 0.8494123252],
                          [-1.2326642069, 0.0, -0.3950488380], [-1.2836327508, 0.0, 0.8382056815], [-0.0071493753, 0.0, -2.2163348766],
                          [0.3688205810, 0.0, 1.1962401614], [-2.1703435327, 0.0, -0.9688628849], [2.1404514581, 0.0, -0.9796660170]])
MALON_BOHR = 0.52917721092
MALON_MASS = np.array([12.0, 12.0, 15.9949, 12.0, 15.9949, 1.00783, 1.00783, 1.00783, 1.00783]) * 1822.888


def malon_geometries(nbatch, seed=0, sigma=0.08):
    """x(3,9,nbatch) F-order, bohr: the minimum-energy structure, randomly rotated, every coordinate displaced by N(0, sigma)"""
    rng = np.random.default_rng(seed)
    x = np.empty((3, 9, nbatch), order="F")
    for k in range(nbatch):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        x[:, :, k] = (MALON_MIN_ANG / MALON_BOHR @ q.T).T + rng.normal(scale=sigma, size=(3, 9))
    return x
