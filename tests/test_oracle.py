"""CPU tests of the oracle (oracle/): pinned against the reference's only golden vectors and against
mathematical identities.  No GPU needed."""
import json
import math
import os

import ctypes

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle_lib import (GOLDEN_GEOM_ANG, GOLDEN_VAL, GOLDEN_VALM, REF_DATA, SAPT_FOR_SURF, Oracle,
                        thermal_dimer_geometries)

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def orc():
    return Oracle()


# ---- the reference's own known-answer vectors: main_CCpol-8sf.f:180-189 (test_parameters) ----------
@pytest.mark.parametrize("isurf", range(1, 11))
def test_golden_interaction_energy(orc, isurf):
    orc.load_ccpol(isurf, 0)
    e = orc.ccpol_energy_ang(GOLDEN_GEOM_ANG)
    assert abs(e - GOLDEN_VAL[isurf - 1]) < 5.1e-6  # printed f10.5


@pytest.mark.parametrize("isurf", range(1, 11))
def test_golden_energy_with_monomers(orc, isurf):
    """valm(1:10) were produced by a build WITHOUT -r8 (PJT2's default-REAL literals in single
    precision); with the repo makefile's -r8 the same code gives values 1.7e-5 kcal/mol lower."""
    orc.load_ccpol(isurf, 1)
    orc.L.orc_ccpol_set_pjt2_r8(0)
    e = orc.ccpol_energy_ang(GOLDEN_GEOM_ANG)
    assert abs(e - GOLDEN_VALM[isurf - 1]) < 5.1e-6
    orc.L.orc_ccpol_set_pjt2_r8(1)
    e8 = orc.ccpol_energy_ang(GOLDEN_GEOM_ANG)
    assert abs(e8 - GOLDEN_VALM[isurf - 1]) < 3e-5 and abs(e8 - e) > 1e-5


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference data files not present on this box")
@pytest.mark.parametrize("isurf", [1, 3, 8, 10])
def test_text_and_packed_loaders_agree(orc, isurf):
    orc.load_ccpol(isurf, 1, text_dir=REF_DATA)
    a = orc.tables_image()
    orc.load_ccpol(isurf, 1)
    assert a == orc.tables_image()


def test_frozen_vectors(orc):
    """oracle bits are frozen in tests/golden/oracle_vectors.json (tests/golden/make_golden.py)"""
    G = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))
    for name, shape in (("ccpol8sf", (3, 6)), ("2dtest", (2, 1)), ("1d", (1, 1))):
        orc.select(name)
        d = G[name]
        nb = d["nbatch"]
        x = np.array([float.fromhex(h) for h in d["x"]]).reshape(shape + (nb,), order="F")
        v, g, xd = orc.pes_eval(x)
        assert [float(t).hex() for t in v] == d["v"]
        assert [float(t).hex() for t in g.reshape(-1, order="F")] == d["grad"]
        if "x_after_vprime" in d:
            assert [float(t).hex() for t in xd.reshape(-1, order="F")] == d["x_after_vprime"]
    for r in G["normals"]:
        assert float(orc.L.orc_normal(r["seed"], r["stream"], r["step"], r["gid"], r["idx"])).hex() == r["z"]


def test_opcount(orc):
    orc.load_ccpol(3, 1)
    c, e = orc.ccpol_opcount(GOLDEN_GEOM_ANG)
    assert abs(e - orc.ccpol_energy_ang(GOLDEN_GEOM_ANG)) == 0.0
    # the census bench.py's roofline uses (FLOP_PER_ENERGY)
    assert c == {"add": 25156, "mul": 33156, "div": 1760, "sqrt": 806, "exp": 1129, "pow": 33, "trig": 28}


def test_fd_gradient_is_central_difference(orc):
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(2, seed=3)
    _, g, xd = orc.pes_eval(x)
    for b in range(2):
        for i in range(3):
            for j in range(6):
                xp = x[:, :, b:b + 1].copy(order="F")
                xm = xp.copy(order="F")
                xp[i, j, 0] += 1e-4
                xm[i, j, 0] -= 1e-4
                vp = orc.pes_eval(xp, gradient=False)[0][0]
                vm = orc.pes_eval(xm, gradient=False)[0][0]
                assert abs((vp - vm) / 2e-4 - g[i, j, b]) < 2e-8 * np.abs(g[:, :, b]).max()
    assert np.abs(xd - x).max() < 1e-15 * 10  # in-place perturbation drift is at the ulp level


# ---- shared deterministic math policy vs mpmath ------------------------------------------------------
def test_detmath_accuracy():
    import ctypes
    import subprocess
    import tempfile

    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    root = os.path.dirname(HERE)
    src = os.path.join(tempfile.mkdtemp(), "dm.c")
    with open(src, "w") as f:
        f.write('#include "pimdk_detmath.h"\n'
                "double dm_exp(double x){return pimdk_exp(x);} double dm_log(double x){return pimdk_log(x);}\n"
                "double dm_pow(double x,double y){return pimdk_pow(x,y);} double dm_sin(double x){return pimdk_sin(x);}\n"
                "double dm_cos(double x){return pimdk_cos(x);} double dm_acos(double x){return pimdk_acos(x);}\n"
                "double dm_tanh(double x){return pimdk_tanh(x);} double dm_atan(double x){return pimdk_atan(x);}\n")
    so = src[:-2] + ".so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(root, "include"), "-o", so,
                    src, "-lm"], check=True)
    L = ctypes.CDLL(so)
    for f in "exp log sin cos acos tanh atan".split():
        getattr(L, "dm_" + f).restype = ctypes.c_double
        getattr(L, "dm_" + f).argtypes = [ctypes.c_double]
    L.dm_pow.restype = ctypes.c_double
    L.dm_pow.argtypes = [ctypes.c_double] * 2
    rng = np.random.default_rng(7)

    def ulps(got, exact):
        return float(abs(mp.mpf(got) - exact) / mp.mpf(np.spacing(abs(float(exact)))))

    # 10 000+ points per function over the ranges the kernels reach (exp down to the flush-to-zero edge and around 0, acos up to
    # the end points, tanh and atan into their tails).  Tolerances sit just above the worst error found (in ulp: exp 0.99,
    # log 1.29, sin 1.10, cos 1.28, acos 1.11, tanh 3.15, atan 2.92; pow 2.2 / 2.0 / 3.9 for the three exponents the path uses):
    # oracle and kernels share this header, so this test is the only guard of the functions themselves.
    N = 10000
    cases = [("exp", mp.exp, np.concatenate([rng.uniform(-700, 50, N), rng.uniform(-1, 1, N // 2), rng.uniform(-1e-3, 1e-3, N // 4)]), 1.0),
             ("log", mp.log, np.exp(rng.uniform(-30, 30, N)), 1.5),
             ("sin", mp.sin, rng.uniform(-7, 7, N), 1.5), ("cos", mp.cos, rng.uniform(-7, 7, N), 1.5),
             ("acos", mp.acos, np.concatenate([rng.uniform(-1, 1, N), 1 - np.exp(rng.uniform(-30, 0, N // 4)),
                                               -1 + np.exp(rng.uniform(-30, 0, N // 4))]), 1.5),
             ("tanh", mp.tanh, np.concatenate([rng.uniform(-3, 3, N), rng.uniform(-20, 20, N // 4)]), 3.5),
             ("atan", mp.atan, np.concatenate([rng.uniform(-3, 3, N), rng.uniform(-1e4, 1e4, N // 2)]), 3.2)]
    for name, ref, xs, tol in cases:
        worst = max(ulps(getattr(L, "dm_" + name)(float(x)), ref(mp.mpf(float(x)))) for x in xs)
        assert worst < tol, (name, worst)
    for y, tol in ((-1.5, 2.5), (-3.0, 2.5), (0.66666666666666666, 4.5)):
        worst = max(ulps(L.dm_pow(float(x), y), mp.power(mp.mpf(float(x)), mp.mpf(y))) for x in np.exp(rng.uniform(-8, 3, N)))
        assert worst < tol, (y, worst)


def test_philox_known_answer(orc):
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors: zero and all-ones counter/key)."""
    import ctypes

    # reach the raw generator through the normal: reproduce Box-Muller from the published words
    def normal_from_words(r):
        u1 = (((r[0] << 32) | r[1]) >> 11) + 0.5
        u2 = (((r[2] << 32) | r[3]) >> 11) + 0.5
        u1 *= 2.0 ** -53
        u2 *= 2.0 ** -53
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586 * u2)

    kat0 = [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]  # ctr=0 key=0
    z = orc.L.orc_normal(0, 0, 0, 0, 0)
    assert abs(z - normal_from_words(kat0)) < 1e-13
    kat1 = [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]  # ctr=ff..f key=ff..f
    seed = 0xffffffffffffffff
    step = (0xffffff << 32) | 0xffffffff
    z = orc.L.orc_normal(ctypes.c_ulonglong(seed), 0xff, ctypes.c_ulonglong(step), 0xffffffff, ctypes.c_ulonglong(0x1fffffffe))
    assert abs(z - normal_from_words(kat1)) < 1e-13


# ---- module verletint: identities (parity is UNPINNED by the reference) ------------------------------
def test_transmatrix_identities(orc):
    orc.select("1d")
    n, betan = 33, 0.37
    orc.nm_setup(n, [1.0], betan)
    orc.init_nm(np.array([[-1.0]]), np.array([[1.0]]))
    T, lam, bm, bv = orc.get_nm()
    assert np.abs(T @ T - np.eye(n)).max() < 1e-13 and np.abs(T - T.T).max() == 0.0
    # T diagonalises the open-chain spring matrix with eigenvalues (betan*lam_k)^2
    K = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = T @ K @ T
    assert np.abs(D - np.diag((lam * betan) ** 2)).max() < 1e-12
    # beadvec = T * (spring-force image of the fixed ends) / (lam betan)^2
    e = np.zeros(n)
    e[0], e[-1] = -1.0, 1.0
    assert np.abs(bv[:, 0] - (T @ e) / (lam * betan) ** 2).max() < 1e-12


def test_gauleg_and_splines(orc):
    for n in (1, 2, 5, 16):
        x, w = orc.gauleg(0.0, 1.0, n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.abs(np.sort(x) - (xr + 1) / 2).max() < 1e-13 and np.abs(w[np.argsort(x)] - wr / 2).max() < 1e-13
    from scipy.interpolate import CubicSpline

    xs = np.cumsum(np.random.default_rng(0).uniform(0.1, 1.0, 12))
    ys = np.sin(xs)
    y2 = orc.spline(xs, ys)
    cs = CubicSpline(xs, ys, bc_type="natural")
    for t in np.linspace(xs[0], xs[-1], 37):
        assert abs(orc.splint(xs, ys, y2, t) - cs(t)) < 1e-13
        assert abs(orc.splin_grad(xs, ys, y2, t) - cs(t, 1)) < 1e-12


def test_nve_energy_conservation_and_um_gradient(orc):
    orc.select("2dtest")
    n, beta = 16, 8.0
    betan = beta / (n + 1)
    a = np.array([[3.0], [0.0]])
    b = np.array([[1.5], [2.598076211353316]])
    orc.nm_setup(n, [1.0], betan, 1.0, 0.0, 1e-3)  # gamma = 0: PILE without noise or friction = NVE
    orc.init_nm(a, b)
    rng = np.random.default_rng(4)
    x = np.empty((n, 2, 1), order="F")
    for k in range(n):
        x[k, :, 0] = (a + (b - a) * (k + 1) / (n + 1))[:, 0] + rng.normal(0, 0.05, 2)
    T, lam, bm, bv = orc.get_nm()
    p = np.asfortranarray(rng.normal(0, 0.02, size=(n, 2, 1)))

    def energy(x, p):
        # H = sum_k P_k^2/(2 beadmass_k) + UM(x): fictitious-mass kinetic energy + springs + V
        P = np.stack([T @ p[:, d, 0] for d in range(2)], axis=1)
        return np.sum(P ** 2 / (2.0 * bm[0][:, None])) + orc.UM(x, a, b)

    e0 = energy(x, p)
    # time_step_pile = V(dt).NM(dt/2).O.NM(dt/2) is a first-order (Lie-Trotter) splitting when gamma=0;
    # time_step_nm = NM(dt/2).V(dt).NM(dt/2) is symmetric (second order).  Symplectic: bounded error, no drift.
    for thermostat, order in ((2, 1), (1, 2)):
        errs = []
        for dt, steps in ((1e-3, 200), (5e-4, 400)):
            orc.nm_setup(n, [1.0], betan, 1.0, 0.0, dt)
            orc.init_nm(a, b)
            x1, p1, _ = orc.propagate(thermostat, x, p, np.asfortranarray(b - a), steps, 0, 10 ** 9)
            errs.append(abs(energy(x1, p1) - e0))
        assert errs[0] < 1e-3 * abs(e0), errs
        assert 0.8 * 2 ** order < errs[0] / errs[1] < 1.25 * 2 ** order, (thermostat, errs)
    # UMprime is the gradient of UM
    g = orc.UMprime(x, a, b)
    g2, f = orc.UMforceenergy(x, a, b)
    assert np.abs(g - g2).max() < 1e-13 and abs(f - orc.UM(x, a, b)) < 1e-12
    for (k, d) in ((0, 0), (5, 1), (n - 1, 0)):
        xp = x.copy(order="F")
        xm = x.copy(order="F")
        xp[k, d, 0] += 1e-5
        xm[k, d, 0] -= 1e-5
        assert abs((orc.UM(xp, a, b) - orc.UM(xm, a, b)) / 2e-5 - g[k, d, 0]) < 1e-6 * max(1.0, abs(g[k, d, 0]))


def test_second_derivatives_restatement(orc):
    """Pins the oracle's Vdoubleprime / UMhessian (unpinned by the reference: no tests, no fixtures) by identities, and
    documents the two reference quirks that are restated literally."""
    # 1D: central difference of the analytic gradient -> analytic second derivative 12 x^2 - 4 (Vheight = x0 = 1)
    orc.select("1d")
    for xv in (-1.2, 0.3, 0.9):
        h, x = orc.Vdoubleprime(np.array([[xv]]))
        assert abs(h[0, 0, 0, 0] - (12 * xv * xv - 4)) < 1e-6
        assert abs(x[0, 0] - xv) < 1e-15            # in-place x + eps - 2 eps + eps: restored up to round-off
    # 2D: the reference assigns inside its loop over the wells (mcmod_2dtest.f90:69-82): only well k = m survives
    orc.select("2dtest")
    xy = np.array([[2.7], [0.4]])
    h, _ = orc.Vdoubleprime(xy)
    a0, b0, rho0, m, pi = 2.0, 0.2, 3.0, 6, 3.14159265358979
    wx, wy = rho0 * np.cos(m * 2.0 * pi / m), rho0 * np.sin(m * 2.0 * pi / m)
    dx, dy = xy[0, 0] - wx, xy[1, 0] - wy
    u = dx * dx + dy * dy
    dv = a0 * np.exp(-a0 * u) + b0 * np.exp(-b0 * u)
    d2 = -a0 ** 2 * np.exp(-a0 * u) - b0 ** 2 * np.exp(-b0 * u)
    expect = np.array([[(d2 * dx + dv) * dx, (d2 * dy + dv) * dx], [(d2 * dy + dv) * dx, (d2 * dy + dv) * dy]])
    assert np.abs(h.reshape(2, 2, order="F") - expect).max() < 1e-13
    # UMhessian: equals the mass-weighted numerical Hessian of UM (central differences of the oracle's own UMprime) in
    # every entry except the bead 1 - bead 2 spring coupling, which the reference never writes (instantonmod.f90:203)
    orc.select("1d")
    n, mass, betan = 7, [1.7], 0.35
    orc.nm_setup(n, mass, betan, 1.0, 1.0, 1e-3, False, True)
    rng = np.random.default_rng(4)
    x = np.asfortranarray(rng.uniform(-1.2, 1.2, size=(n, 1, 1)))
    a, b = np.array([[-1.0]]), np.array([[1.0]])
    band = orc.UMhessian(x, False)
    dense = np.zeros((n, n))
    for c in range(n):
        for r in range(c, min(n, c + 2)):
            dense[r, c] = dense[c, r] = band[r - c, c]
    num = np.zeros((n, n))
    eps = 1e-5
    for j in range(n):
        xp, xm = x.copy(order="F"), x.copy(order="F")
        xp[j] += eps
        xm[j] -= eps
        num[:, j] = (orc.UMprime(xp, a, b) - orc.UMprime(xm, a, b)).reshape(n) / (2 * eps) / mass[0]
    mask = np.ones((n, n), bool)
    mask[0, 1] = mask[1, 0] = False
    assert np.abs(dense - num)[mask].max() < 1e-5 * np.abs(num).max()
    assert dense[1, 0] == 0.0 and abs(num[1, 0] + 1.0 / betan ** 2) < 1e-5 / betan ** 2


def test_oracle_ti_ratio_matches_exact_path_integral(orc):
    """The oracle's verletint restatement (init_path, PILE step, estimator) is parity-unpinned by the reference (no
    fixtures, clock-seeded RNG); this pins it to the mathematics instead: ln(q/q0) of a 1D run against the transfer-
    matrix value of tests/exact_pi.py, within 4.5 standard errors."""
    import exact_pi
    from pimd_tunneling_b200 import path as P

    orc.select("1d")
    n, beta, dt, NMC, imin, nint, nrep = 16, 3.0, 5e-3, 20000, 1000, 8, 20
    betan = beta / (n + 1)
    a, b = np.array([[-1.0]]), np.array([[1.0]])
    lam, path, spl = P.build_path(np.stack([a, b], axis=0))
    xi, w = np.polynomial.legendre.leggauss(nint)
    xi, w = (xi + 1) / 2, w / 2
    xint, dbd = P.endpoints(lam, path, spl, xi)
    I = np.empty((nint, nrep))
    for k in range(nint):
        for r in range(nrep):
            orc.nm_setup(n, [1.0], betan, 1.0, 1.0, dt)
            orc.init_nm(a, xint[..., k])
            orc.set_rng(99, 5000 + k * nrep + r)
            x0, p0 = orc.init_path(float(xi[k]), lam, path, spl)
            _, _, d = orc.propagate(2, x0, p0, dbd[..., k], NMC, imin)
            I[k, r] = d / betan ** 2
    m, s = I.mean(axis=1), I.std(axis=1, ddof=1) / np.sqrt(nrep)
    got, se = -betan * np.sum(w * m), betan * np.sqrt(np.sum(w ** 2 * s ** 2))
    exact = exact_pi.log_ratio_1d(-1.0, 1.0, n, beta)
    assert se < 0.12 and abs(got - exact) < 4.5 * se, (got, exact, se)


def test_oracle_instanton_matches_the_analytic_kink(orc):
    """UM / UMprime / UMhessian of the oracle (parity-unpinned by the reference) pinned to the analytic instanton of
    V = (x^2-1)^2: kink action (4/3) sqrt(2m), one-loop splitting 2 w sqrt(6 S/pi) exp(-S)."""
    from scipy.linalg import eigvals_banded
    from scipy.optimize import fmin_l_bfgs_b

    orc.select("1d")
    orc.set_V0(0.0)
    m, n, beta = 20.0, 256, 40.0
    betan = beta / n
    a, b = np.array([[-1.0]]), np.array([[1.0]])
    orc.nm_setup(n, [m], betan, 1.0, 1.0, 1e-3, False, True)
    x0 = np.empty((n, 1, 1), order="F")
    for i in range(n):
        x0[i] = a + (b - a) * i / (n - 1)

    def fg(v):
        g_, f_ = orc.UMforceenergy(v.reshape(x0.shape, order="F"), a, b)
        return f_, g_.reshape(-1, order="F")

    xs, _, _ = fmin_l_bfgs_b(fg, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-7, maxls=40, maxiter=20000)
    xt = np.asfortranarray(xs.reshape(x0.shape, order="F"))
    xh = np.empty_like(xt)
    xh[:] = a.reshape(1, 1, 1)
    e0 = eigvals_banded(orc.UMhessian(xh, True), lower=True)
    e1 = eigvals_banded(orc.UMhessian(xt, False), lower=True)[1:]
    assert (e0 > 0).all() and (e1 > 0).all()
    phi = np.exp(0.5 * (np.sum(np.log(e1)) - np.sum(np.log(e0))))
    sk = betan * orc.UM(xt, a, b)
    delta = 2.0 * np.exp(-sk) * np.sqrt(sk / (2.0 * np.pi)) / phi
    S = 4.0 / 3.0 * np.sqrt(2.0 * m)
    exact = 2.0 * np.sqrt(8.0 / m) * np.sqrt(6.0 * S / np.pi) * np.exp(-S)
    assert abs(sk - S) < 2e-4 * S and abs(delta - exact) < 2e-3 * exact, (sk, S, delta, exact)


def test_fd_gradient_against_the_analytic_gradient(orc):
    """The reference's Vprime is a central difference with eps = 1e-4 bohr (mcmod_waterdimer_ccpol.f90:40-58).  Dual
    numbers through the oracle's own templates give the analytic gradient of the same V: identical energy, translation
    invariant to 1e-11, and the finite-difference gradient sits 2.6e-8 (median; < 1e-7) of max|grad| away from it — the
    size of the difference an analytic-gradient mode would have against the reference, and the reason
    the 1e-10 contract can only be met by evaluating the reference's difference quotient bit for bit."""
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(16, seed=4)
    v, g, _ = orc.pes_eval(x)
    errs = []
    for k in range(x.shape[2]):
        va, ga = orc.ccpol_analytic_gradient(x[:, :, k])
        assert va == v[k]
        assert np.abs(ga.sum(axis=1)).max() <= 1e-11 * np.abs(ga).max()
        errs.append(np.abs(ga - g[:, :, k]).max() / np.abs(g[:, :, k]).max())
    assert 1e-9 < np.median(errs) < 1e-7 and max(errs) < 2e-7, (np.median(errs), max(errs))


@pytest.mark.parametrize("isurf,iemon", [(3, 1), (3, 0), (10, 1)])
def test_analytic_gradient_mathematics_matches_dual_numbers(orc, isurf, iemon):
    """The analytic-gradient mode of the CUDA library (pimd_tunneling_b200/csrc/ccpol_grad.cuh: hand-written adjoints of
    the site-pair sums, fixed-point derivative of the induction iteration, rigid-body reduction of the embedded monomers)
    compiled for the HOST and wired per bead by tests/agrad_host.cpp, against forward-mode dual numbers through the oracle's
    own templates (oracle/dual.hpp) — two independent routes to the derivative of the same energy expression
    (main_CCpol-8sf.f:210-380).  Energy to 1e-12, gradient to 1e-10 of max|grad| (measured: 5e-13)."""
    import ctypes

    from agrad_lib import AgradHost, _P

    orc.load_ccpol(isurf, iemon)
    ag = AgradHost(isurf, iemon)
    X = thermal_dimer_geometries(12, seed=21, sigma=0.08)
    for t in range(X.shape[2]):
        vo, go = orc.ccpol_analytic_gradient(X[:, :, t])
        x18 = np.ascontiguousarray(X[:, :, t].T.reshape(-1))
        va, ga = ag.energy_gradient(x18)
        assert abs(va - vo) <= 1e-12 * max(1.0, abs(vo))
        assert np.abs(ga.reshape(6, 3).T - go).max() <= 1e-10 * np.abs(go).max()
        if isurf == 3 and iemon == 1 and t < 3:   # and the SAPT-5s'f site model alone (stage-wise yardstick)
            A = x18 * 0.529177
            v, g = ctypes.c_double(), np.empty(18)
            orc.L.orc_sapt5sf_dual(A[:9].ctypes.data_as(_P), A[9:].ctypes.data_as(_P), ctypes.byref(v), g.ctypes.data_as(_P))
            v2, g2 = ag.sapt(A[:9], A[9:])
            # the value is a sum of ~1e3 kcal/mol of cancelling electrostatic terms: absolute tolerance
            assert abs(v2 - v.value) <= 1e-9 and np.abs(g2 - g).max() <= 1e-10 * np.abs(g).max()
    orc.load_ccpol(3, 1)


def test_so2_ring_surface_restatement(orc):
    """mcmod_so2.f90:10-84: V = omegaforce**2/2 (r - r0)**2 on a ring; gradient against a central difference; the Hessian
    literally as written there (x_i x_j omegaforce**2 r0 / r**3, without the diagonal (1 - r0/r) term)"""
    orc.select("so2")
    orc.set_so2(3.0, 2.5)
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = rng.normal(size=(2, 1)) * 2.0 + 0.5
        r = np.hypot(x[0, 0], x[1, 0])
        v, g, _ = orc.pes_eval(x.reshape(2, 1, 1))
        assert abs(v[0] - 0.5 * 9.0 * (r - 2.5) ** 2) <= 1e-14 * max(1.0, abs(v[0]))
        e = 1e-6
        for d in range(2):
            xp, xm = x.copy(), x.copy()
            xp[d, 0] += e
            xm[d, 0] -= e
            fd = (orc.pes_eval(xp.reshape(2, 1, 1))[0][0] - orc.pes_eval(xm.reshape(2, 1, 1))[0][0]) / (2 * e)
            assert abs(fd - g[d, 0, 0]) < 1e-7 * max(1.0, abs(fd))
    orc.set_so2(10000.0, 20.0)   # the reference's own parameters: minimum on the ring r = 20
    assert orc.pes_eval(np.array([[12.0], [16.0]]).reshape(2, 1, 1))[0][0] == 0.0
    orc.select("ccpol8sf")


def test_water_methane_surface_restatement(orc):
    """watermethane.f90 (wmrb / wmrb_grad, Numerical Recipes gammp with EPS = 3e-7) behind mcmod_watmeth.f90: the
    incomplete gamma function against scipy to its own tolerance; the analytic gradient against a central difference of V
    (limited by that same 3e-7); rigid-motion invariance; a bound state of the right size (the paper's well is ~1 kcal/mol)."""
    from scipy.special import gammainc

    from oracle_lib import watmeth_geometries
    orc.L.orc_wm_gammp.restype = ctypes.c_double
    orc.L.orc_wm_gammp.argtypes = [ctypes.c_double, ctypes.c_double]
    for a in (7.0, 9.0, 11.0):
        for xx in (0.0, 0.3, 2.0, a - 0.5, a + 0.99, a + 1.01, 15.0, 40.0):
            assert abs(orc.L.orc_wm_gammp(a, xx) - gammainc(a, xx)) < 1e-6
    orc.select("watmeth")
    x = watmeth_geometries(24, seed=5)
    v, g, _ = orc.pes_eval(x)
    assert np.isfinite(v).all() and v.min() < -5e-4 and v.min() > -1e-2     # Hartree: a well of the order of 1 kcal/mol
    assert np.abs(g.sum(axis=1)).max() <= 1e-12 * np.abs(g).max()           # no net force on the pair of rigid bodies
    assert np.abs(g[:, 7, :]).max() == 0.0                                  # the water O site carries no interaction
    e = 1e-4
    for k in range(4):
        for (d, s) in ((0, 0), (2, 3), (1, 12), (2, 16)):
            xp, xm = x[:, :, k:k + 1].copy(order="F"), x[:, :, k:k + 1].copy(order="F")
            xp[d, s, 0] += e
            xm[d, s, 0] -= e
            fd = (orc.pes_eval(xp)[0][0] - orc.pes_eval(xm)[0][0]) / (2 * e)
            assert abs(fd - g[d, s, k]) < 2e-5 * np.abs(g[:, :, k]).max() + 1e-9
    orc.select("ccpol8sf")


def test_malonaldehyde_surface_restatement(orc):
    """pes_malonaldehyde.f90 (`pes`, iopt 0/1/2) behind mcmod_malon.f90.  Reference-held known answer: the file's header lists
    the minimum-energy structure (:12-21) and defines the energy as "above equilibrium" — V vanishes there and so does the
    gradient (to the 10 printed decimals of the coordinates).  Then: analytic gradient against a central difference of V,
    analytic Hessian against a central difference of the gradient, symmetry, rigid-motion invariance, V0."""
    from oracle_lib import MALON_BOHR, MALON_MIN_ANG, malon_geometries
    orc.select("malon")
    xmin = np.asfortranarray((MALON_MIN_ANG / MALON_BOHR).T.reshape(3, 9, 1))
    v, g, _ = orc.pes_eval(xmin)
    assert abs(v[0]) < 1e-12 and np.abs(g).max() < 1e-8
    x = malon_geometries(6, seed=2)
    v, g, _ = orc.pes_eval(x)
    assert (v > 0).all() and v.max() < 0.5                                   # Hartree above the minimum
    assert np.abs(g.sum(axis=1)).max() <= 1e-12 * np.abs(g).max()           # translation invariance
    e = 1e-5
    x0 = np.array(x[:, :, 0], order="F")
    num = np.zeros((3, 9))
    for d in range(3):
        for s in range(9):
            xp, xm = x0.reshape(3, 9, 1).copy(order="F"), x0.reshape(3, 9, 1).copy(order="F")
            xp[d, s, 0] += e
            xm[d, s, 0] -= e
            num[d, s] = (orc.pes_eval(xp, gradient=False)[0][0] - orc.pes_eval(xm, gradient=False)[0][0]) / (2 * e)
    assert np.abs(num - g[:, :, 0]).max() < 1e-8 * np.abs(g[:, :, 0]).max()
    h = orc.Vdoubleprime(x0.copy(order="F"))[0].reshape(27, 27, order="F")
    assert np.array_equal(h, h.T)
    numh = np.zeros((27, 27))
    for d in range(27):
        xp, xm = x0.reshape(-1, order="F").copy(), x0.reshape(-1, order="F").copy()
        xp[d] += e
        xm[d] -= e
        gp = orc.pes_eval(xp.reshape(3, 9, 1, order="F"), energy=False)[1][:, :, 0].reshape(-1, order="F")
        gm = orc.pes_eval(xm.reshape(3, 9, 1, order="F"), energy=False)[1][:, :, 0].reshape(-1, order="F")
        numh[d] = (gp - gm) / (2 * e)
    assert np.abs(numh - h).max() < 1e-8 * np.abs(h).max()
    orc.set_V0(0.25)                                                         # mcmod_malon.f90:21: V = V - V0
    assert orc.pes_eval(x[:, :, :1], gradient=False)[0][0] == v[0] - 0.25
    orc.set_V0(0.0)
    orc.select("ccpol8sf")


@pytest.mark.skipif(not os.path.exists("/root/reference/pes_malonaldehyde.f90"), reason="the reference tree is not on this box")
def test_malonaldehyde_table_file_is_the_reference_data(tmp_path):
    """pimd_tunneling_b200/data/malonaldehyde.tbl is exactly what tools/pack_malon_tables.py makes of the reference's DATA
    statements (run where /root/reference exists; the GPU box only has the packed file)."""
    import subprocess
    import sys
    out = tmp_path / "m.tbl"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pack_malon_tables.py"), "/root/reference/pes_malonaldehyde.f90", str(out)],
                   check=True, capture_output=True)
    assert out.read_bytes() == open(os.path.join(ROOT, "pimd_tunneling_b200", "data", "malonaldehyde.tbl"), "rb").read()
