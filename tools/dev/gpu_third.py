import sys, time, os, ctypes, numpy as np
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import Oracle, thermal_dimer_geometries, GOLDEN_GEOM_ANG
orc = Oracle()
pk.init()
mO, mH = 15.9949146221*1822.888486, 1.0078250321*1822.888486

def run_case(name, n, ntraj, nsteps, thermostat, beta, mass, a, b, sigma, Noutput=100000, gamma=1.0, shape=None):
    pes = pk.McmodMass(name)
    if shape: pes.ndim, pes.natom = shape; pes.ndof = shape[0]*shape[1]
    pes.V_init()
    if shape: orc.select(name, shape[0], shape[1])
    else: orc.select(name)
    nd, na = pes.ndim, pes.natom
    vi = pk.VerletInt(pes, n, mass, beta, dt=1e-3, gamma=gamma, NMC=nsteps, Noutput=Noutput, seed=1234).init_nm()
    rng = np.random.default_rng(5)
    bb = np.asfortranarray(np.stack([a + (b - a) * (0.3 + 0.7 * t / max(1, ntraj - 1)) for t in range(ntraj)], axis=-1))
    dbdl = np.asfortranarray(np.repeat((b - a)[..., None], ntraj, axis=-1))
    x = np.empty((n, nd, na, ntraj), order="F"); p = np.empty_like(x)
    for t in range(ntraj):
        for k in range(n):
            x[k, :, :, t] = a + (bb[..., t] - a) * k / (n - 1) + rng.normal(0, sigma, size=(nd, na))
    p[...] = rng.normal(0, 1.0, size=p.shape) * np.sqrt(np.asarray(mass))[None, None, :, None] * 0.05
    gid = np.arange(ntraj, dtype=np.int64) + 7
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    xg, pg, dg = fn(x, p, a, bb, dbdl, traj_gid=gid)
    errs = []
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, gamma, 1e-3, False, True)
        orc.init_nm(a, bb[..., t]); orc.set_rng(1234, int(gid[t]))
        xo, po, do = orc.propagate(thermostat, x[..., t], p[..., t], dbdl[..., t], nsteps, 0, Noutput)
        ex = np.abs(xg[..., t] - xo).max() / np.abs(xo).max(); ep = np.abs(pg[..., t] - po).max() / np.abs(po).max()
        ed = abs(dg[t] - do) / max(abs(do), 1e-300)
        errs.append((ex, ep, ed))
    errs = np.array(errs)
    print("%-9s thermo=%d n=%d traj=%d steps=%d gamma=%g rel err x %.2e p %.2e dHdr %.2e |p|max %.2e" % (name, thermostat, n, ntraj, nsteps, gamma, errs[:,0].max(), errs[:,1].max(), errs[:,2].max(), np.abs(pg).max()))

g0 = (GOLDEN_GEOM_ANG/0.529177).reshape(6,3).T.copy()
g1 = g0.copy(); g1[:,[4,5]] = g0[:,[5,4]]
M = [mO,mH,mH,mO,mH,mH]
run_case("1d", 6, 2, 4, 2, 12000.0, M, g0, g1, 0.01, shape=(3,6))
run_case("1d", 6, 2, 4, 1, 12000.0, M, g0, g1, 0.01, Noutput=2, shape=(3,6))
run_case("1d", 6, 2, 4, 2, 10.0, M, g0, g1, 0.01, shape=(3,6))
for steps in (1, 2, 4):
    run_case("ccpol8sf", 6, 2, steps, 2, 12000.0, M, g0, g1, 0.01, gamma=0.0)
run_case("ccpol8sf", 6, 2, 4, 2, 12000.0, M, g0, g1, 0.01)
run_case("ccpol8sf", 6, 2, 4, 2, 100.0, M, g0, g1, 0.01)
run_case("ccpol8sf", 6, 2, 4, 1, 100.0, M, g0, g1, 0.01, Noutput=2)
