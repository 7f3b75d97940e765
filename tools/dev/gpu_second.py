import sys, time, os, ctypes, numpy as np
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import Oracle, thermal_dimer_geometries, GOLDEN_GEOM_ANG
orc = Oracle()
pk.init()

def run_case(name, n, ntraj, nsteps, thermostat, beta, mass, a, b, sigma, Noutput=100000, cayley=False):
    pes = pk.McmodMass(name).V_init(); orc.select(name)
    nd, na = pes.ndim, pes.natom
    vi = pk.VerletInt(pes, n, mass, beta, dt=1e-3, NMC=nsteps, Noutput=Noutput, cayley=cayley, seed=1234).init_nm()
    rng = np.random.default_rng(5)
    # straight-line path a->b(traj); beads on the line + noise
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
        orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, cayley, True)
        orc.init_nm(a, bb[..., t]); orc.set_rng(1234, int(gid[t]))
        xo, po, do = orc.propagate(thermostat, x[..., t], p[..., t], dbdl[..., t], nsteps, 0, Noutput)
        ex = np.abs(xg[..., t] - xo).max() / np.abs(xo).max(); ep = np.abs(pg[..., t] - po).max() / np.abs(po).max()
        ed = abs(dg[t] - do) / max(abs(do), 1e-300)
        errs.append((ex, ep, ed))
    errs = np.array(errs)
    print("%-9s thermo=%d n=%d traj=%d steps=%d  rel err x %.2e p %.2e dHdr %.2e" % (name, thermostat, n, ntraj, nsteps, errs[:,0].max(), errs[:,1].max(), errs[:,2].max()))

mO, mH = 15.9949146221*1822.888486, 1.0078250321*1822.888486
a1 = np.array([[-1.0]]); b1 = np.array([[1.0]])
run_case("1d", 16, 3, 20, 2, 10.0, [1.0], a1, b1, 0.05)
run_case("1d", 16, 3, 20, 1, 10.0, [1.0], a1, b1, 0.05, Noutput=5)
a2 = np.array([[3.0],[0.0]]); b2 = np.array([[1.5],[2.598]])
run_case("2dtest", 24, 4, 20, 2, 10.0, [1.0], a2, b2, 0.05)
run_case("2dtest", 24, 4, 20, 1, 10.0, [1.0], a2, b2, 0.05, Noutput=4)
run_case("2dtest", 24, 2, 10, 2, 10.0, [1.0], a2, b2, 0.05, cayley=True)
g0 = (GOLDEN_GEOM_ANG/0.529177).reshape(6,3).T.copy()
g1 = g0.copy(); g1[:,[4,5]] = g0[:,[5,4]]
run_case("ccpol8sf", 6, 2, 4, 2, 12000.0, [mO,mH,mH,mO,mH,mH], g0, g1, 0.01)
run_case("ccpol8sf", 6, 2, 4, 1, 12000.0, [mO,mH,mH,mO,mH,mH], g0, g1, 0.01, Noutput=2)

# ---- throughput of the CCpol gradient kernel
import torch
pes = pk.McmodMass("ccpol8sf").V_init()
tf = ctypes.c_double(); check(lib().pimdk_fp64_peak(ctypes.byref(tf))); print("FP64 DFMA peak TFLOP/s:", tf.value)
for mode in (0, 1):
    check(lib().pimdk_set_mode(mode))
    for nb in (148*7*4, 148*7*16):
        x = torch.from_numpy(np.ascontiguousarray(thermal_dimer_geometries(nb, seed=3).reshape(18, nb, order="F").T)).cuda()
        g = torch.empty_like(x)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            check(lib().pimdk_pes_eval_dev(nb, 3, 6, x.data_ptr(), None, g.data_ptr()))
            torch.cuda.synchronize(); dt = time.time() - t0
        print("mode %d  nb=%d  grad time %.4f s  -> %.3e bead-grad/s  (%.3e energies/s)" % (mode, nb, dt, nb/dt, nb*36/dt))
